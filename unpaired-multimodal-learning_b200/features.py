"""Feature-bank files: the on-disk contract between the (offline) encoders and the training path.

The reference's ``vision_language/features.py`` runs frozen CLIP / timm / HF encoders and saves their
outputs; the encoders are out of scope here (frozen, offline, need downloaded weights), the FILE
LAYOUT is not - it is what ``finetune.py`` consumes:

  text   {feature_dir}/text/{enc}/{dataset}/{text_type}.pth
         {'features' [Nt,D], 'labels' [Nt], 'eot_indices' [Nt], 'prompts', 'lab2cname'}   (features.py:38-44,96-103)
  image  {feature_dir}/image/{enc}/{dataset}/{aug}/shot_{k}-seed_{s}.pth
         {'train': {'features','labels','paths'}, 'val': {...}, 'lab2cname'}               (features.py:32-36,239-246)
  test   {feature_dir}/image/{enc}/{dataset}/test.pth
         {'features','labels','paths','lab2cname'}
with ``enc = encoder.replace('/', '-')``.  ``write_*`` produce files the reference can read back;
``load_*`` accept files the reference wrote.
"""
import os

import torch

from .engine.datasets.utils import get_few_shot_setup_name
from .engine.tools.utils import makedirs


def img_outdir(outdir, encoder, ds, augmentation, tr_shot, seed, mode="train", return_tokens=False):
    sub = "patch-token" if return_tokens else ""
    enc = encoder.replace("/", "-")
    if mode == "train":
        return os.path.join(outdir, sub, "image", enc, ds, augmentation, f"{get_few_shot_setup_name(tr_shot, seed)}.pth")
    return os.path.join(outdir, sub, "image", enc, ds, "test.pth")


def text_outdir(outdir, encoder, ds, text_augmentation, return_tokens=False):
    sub = "patch-token" if return_tokens else ""
    return os.path.join(outdir, sub, "text", encoder.replace("/", "-"), ds, f"{text_augmentation}.pth")


def descriptor_outdir(outdir, encoder, ds, descriptor_type, return_tokens=False):
    return text_outdir(outdir, encoder, ds, descriptor_type, return_tokens)


def _check_rows(d, what):
    f, l = d["features"], d["labels"]
    if f.dim() != 2 or l.dim() != 1 or f.shape[0] != l.shape[0]:
        raise ValueError(f"{what}: expected features [N,D] and labels [N], got {tuple(f.shape)} / {tuple(l.shape)}")
    return d


def load_text_bank(path):
    """The text bank dict of the reference's layout (features.py:152-184 there).  When a fresh v2 file sits next to it
    (``<stem>.bank2``, ``convert_bank``) the rows come from that file - no unpickling - together with its class-sorted row
    index (``class_order`` / ``class_starts``), which ``TextTensorDataset`` consumes for n-shot selection / averaging."""
    v2 = os.path.splitext(path)[0] + ".bank2"
    if os.path.exists(v2) and (not os.path.exists(path) or os.path.getmtime(v2) >= os.path.getmtime(path)):
        t, _, meta = load_bank_v2(v2, "cpu", sections=("features", "labels", "eot_indices", "class_order", "class_starts"))
        d = dict(meta or {})
        d.update(t)
        if "eot_indices" not in d:
            d["eot_indices"] = torch.zeros(d["labels"].shape[0], dtype=torch.int64)
        return _check_rows(d, v2)
    d = torch.load(path, map_location="cpu")
    for k in ("features", "labels", "eot_indices"):
        if k not in d:
            raise KeyError(f"{path}: text bank lacks '{k}'")
    return _check_rows(d, path)


def load_image_bank(path):
    d = torch.load(path, map_location="cpu")
    if "train" in d:
        _check_rows(d["train"], path + "[train]")
        _check_rows(d["val"], path + "[val]")
    else:
        _check_rows(d, path)
    return d


def write_text_bank(path, features, labels, eot_indices=None, prompts=None, lab2cname=None):
    makedirs(os.path.dirname(path))
    eot = torch.zeros(labels.shape[0], dtype=torch.int64) if eot_indices is None else eot_indices
    torch.save({"features": features.float().cpu(), "labels": labels.long().cpu(), "eot_indices": eot.long().cpu(),
                "prompts": prompts if prompts is not None else {}, "lab2cname": lab2cname}, path)


def write_image_bank(path, train=None, val=None, test=None, lab2cname=None):
    """train/val/test are (features, labels[, paths]) tuples.  With ``test`` writes the flat test layout."""
    makedirs(os.path.dirname(path))

    def pack(t):
        return {"features": t[0].float().cpu(), "labels": t[1].long().cpu(),
                "paths": list(t[2]) if len(t) > 2 else [""] * t[1].shape[0]}

    if test is not None:
        d = pack(test)
    else:
        d = {"train": pack(train), "val": pack(val)}
    d["lab2cname"] = lab2cname
    torch.save(d, path)


# ------------------------------------------------------------------------------------------------
# bank format v2: one flat, memory-mappable file per bank
# ------------------------------------------------------------------------------------------------
# The v1 files above are pickled dicts of fp32 tensors: torch.load materialises a second copy on the host and the
# per-class operations of TextTensorDataset scan all labels once per class.  v2 is what the HBM-resident banks want:
#
#   [0, 4096)            header: magic "UMLBANK2", then a JSON object (padded with spaces)
#   features             [N, D] fp32, row-major            (4096-byte aligned)
#   features_bf16        [N, D] bf16, optional             (the tensor-core path's shadow bank, written once)
#   labels               [N] int64
#   class_order          [N] int64   row indices sorted by (label, row) - a stable argsort
#   class_starts         [C + 1] int64  class c owns class_order[class_starts[c] : class_starts[c + 1]]
#   eot_indices          [N] int64, optional (text banks)
# Everything else a v1 file carries (paths, prompts, lab2cname ...) goes to a small pickled side-car "<file>.meta".
# Loading maps the file and copies section by section through pinned memory; nothing is parsed or re-laid out.
# convert_bank() rewrites a v1 file loss-lessly (features/labels bit-identical, extras preserved).
import json
import mmap
import warnings

import numpy as np

_MAGIC = b"UMLBANK2"
_ALIGN = 4096


def _np_dtype(name):
    return {"float32": np.float32, "int64": np.int64, "bfloat16": np.uint16}[name]


def write_bank_v2(path, features, labels, eot_indices=None, meta=None, with_bf16=True):
    """Write one bank (any split) in the v2 layout.  ``features`` [N, D] (stored as fp32, plus a bf16 copy when
    ``with_bf16``), ``labels`` [N]; ``meta`` is any picklable dict (paths, prompts, lab2cname ...)."""
    feats = features.detach().to("cpu", torch.float32).contiguous()
    labs = labels.detach().to("cpu", torch.int64).contiguous()
    if feats.dim() != 2 or labs.dim() != 1 or feats.shape[0] != labs.shape[0]:
        raise ValueError("write_bank_v2: expected features [N, D] and labels [N]")
    n, d = feats.shape
    n_classes = int(labs.max()) + 1 if n else 0
    order = torch.argsort(labs, stable=True)
    starts = torch.zeros(n_classes + 1, dtype=torch.int64)
    if n:
        starts[1:] = torch.cumsum(torch.bincount(labs, minlength=n_classes), 0)
    sections = [("features", feats.numpy(), "float32", [n, d])]
    if with_bf16:
        sections.append(("features_bf16", feats.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16), "bfloat16", [n, d]))
    sections += [("labels", labs.numpy(), "int64", [n]), ("class_order", order.numpy(), "int64", [n]),
                 ("class_starts", starts.numpy(), "int64", [n_classes + 1])]
    if eot_indices is not None:
        sections.append(("eot_indices", eot_indices.detach().to("cpu", torch.int64).contiguous().numpy(), "int64", [n]))
    header, off = {"version": 2, "rows": n, "dim": d, "classes": n_classes, "sections": {}}, _ALIGN
    for name, arr, dt, shape in sections:
        header["sections"][name] = {"offset": off, "dtype": dt, "shape": shape, "bytes": int(arr.nbytes)}
        off = (off + arr.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    blob = _MAGIC + json.dumps(header).encode()
    if len(blob) > _ALIGN:
        raise ValueError("write_bank_v2: header does not fit 4096 bytes")
    makedirs(os.path.dirname(path) or ".")
    with open(path, "wb") as f:
        f.write(blob.ljust(_ALIGN, b" "))
        for name, arr, _, _ in sections:
            f.seek(header["sections"][name]["offset"])
            f.write(arr.tobytes())
        f.truncate(off)
    if meta is not None:
        torch.save(meta, path + ".meta")
    return header


def read_bank_v2_header(path):
    with open(path, "rb") as f:
        raw = f.read(_ALIGN)
    if raw[:8] != _MAGIC:
        raise ValueError(f"{path}: not a v2 bank file")
    return json.loads(raw[8:].decode().strip())


def load_bank_v2(path, device="cuda", sections=("features", "features_bf16", "labels", "class_order", "class_starts",
                                                "eot_indices"), chunk_bytes=64 << 20):
    """Map a v2 bank and bring the requested sections to ``device``: each section is copied straight out of the page
    cache in ``chunk_bytes`` pieces through a pinned staging buffer (no second host copy, no unpickling).
    Returns ``(tensors, header, meta)``; ``tensors['features_bf16']`` is a bf16 view when the file has that section."""
    header = read_bank_v2_header(path)
    dev = torch.device(device)
    out = {}
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            stage = torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) if dev.type == "cuda" else None
            for name in sections:
                sec = header["sections"].get(name)
                if sec is None:
                    continue
                flat = np.frombuffer(mm, dtype=np.uint8, count=sec["bytes"], offset=sec["offset"])
                dst = torch.empty(sec["bytes"], dtype=torch.uint8, device=dev)
                if dev.type == "cuda":
                    for lo in range(0, sec["bytes"], chunk_bytes):
                        hi = min(sec["bytes"], lo + chunk_bytes)
                        with warnings.catch_warnings():
                            warnings.simplefilter("ignore")  # read-only mapping: it is only read
                            stage[:hi - lo].copy_(torch.from_numpy(flat[lo:hi]))   # page cache -> pinned
                        dst[lo:hi].copy_(stage[:hi - lo], non_blocking=True)       # pinned -> HBM
                        torch.cuda.current_stream().synchronize()                  # the stage is reused
                else:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")  # read-only mapping: it is only read
                        dst.copy_(torch.from_numpy(flat))
                tdt = {"float32": torch.float32, "int64": torch.int64, "bfloat16": torch.bfloat16}[sec["dtype"]]
                out[name] = dst.view(tdt).view(sec["shape"])
                del flat
        finally:
            mm.close()
    meta = torch.load(path + ".meta", map_location="cpu") if os.path.exists(path + ".meta") else None
    return out, header, meta


def convert_bank(v1_path, v2_path=None, split=None, with_bf16=True):
    """Rewrite a v1 ``.pth`` bank as v2, loss-lessly.  Image train files hold two splits ('train', 'val'): pass
    ``split`` or get ``<v2_path>.train`` / ``<v2_path>.val``.  Returns the list of files written."""
    d = torch.load(v1_path, map_location="cpu")
    v2_path = v2_path or (os.path.splitext(v1_path)[0] + ".bank2")
    written = []
    if "train" in d and "val" in d:
        for sp in ([split] if split else ["train", "val"]):
            part = d[sp]
            meta = {"paths": part.get("paths"), "lab2cname": d.get("lab2cname")}
            write_bank_v2(f"{v2_path}.{sp}", part["features"], part["labels"], meta=meta, with_bf16=with_bf16)
            written.append(f"{v2_path}.{sp}")
    else:
        meta = {k: v for k, v in d.items() if k not in ("features", "labels", "eot_indices")}
        write_bank_v2(v2_path, d["features"], d["labels"], eot_indices=d.get("eot_indices"), meta=meta, with_bf16=with_bf16)
        written.append(v2_path)
    return written


def load_feature_bank(v1_path, device="cuda", split=None):
    """A ``FeatureBank`` for one split of an image bank file: from the v2 file next to it when there is one
    (``<stem>.bank2[.split]``, written by ``convert_bank``; the bf16 shadow comes along), else from the v1 dict."""
    from .engine.datasets.utils import FeatureBank
    v2 = os.path.splitext(v1_path)[0] + ".bank2" + (f".{split}" if split else "")
    # a v1 file regenerated after the conversion wins over the stale v2 cache
    if os.path.exists(v2) and (not os.path.exists(v1_path) or os.path.getmtime(v2) >= os.path.getmtime(v1_path)):
        t, _, meta = load_bank_v2(v2, device, sections=("features", "features_bf16", "labels"))
        bank = FeatureBank.__new__(FeatureBank)
        bank.features, bank.labels = t["features"], t["labels"]
        if "features_bf16" in t:
            bank._bf16 = t["features_bf16"]
        return bank, (meta or {})
    d = load_image_bank(v1_path)
    part = d[split] if split else d
    return FeatureBank(part["features"], part["labels"], device), {"lab2cname": d.get("lab2cname"), "paths": part.get("paths")}
