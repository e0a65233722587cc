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
