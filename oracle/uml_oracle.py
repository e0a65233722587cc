"""TEST INFRASTRUCTURE — CPU restatement of the reference UML hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this file.  The product (the package
``unpaired-multimodal-learning_b200``) never does; it fails loudly without its CUDA library.

What is restated (all arithmetic in fp32 on the CPU, like the reference, which never
enables AMP/TF32):

* sampler protocol      reference ``vision_language/finetune.py:33-39,157-158,370-371``
                         + torch ``RandomSampler.__iter__`` / ``_BaseDataLoaderIter.__init__``
* text bank selection   ``vision_language/engine/datasets/utils.py:48-107``
* zero-shot head init   ``vision_language/engine/models/head.py:22-37``
* head forward          ``head.py:77-84`` (UML) and ``head.py:131-137`` (UMLClip)
* loss                  ``finetune.py:186-188``
* backward              autograd of the above (closed form here)
* optimizers            ``vision_language/engine/optimizer/optim.py:15-71`` -> torch.optim
                         AdamW / Adam / SGD single-tensor update rules
* LR schedule           ``vision_language/engine/optimizer/scheduler.py:58-143``
* validate              ``finetune.py:291-315``
* train loop            ``finetune.py:157-195,247-275``
* Gaussian analogue     ``Gaussian_experiment/{main.py:31-59,model.py:5-49,dataset.py:3-18,data.py:29-61}``

Third-party arithmetic the reference leans on: PyTorch (pinned ``torch==2.8.0`` in the
reference's ``environment.yml:197``; this image has 2.11.0).  ``torch.randperm`` and
``Tensor.random_`` are used here as-is for the integer sampler stream because the
reference's index order *is* torch's mt19937 stream.

Pinning: the reference ships no tests and no golden vectors.  This file is pinned against
the reference *itself*, executed unmodified in the build container by
``oracle/ref_harness.py``; the resulting traces are committed under ``tests/golden/``
(``make_golden.py`` is the generating script) and re-checked by
``tests/test_oracle_golden.py`` everywhere, plus live by
``tests/test_oracle_vs_reference.py`` where ``/root/reference`` exists.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# a-1  sampler protocol
# --------------------------------------------------------------------------------------


def _draw_int64(generator: Optional[torch.Generator]) -> int:
    """One ``torch.empty((), int64).random_()`` draw (global generator when None)."""
    return int(torch.empty((), dtype=torch.int64).random_(generator=generator).item())


class OracleLoader:
    """Index stream of ``DataLoader(ds, batch_size, shuffle=True, drop_last=...)``.

    ``iter()`` consumes the RNG exactly as the torch loader does:
      * every new iterator draws a base seed first (``_BaseDataLoaderIter.__init__``);
      * the RandomSampler then draws its own seed and a permutation - at the first
        ``next()`` when ``num_workers == 0`` (lazy generator), but already inside
        ``iter()`` when ``num_workers > 0`` (the multi-process iterator primes its
        prefetch queue in its constructor);
      * with an explicit ``generator`` (Gaussian loader, ``main.py:141-143``) both the
        base seed and the permutation come from that generator and no sampler seed is drawn.
    """

    def __init__(self, n: int, batch_size: int, shuffle: bool = True, drop_last: bool = False,
                 num_workers: int = 0, generator: Optional[torch.Generator] = None):
        self.n, self.batch_size, self.shuffle = int(n), int(batch_size), shuffle
        self.drop_last, self.num_workers, self.generator = drop_last, num_workers, generator
        self._perm: Optional[torch.Tensor] = None
        self._pos = 0
        self._tail_drawn = False

    def _draw_perm(self):
        if not self.shuffle:
            self._perm = torch.arange(self.n)
        elif self.generator is None:
            g = torch.Generator()
            g.manual_seed(_draw_int64(None))
            self._perm = torch.randperm(self.n, generator=g)
        else:
            self._perm = torch.randperm(self.n, generator=self.generator)
        self._pos = 0

    def iter(self):
        self._perm = None
        self._tail_drawn = False
        _draw_int64(self.generator)  # base seed (value unused when there are no workers)
        if self.num_workers > 0:
            self._draw_perm()
        return self

    def next(self) -> Optional[torch.Tensor]:
        """Next index batch, or None when the epoch is exhausted (StopIteration)."""
        if self._perm is None:
            self._draw_perm()
        left = self.n - self._pos
        if left <= 0 or (self.drop_last and left < self.batch_size):
            if self.shuffle and self.generator is not None and not self._tail_drawn:
                # RandomSampler.__iter__ ends with `randperm(n)[: num_samples % n]`: an empty slice,
                # but the permutation is still drawn - from the caller's generator when one is given.
                torch.randperm(self.n, generator=self.generator)
                self._tail_drawn = True
            return None
        take = min(self.batch_size, left)
        out = self._perm[self._pos:self._pos + take]
        self._pos += take
        return out


def fetch_next_indices(loader: OracleLoader) -> torch.Tensor:
    """``fetch_next`` (finetune.py:33-39): on exhaustion re-iter, then take the first batch."""
    b = loader.next()
    if b is None:
        loader.iter()
        b = loader.next()
    return b


# --------------------------------------------------------------------------------------
# a-2  text bank selection
# --------------------------------------------------------------------------------------


def select_text_rows(features: torch.Tensor, labels: torch.Tensor, eot: torch.Tensor, n_shots):
    """TextTensorDataset.__init__: None -> all rows; int k -> k random rows per class in
    ``torch.unique`` order (one global-RNG ``randperm`` per class); 'average' -> class means."""
    if n_shots is None:
        return features, labels, eot
    if isinstance(n_shots, str):
        if n_shots.lower() != "average":
            raise ValueError("n_shots must be an int, None, or 'average'")
        classes = torch.unique(labels)
        feats = torch.stack([features[labels == c].mean(dim=0) for c in classes])
        e = torch.stack([eot[labels == c][0] for c in classes])
        return feats, classes, e
    if not isinstance(n_shots, int):
        raise ValueError("n_shots must be an int, None, or 'average'")
    keep = []
    for c in torch.unique(labels):
        rows = (labels == c).nonzero(as_tuple=True)[0]
        k = min(n_shots, rows.numel())
        keep.append(rows[torch.randperm(rows.numel())[:k]])
    keep = torch.cat(keep)
    return features[keep], labels[keep], eot[keep]


# --------------------------------------------------------------------------------------
# a-10 zero-shot init
# --------------------------------------------------------------------------------------


def zero_shot_weights(features: torch.Tensor, labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """Rows = L2-normalised class means of the text rows; classes without rows stay zero
    (F.normalize clamps the norm at 1e-12, so 0 stays 0)."""
    dim = features.shape[1]
    w = torch.zeros(num_classes, dim, dtype=torch.float32)
    for c in torch.unique(labels).tolist():
        w[int(c)] = features[labels == c].to(torch.float32).mean(dim=0)
    norm = w.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return w / norm


# --------------------------------------------------------------------------------------
# a-9  LR schedule (closed form of what torch 2.x emits for the reference's wrappers)
# --------------------------------------------------------------------------------------


def lr_at(step: int, base_lr: float, sched: str = "cosine", warmup_iter: int = 50,
          max_iter: int = 12800, warmup_type: Optional[str] = "linear",
          warmup_lr: Optional[float] = 1e-5) -> float:
    """Learning rate the optimizer *uses* at 0-based step ``step``.

    LinearWarmupScheduler.get_lr (scheduler.py:73-81): step 0 -> warmup_lr, then
    base*step/warmup; once ``last_epoch >= warmup`` the wrapped successor is stepped, and it
    starts from its own epoch 0, so the decay phase is evaluated at ``step - warmup``.
    """
    if warmup_iter > 0 and step < warmup_iter:
        if warmup_type == "constant":
            return float(warmup_lr)
        if warmup_type == "linear":
            return float(warmup_lr) if step == 0 else base_lr * step / warmup_iter
        raise ValueError(f"warmup_type {warmup_type!r}")
    t = step - warmup_iter if warmup_iter > 0 else step
    if sched == "cosine":
        return base_lr * (1.0 + math.cos(math.pi * t / float(max_iter))) / 2.0
    if sched == "linear":
        return base_lr * (1.0 - t / float(max_iter))
    raise ValueError(f"scheduler {sched!r}")


# --------------------------------------------------------------------------------------
# a-8  optimizers (torch single-tensor rules, fp32 state)
# --------------------------------------------------------------------------------------


class OracleOptimizer:
    """AdamW (decoupled decay), Adam (L2), SGD (momentum 0.9, L2, no nesterov).

    Parameters whose gradient is None are skipped entirely (no decay, no step count), as
    torch.optim does."""

    def __init__(self, params: Dict[str, torch.Tensor], name: str, lr: float, weight_decay: float,
                 betas=(0.9, 0.999), eps: float = 1e-8, momentum: float = 0.9):
        if name not in ("adamw", "adam", "sgd"):
            raise AssertionError(f"Optimizer {name} not found")
        self.params, self.name, self.lr, self.wd = params, name, lr, weight_decay
        self.betas, self.eps, self.momentum = betas, eps, momentum
        self.state: Dict[str, Dict[str, object]] = {}

    def step(self, grads: Dict[str, Optional[torch.Tensor]], lr: Optional[float] = None):
        lr = self.lr if lr is None else lr
        b1, b2 = self.betas
        for k, p in self.params.items():
            g = grads.get(k)
            if g is None:
                continue
            st = self.state.setdefault(k, {"t": 0})
            if self.name == "sgd":
                if self.wd != 0:
                    g = g + self.wd * p
                if "buf" not in st:
                    st["buf"] = g.clone()
                else:
                    st["buf"].mul_(self.momentum).add_(g)
                p.sub_(st["buf"] * lr)
                continue
            if "m" not in st:
                st["m"], st["v"] = torch.zeros_like(p), torch.zeros_like(p)
            st["t"] += 1
            t = st["t"]
            if self.name == "adamw":
                p.mul_(1.0 - lr * self.wd)
            elif self.wd != 0:
                g = g + self.wd * p
            st["m"].add_((g - st["m"]) * (1.0 - b1))
            st["v"].mul_(b2).add_(g * g * (1.0 - b2))
            bc1 = 1.0 - b1 ** t
            bc2_sqrt = math.sqrt(1.0 - b2 ** t)
            denom = st["v"].sqrt() / bc2_sqrt + self.eps
            p.sub_((lr / bc1) * st["m"] / denom)


# --------------------------------------------------------------------------------------
# a-4..a-7  head forward / CE / backward (closed form)
# --------------------------------------------------------------------------------------


def _ce_and_grad(logits: torch.Tensor, labels: torch.Tensor):
    """mean CE and d(mean CE)/d logits = (softmax - onehot)/B."""
    b = logits.shape[0]
    lse = torch.logsumexp(logits, dim=1)
    picked = logits.gather(1, labels.view(-1, 1)).squeeze(1)
    loss = (lse - picked).mean()
    g = torch.softmax(logits, dim=1)
    g[torch.arange(b), labels] -= 1.0
    g /= b
    acc = (logits.argmax(dim=1) == labels).float().mean().item()
    return loss, g, acc


@dataclass
class HeadState:
    """Parameters in the reference's registration order (head.py:63-70):
    img_proj.weight (optional), head.weight, img_scale, txt_scale (when learnable)."""
    head: torch.Tensor                       # [C, D]
    img_proj: Optional[torch.Tensor] = None  # [D, Dv]
    img_scale: float = 1.0
    txt_scale: float = 1.0
    learnable_temp: bool = False
    scales: Dict[str, torch.Tensor] = field(default_factory=dict)

    def param_dict(self) -> Dict[str, torch.Tensor]:
        d: Dict[str, torch.Tensor] = {}
        if self.img_proj is not None:
            d["img_proj.weight"] = self.img_proj
        d["head.weight"] = self.head
        if self.learnable_temp:
            if not self.scales:
                self.scales = {"img_scale": torch.tensor(float(self.img_scale)),
                               "txt_scale": torch.tensor(float(self.txt_scale))}
            d.update(self.scales)
        return d

    def s_img(self) -> float:
        return float(self.scales["img_scale"]) if self.learnable_temp and self.scales else float(self.img_scale)

    def s_txt(self) -> float:
        return float(self.scales["txt_scale"]) if self.learnable_temp and self.scales else float(self.txt_scale)


def uml_step_grads(st: HeadState, x_img: Optional[torch.Tensor], y_img: Optional[torch.Tensor],
                   x_txt: Optional[torch.Tensor], y_txt: Optional[torch.Tensor], alpha: float):
    """One forward/backward of ``loss = 1.0*CE(img) + alpha*CE(txt)`` through the shared head.
    Returns (stats, grads) with grads keyed like the state dict.  ``dW_img``/``dW_txt`` are the
    per-modality head gradients the reference extracts at finetune.py:190-191 (unweighted)."""
    W = st.head
    grads: Dict[str, Optional[torch.Tensor]] = {k: None for k in st.param_dict()}
    stats = {"image_loss": 0.0, "text_loss": 0.0, "img_acc": 0.0, "text_acc": 0.0}
    dW = torch.zeros_like(W)
    dW_img = torch.zeros_like(W)
    dW_txt = torch.zeros_like(W)
    if x_img is not None:
        z = x_img @ st.img_proj.t() if st.img_proj is not None else x_img
        raw = z @ W.t()
        s = st.s_img()
        loss, g, acc = _ce_and_grad(raw * s, y_img)
        stats["image_loss"], stats["img_acc"] = float(loss), acc
        dW_img = s * (g.t() @ z)
        dW += dW_img
        if st.img_proj is not None:
            dz = s * (g @ W)
            grads["img_proj.weight"] = dz.t() @ x_img
        if st.learnable_temp:
            grads["img_scale"] = (g * raw).sum()
    if x_txt is not None:
        raw = x_txt @ W.t()
        s = st.s_txt()
        loss, g, acc = _ce_and_grad(raw * s, y_txt)
        stats["text_loss"], stats["text_acc"] = float(loss), acc
        dW_txt = s * (g.t() @ x_txt)
        dW += alpha * dW_txt
        if st.learnable_temp:
            grads["txt_scale"] = alpha * (g * raw).sum()
    grads["head.weight"] = dW
    stats["dW_img"], stats["dW_txt"] = dW_img, dW_txt
    return stats, grads


def uml_step_grads_autograd(st: HeadState, x_img, y_img, x_txt, y_txt, alpha: float):
    """Same quantities as ``uml_step_grads`` but computed the way the reference does it -
    ``nn.functional.linear`` + ``F.cross_entropy`` and THREE backward sweeps
    (two ``autograd.grad`` diagnostics, finetune.py:190-191, then ``loss.backward``, :193).
    Used (a) to cross-check the closed form and (b) as the CPU cost model of the reference
    step in ``bench.py --impl reference``."""
    F = torch.nn.functional
    params = {k: v.detach().clone().requires_grad_(True) for k, v in st.param_dict().items()}
    W = params["head.weight"]
    stats = {"image_loss": 0.0, "text_loss": 0.0, "img_acc": 0.0, "text_acc": 0.0}
    zero = torch.tensor(0.0)
    il, tl = zero, zero
    dW_img, dW_txt = torch.zeros_like(W), torch.zeros_like(W)
    if x_img is not None:
        z = F.linear(x_img, params["img_proj.weight"]) if "img_proj.weight" in params else x_img
        s = params["img_scale"] if "img_scale" in params else st.s_img()
        logits = F.linear(z, W) * s
        il = F.cross_entropy(logits, y_img)
        stats["image_loss"] = float(il.detach())
        stats["img_acc"] = (logits.argmax(1) == y_img).float().mean().item()
    if x_txt is not None:
        s = params["txt_scale"] if "txt_scale" in params else st.s_txt()
        tlog = F.linear(x_txt, W) * s
        tl = F.cross_entropy(tlog, y_txt)
        stats["text_loss"] = float(tl.detach())
        stats["text_acc"] = (tlog.argmax(1) == y_txt).float().mean().item()
    loss = 1.0 * il + alpha * tl
    if x_img is not None:
        (dW_img,) = torch.autograd.grad(il, W, retain_graph=True)
    if x_txt is not None:
        (dW_txt,) = torch.autograd.grad(tl, W, retain_graph=True)
    loss.backward(retain_graph=True)
    grads = {k: (v.grad.detach() if v.grad is not None else None) for k, v in params.items()}
    stats["dW_img"], stats["dW_txt"] = dW_img.detach(), dW_txt.detach()
    return stats, grads


# --------------------------------------------------------------------------------------
# a-11 validate
# --------------------------------------------------------------------------------------


def validate(st: HeadState, feats: torch.Tensor, labels: torch.Tensor, batch_size: int,
             loader_protocol: bool = True) -> Tuple[float, float]:
    """(val_loss, val_acc): accuracy over all rows; loss = mean over batches of the
    batch-mean CE, so a short last batch is over-weighted (finetune.py:304-312).

    ``for batch in val_loader`` (finetune.py:295) builds a fresh DataLoader iterator, and every
    such iterator draws a base seed from the *global* generator even with shuffle=False - so each
    validate() call advances the RNG stream the training samplers later draw from."""
    if loader_protocol:
        _draw_int64(None)
    losses, hits = [], 0
    for s in range(0, feats.shape[0], batch_size):
        x, y = feats[s:s + batch_size], labels[s:s + batch_size]
        z = x @ st.img_proj.t() if st.img_proj is not None else x
        logits = (z @ st.head.t()) * st.s_img()
        lse = torch.logsumexp(logits, dim=1)
        losses.append((lse - logits.gather(1, y.view(-1, 1)).squeeze(1)).mean())
        hits += int((logits.argmax(1) == y).sum())
    return float(torch.stack(losses).mean()), hits / feats.shape[0]


# --------------------------------------------------------------------------------------
# a-12 train loop
# --------------------------------------------------------------------------------------


@dataclass
class TrainTrace:
    img_idx: List[np.ndarray] = field(default_factory=list)
    txt_idx: List[np.ndarray] = field(default_factory=list)
    image_loss: List[float] = field(default_factory=list)
    text_loss: List[float] = field(default_factory=list)
    img_acc: List[float] = field(default_factory=list)
    text_acc: List[float] = field(default_factory=list)
    lr: List[float] = field(default_factory=list)
    evals: List[Tuple[int, float, float]] = field(default_factory=list)  # (iter, val_loss, val_acc)
    weights: List[Dict[str, torch.Tensor]] = field(default_factory=list)


def train(st: HeadState, img_bank, txt_bank, val_bank, *, batch_size: int, optim: str = "adamw",
          lr: float = 1e-3, weight_decay: float = 0.0, sched: str = "cosine", warmup_iter: int = 50,
          sched_max_iter: Optional[int] = None, warmup_type: Optional[str] = "linear",
          warmup_lr: Optional[float] = 1e-5, max_iters: int = 1000, alpha: float = 1.0,
          eval_freq: int = 100, patience: int = 5, num_workers: int = 0,
          record_weights_every: int = 0, use_autograd: bool = False,
          per_sample_collate: bool = False) -> Tuple[dict, TrainTrace]:
    """Restatement of finetune.train (finetune.py:157-288) over feature banks.

    ``img_bank`` / ``txt_bank`` / ``val_bank`` are ``(features, labels)`` tuples (None to drop
    a modality, as ``--modality image|text`` does at finetune.py:373-380).
    ``per_sample_collate`` fetches rows one by one and stacks them, as DataLoader's
    default_collate over a map-style dataset does (cost model only; identical values)."""
    assert img_bank is not None or txt_bank is not None
    sched_max_iter = max_iters if sched_max_iter is None else sched_max_iter
    opt = OracleOptimizer(st.param_dict(), optim, lr, weight_decay)
    trace = TrainTrace()
    il = OracleLoader(img_bank[0].shape[0], batch_size, num_workers=num_workers) if img_bank is not None else None
    tl = OracleLoader(txt_bank[0].shape[0], batch_size, num_workers=num_workers) if txt_bank is not None else None
    if il is not None:
        il.iter()
    if tl is not None:
        tl.iter()
    step_fn = uml_step_grads_autograd if use_autograd else uml_step_grads

    def rows(bank, idx):
        if per_sample_collate:
            return (torch.stack([bank[0][int(i)] for i in idx]),
                    torch.stack([bank[1][int(i)] for i in idx]))
        return bank[0][idx], bank[1][idx]

    out = {"iter": None, "val_acc": None, "val_loss": None, "model": None}
    no_improve = 0
    for i in range(max_iters):
        xi = yi = xt = yt = None
        if il is not None:
            idx = fetch_next_indices(il)
            trace.img_idx.append(idx.numpy().copy())
            xi, yi = rows(img_bank, idx)
        if tl is not None:
            idx = fetch_next_indices(tl)
            trace.txt_idx.append(idx.numpy().copy())
            xt, yt = rows(txt_bank, idx)
        stats, grads = step_fn(st, xi, yi, xt, yt, alpha)
        cur_lr = lr_at(i, lr, sched, warmup_iter, sched_max_iter, warmup_type, warmup_lr)
        opt.step(grads, cur_lr)
        trace.lr.append(cur_lr)
        for k in ("image_loss", "text_loss", "img_acc", "text_acc"):
            getattr(trace, k).append(stats[k])
        if record_weights_every and i % record_weights_every == 0:
            trace.weights.append({k: v.clone() for k, v in st.param_dict().items()})
        if i % eval_freq == 0:
            snap = {k: v.clone() for k, v in st.param_dict().items()}
            vloss, vacc = validate(st, val_bank[0], val_bank[1], batch_size)
            trace.evals.append((i, vloss, vacc))
            if out["val_acc"] is None or vacc > out["val_acc"]:
                out.update(iter=i, val_acc=vacc, val_loss=vloss, model=snap)
                no_improve = 0
            else:
                no_improve += 1
            if no_improve >= patience:
                break
    for k, v in out["model"].items():
        st.param_dict()[k].copy_(v)
    return out, trace


# --------------------------------------------------------------------------------------
# a-14 Gaussian linear analogue
# --------------------------------------------------------------------------------------


def gaussian_generate(seed: int, num_samples: int, dim_c: int, dim_x: int, dim_y: int, dim_obs: int,
                      noise_std: float, attenuate_x: bool, attenuation: float,
                      latent: str = "gaussian") -> Dict[str, torch.Tensor]:
    """data.py:29-61 - draw order: shared latent, private x/y latents, noises, then the four
    mixing matrices; X sees an attenuated copy of the shared latent (first 10% of dims kept)."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    if latent == "gaussian":
        tc = torch.randn(num_samples, dim_c)
        tc = tc - tc.mean(0)
    elif latent == "laplace":
        lap = torch.distributions.Laplace(torch.tensor([0.0]), torch.tensor([1.0]))
        tc = lap.sample((num_samples, dim_c)).squeeze(-1)
    else:
        raise ValueError(latent)
    tx = torch.randn(num_samples, dim_x)
    ty = torch.randn(num_samples, dim_y)
    nx = torch.randn(num_samples, dim_obs) * noise_std
    ny = torch.randn(num_samples, dim_obs) * noise_std
    a_c, a_x = torch.randn(dim_obs, dim_c), torch.randn(dim_obs, dim_x)
    b_c, b_y = torch.randn(dim_obs, dim_c), torch.randn(dim_obs, dim_y)
    if attenuate_x:
        att = torch.full((dim_c,), attenuation)
        att[: int(dim_c * 0.1)] = 1.0
        tcx = tc * att
    else:
        tcx = tc
    return {"x": tcx @ a_c.T + tx @ a_x.T + nx, "y": tc @ b_c.T + ty @ b_y.T + ny}


GAUSS_LAYERS = ("in_head_x", "in_head_y", "shared_encoder.0", "shared_encoder.2",
                "shared_decoder.0", "shared_decoder.2", "out_head_x", "out_head_y")


def gaussian_init(dim_obs: int, dim_common: int, dim_latent: int) -> Dict[str, torch.Tensor]:
    """nn.Linear default init in SharedAutoencoder's construction order (model.py:9-25):
    weight ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (kaiming_uniform with a=sqrt(5)), then bias ~ same bound."""
    dims = {"in_head_x": (dim_obs, dim_common), "in_head_y": (dim_obs, dim_common),
            "shared_encoder.0": (dim_common, dim_latent), "shared_encoder.2": (dim_latent, dim_latent),
            "shared_decoder.0": (dim_latent, dim_latent), "shared_decoder.2": (dim_latent, dim_common),
            "out_head_x": (dim_common, dim_obs), "out_head_y": (dim_common, dim_obs)}
    p = {}
    for name in GAUSS_LAYERS:
        fin, fout = dims[name]
        bound = 1.0 / math.sqrt(fin)
        p[name + ".weight"] = torch.empty(fout, fin).uniform_(-bound, bound)
        p[name + ".bias"] = torch.empty(fout).uniform_(-bound, bound)
    return p


def _gauss_branch(p, v, m):
    """recon = out_head_m(dec(enc(in_head_m(v)))); returns loss and grads for the branch."""
    lin = lambda n, a: a @ p[n + ".weight"].t() + p[n + ".bias"]
    a0 = lin(f"in_head_{m}", v)
    h1 = lin("shared_encoder.0", a0); r1 = torch.relu(h1)
    lat = lin("shared_encoder.2", r1)
    h2 = lin("shared_decoder.0", lat); r2 = torch.relu(h2)
    a3 = lin("shared_decoder.2", r2)
    rec = lin(f"out_head_{m}", a3)
    diff = rec - v
    loss = (diff * diff).mean()
    g = {}
    d = 2.0 * diff / diff.numel()

    def back(name, inp, dout):
        g[name + ".weight"] = dout.t() @ inp
        g[name + ".bias"] = dout.sum(0)
        return dout @ p[name + ".weight"]

    d = back(f"out_head_{m}", a3, d)
    d = back("shared_decoder.2", r2, d) * (h2 > 0)
    d = back("shared_decoder.0", lat, d)
    d = back("shared_encoder.2", r1, d) * (h1 > 0)
    d = back("shared_encoder.0", a0, d)
    back(f"in_head_{m}", v, d)
    return loss, g, rec, lat


def gaussian_step_grads(p, x, y, mode: str, alpha_x: float, alpha_y: float):
    """main.py:47-59: xy -> alpha_x*MSE(x)+alpha_y*MSE(y); x -> MSE(x) only (y-heads get no grad)."""
    lx, gx, _, _ = _gauss_branch(p, x, "x")
    ly, gy, _, _ = _gauss_branch(p, y, "y")
    grads: Dict[str, Optional[torch.Tensor]] = {k: None for k in p}
    if mode == "xy":
        for k, v in gx.items():
            grads[k] = alpha_x * v
        for k, v in gy.items():
            grads[k] = alpha_y * v if grads[k] is None else grads[k] + alpha_y * v
    elif mode == "x":
        grads.update(gx)
    else:
        raise ValueError(mode)
    return float(lx), float(ly), grads


def gaussian_train(p, data_x, data_y, *, num_steps: int, batch_size: int = 512, lr: float = 1e-3,
                   mode: str = "xy", alpha_x: float = 1.0, alpha_y: float = 1.0, loader_seed: int = 42):
    """train_model_steps (main.py:31-59) + loader of main.py:139-143 (Generator(42), drop_last=True)
    + UnpairedDataset index wrap (dataset.py:14-18).  Adam(lr) with torch defaults."""
    g = torch.Generator()
    g.manual_seed(loader_seed)
    n = max(len(data_x), len(data_y))
    loader = OracleLoader(n, batch_size, drop_last=True, generator=g)
    loader.iter()
    opt = OracleOptimizer(p, "adam", lr, 0.0)
    hist = []
    for _ in range(num_steps):
        idx = fetch_next_indices(loader)
        x, y = data_x[idx % len(data_x)], data_y[idx % len(data_y)]
        lx, ly, grads = gaussian_step_grads(p, x, y, mode, alpha_x, alpha_y)
        opt.step(grads)
        hist.append((lx, ly, idx.numpy().copy()))
    return hist
