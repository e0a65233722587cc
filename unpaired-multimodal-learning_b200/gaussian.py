"""The linear analogue of the UML step: ``Gaussian_experiment`` of the reference.

Mirrors ``Gaussian_experiment/{model.py, dataset.py, data.py, main.py}``: a ``SharedAutoencoder`` whose
encoder/decoder are shared by two unpaired modalities, trained with ``alpha_x MSE(x) + alpha_y MSE(y)``
("xy") or ``MSE(x)`` ("x") and Adam.  The whole step (gather by sampler index, six layers forward, MSE,
six layers backward, Adam) runs in two hand-written kernels (``csrc/gauss.cu``); the per-step losses land
in a device log that is read once at the end.  Same names and argument meaning as the reference so its
``main.py`` reads the same against this module; the per-step CKA / mutual-kNN probes of the reference's loop
(``main.py:67-84``) are diagnostics outside the training arithmetic and are not run.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check
from .engine.datasets.utils import BankLoader

LAYERS = ("in_head_x", "in_head_y", "shared_encoder.0", "shared_encoder.2",
          "shared_decoder.0", "shared_decoder.2", "out_head_x", "out_head_y")


def generate_data(config: Dict) -> Dict[str, torch.Tensor]:
    """``data.generate_data`` (data.py:29-61): shared latent (gaussian, centred, or laplace), private latents,
    observation noise, four random mixing matrices; X sees the shared latent attenuated except for its first
    10 % of dimensions.  Host-side torch draws in the reference's order, so the tensors are bit-identical."""
    seed, n = config["seed"], config["num_samples"]
    dim_c, dim_x, dim_y, dim_obs = config["dim_c"], config["dim_x"], config["dim_y"], config["dim_obs"]
    torch.manual_seed(seed)
    np.random.seed(seed)
    kind = config.get("shared_latent_distribution_type", "gaussian")
    if kind == "gaussian":
        tc = torch.randn(n, dim_c)
        tc = tc - tc.mean(0)
    elif kind == "laplace":
        tc = torch.distributions.Laplace(torch.tensor([0.0]), torch.tensor([1.0])).sample((n, dim_c)).squeeze(-1)
    else:
        raise ValueError(f"unknown shared_latent_distribution_type {kind!r}")
    tx, ty = torch.randn(n, dim_x), torch.randn(n, dim_y)
    nx, ny = torch.randn(n, dim_obs) * config["noise_std"], torch.randn(n, dim_obs) * config["noise_std"]
    a_c, a_x = torch.randn(dim_obs, dim_c), torch.randn(dim_obs, dim_x)
    b_c, b_y = torch.randn(dim_obs, dim_c), torch.randn(dim_obs, dim_y)
    if config.get("attenuate_x", False):
        att = torch.full((dim_c,), float(config["attenuation"]))
        att[: int(dim_c * 0.1)] = 1.0
        tcx = tc * att
    else:
        tcx = tc
    return {"x": tcx @ a_c.T + tx @ a_x.T + nx, "y": tc @ b_c.T + ty @ b_y.T + ny}


class UnpairedDataset:
    """``dataset.UnpairedDataset``: length = max of the two; item ``i`` pairs ``x[i % len_x]`` with ``y[i % len_y]``
    (the wrap is applied inside the kernel).  The rows live in HBM."""

    def __init__(self, data_x: torch.Tensor, data_y: torch.Tensor, device="cuda"):
        self.data_x = data_x.detach().to(device=device, dtype=torch.float32).contiguous()
        self.data_y = data_y.detach().to(device=device, dtype=torch.float32).contiguous()
        self.len_x, self.len_y = len(data_x), len(data_y)
        self.length = max(self.len_x, self.len_y)
        self.device = self.data_x.device

    def __len__(self):
        return self.length


def unpaired_loader(dataset: UnpairedDataset, batch_size: int, generator: Optional[torch.Generator] = None) -> BankLoader:
    """``DataLoader(dataset, batch_size, shuffle=True, drop_last=True, generator=g)`` (main.py:141-143) as an index
    loader with the same RNG protocol (explicit generator: one permutation per epoch plus the discarded trailing one)."""
    return BankLoader(dataset, batch_size, shuffle=True, drop_last=True, generator=generator)


class SharedAutoencoder:
    """``model.SharedAutoencoder(dim_obs, dim_common, dim_latent)``: parameters live in ONE flat CUDA buffer in the
    reference's construction order; ``state_dict`` uses the reference's keys.  Initialisation draws the same
    ``nn.Linear`` defaults from the global CPU generator in the same order as the reference constructor."""

    def __init__(self, dim_obs: int, dim_common: int, dim_latent: int, device="cuda"):
        self.dim_obs, self.dim_common, self.dim_latent = dim_obs, dim_common, dim_latent
        shapes = {"in_head_x": (dim_common, dim_obs), "in_head_y": (dim_common, dim_obs),
                  "shared_encoder.0": (dim_latent, dim_common), "shared_encoder.2": (dim_latent, dim_latent),
                  "shared_decoder.0": (dim_latent, dim_latent), "shared_decoder.2": (dim_common, dim_latent),
                  "out_head_x": (dim_obs, dim_common), "out_head_y": (dim_obs, dim_common)}
        # construction order of the reference: in heads, encoder (0, 2), decoder (0, 2), out heads
        host = {}
        for name in LAYERS:
            lin = torch.nn.Linear(shapes[name][1], shapes[name][0])
            host[name + ".weight"], host[name + ".bias"] = lin.weight.detach(), lin.bias.detach()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SharedAutoencoder: parameters must live on a CUDA device (no CPU path)")
        n = int(_lib.load().uml_gauss_param_count(dim_obs, dim_common, dim_latent))
        self.flat = torch.empty(n, device=self.device)
        self._views, off = {}, 0
        for name in LAYERS:
            for suffix in (".weight", ".bias"):
                t = host[name + suffix]
                self._views[name + suffix] = self.flat[off:off + t.numel()].view(t.shape)
                self._views[name + suffix].copy_(t)
                off += t.numel()
        assert off == n

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self._views.items()}

    def load_state_dict(self, sd):
        for k, v in self._views.items():
            v.copy_(sd[k].to(self.device))

    def parameters(self):
        return list(self._views.values())

    def train(self):
        return self

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("SharedAutoencoder has no CPU path")
        return self


class Adam:
    """``optim.Adam(model.parameters(), lr)`` with torch's defaults; state is two flat buffers next to the model's."""

    def __init__(self, model: SharedAutoencoder, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.model, self.lr, self.betas, self.eps = model, float(lr), betas, float(eps)
        self.m, self.v = torch.zeros_like(model.flat), torch.zeros_like(model.flat)
        self.step_count = 0


def validate(model: SharedAutoencoder, val_data_x: torch.Tensor, val_data_y: torch.Tensor):
    """(val_loss_x, val_loss_y) = MSE of both reconstructions over the validation rows (main.py:68-72)."""
    x = val_data_x.to(device=model.device, dtype=torch.float32).contiguous()
    y = val_data_y.to(device=model.device, dtype=torch.float32).contiguous()
    if x.shape != y.shape:
        raise ValueError("validate: x and y validation sets must have the same shape")
    n = x.shape[0]
    ws = torch.empty(2 * ((n + 15) // 16), device=model.device)
    out = torch.empty(2, device=model.device)
    check(_lib.load().uml_gauss_eval(model.flat.data_ptr(), model.dim_obs, model.dim_common, model.dim_latent, x.data_ptr(),
                                     y.data_ptr(), n, ws.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    lx, ly = out.tolist()
    return lx, ly


def get_embeddings(model: SharedAutoencoder, x=None, y=None):
    """``model.get_embeddings`` of the reference (model.py:51-60): latent = shared_encoder(in_head(rows)) per modality,
    ``[n, dim_latent]`` device tensors (None for an absent modality)."""
    def prep(t):
        return None if t is None else t.to(device=model.device, dtype=torch.float32).contiguous()
    x, y = prep(x), prep(y)
    if x is None and y is None:
        return None, None
    n = (x if x is not None else y).shape[0]
    if x is not None and y is not None and x.shape[0] != y.shape[0]:
        raise ValueError("get_embeddings: x and y must have the same number of rows")
    ex = torch.empty((n, model.dim_latent), device=model.device) if x is not None else None
    ey = torch.empty((n, model.dim_latent), device=model.device) if y is not None else None
    check(_lib.load().uml_gauss_embed(model.flat.data_ptr(), model.dim_obs, model.dim_common, model.dim_latent,
                                      x.data_ptr() if x is not None else None, y.data_ptr() if y is not None else None, n,
                                      ex.data_ptr() if ex is not None else None, ey.data_ptr() if ey is not None else None,
                                      torch.cuda.current_stream().cuda_stream))
    return ex, ey


def alignment(model: SharedAutoencoder, val_data_x, val_data_y, topk: int = 10):
    """(cka, mknn) of the two modalities' validation embeddings - the probes main.py:21-29,78-83 logs as val/cka and
    val/mknn - computed on the device (``csrc/probes.cu``); one read-back of two floats."""
    from . import ops
    ex, ey = get_embeddings(model, val_data_x, val_data_y)
    both = torch.cat([ops.cka_linear(ex, ey), ops.mutual_knn(ex, ey, topk)]).tolist()
    return both[0], both[1]


def train_model_steps(model: SharedAutoencoder, data_loader: BankLoader, optimizer: Adam, num_steps: int,
                      val_data_x=None, val_data_y=None, device="cuda", args=None, eval_every: int = 0, trace=None):
    """``main.train_model_steps`` (main.py:31-86).  ``args`` carries ``mode`` ('xy' | 'x'), ``alpha_x``, ``alpha_y``.
    Returns ``{'loss_x': [...], 'loss_y': [...], 'loss': [...], 'val': [(step, val_x, val_y), ...], 'align': [(step, cka,
    mknn), ...]}``; the
    training losses are read back from the device log once, after the last step (or at validation points)."""
    mode = getattr(args, "mode", "xy")
    alpha_x, alpha_y = float(getattr(args, "alpha_x", 1.0)), float(getattr(args, "alpha_y", 1.0))
    if mode not in ("xy", "x"):
        raise ValueError("mode must be 'xy' or 'x'")
    ds: UnpairedDataset = data_loader.bank
    lib = _lib.load()
    B = data_loader.batch_size
    ws = torch.empty(int(lib.uml_gauss_workspace_floats(model.dim_obs, model.dim_common, model.dim_latent, B)), device=model.device)
    log = torch.zeros((num_steps, 2), device=model.device)
    stream = torch.cuda.current_stream().cuda_stream
    import ctypes as C
    it = iter(data_loader)
    out = {"loss_x": [], "loss_y": [], "loss": [], "val": [], "align": []}
    chunk_max = 32  # steps enqueued per library call (a Python round trip per step costs more than the step's kernels)
    step = 0
    while step < num_steps:
        n = min(chunk_max, num_steps - step)
        if eval_every and val_data_x is not None:
            n = min(n, eval_every - step % eval_every)  # a chunk ends at the next validation point
        batches = []
        for _ in range(n):
            try:
                batch = next(it)
            except StopIteration:
                it = iter(data_loader)
                batch = next(it)
            if trace is not None:
                trace.setdefault("idx", []).append(batch.host_idx.clone())
            batches.append(batch)
        ptrs = (C.c_void_p * n)(*[b.idx.data_ptr() for b in batches])
        check(lib.uml_gauss_run(model.flat.data_ptr(), optimizer.m.data_ptr(), optimizer.v.data_ptr(), model.dim_obs,
                                model.dim_common, model.dim_latent, ds.data_x.data_ptr(), ds.len_x, ds.data_y.data_ptr(),
                                ds.len_y, ptrs, n, B, int(mode == "xy"), alpha_x, alpha_y, optimizer.lr, optimizer.betas[0],
                                optimizer.betas[1], optimizer.eps, optimizer.step_count + 1, ws.data_ptr(),
                                log[step].data_ptr(), stream))
        optimizer.step_count += n
        step += n
        if eval_every and val_data_x is not None and step % eval_every == 0:
            out["val"].append((step - 1,) + validate(model, val_data_x, val_data_y))
            # val/cka and val/mknn of the validation embeddings, as the reference logs them at every evaluation point
            out["align"].append((step - 1,) + alignment(model, val_data_x, val_data_y))
    host = log.cpu()
    out["loss_x"], out["loss_y"] = host[:, 0].tolist(), host[:, 1].tolist()
    out["loss"] = [(alpha_x * a + alpha_y * b) if mode == "xy" else a for a, b in zip(out["loss_x"], out["loss_y"])]
    return out
