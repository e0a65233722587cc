"""Sweep-level batching: K heads of one hyper-parameter sweep advanced in lock step.

The reference's ``sweep`` (finetune.py:406-448) trains every lr x weight-decay combination of ``HYPER_DICT``
(engine/optimizer/default.py) one after the other over the same banks; at its batch size of 32 a step cannot fill one
SM.  ``HeadGroup`` keeps the K heads' weights and optimizer state in three ``[K, C*D]`` slabs and runs one step of all
of them with two launches - logits + softmax / CE and dW + update, both on the tensor cores (``uml_sweep_run``, csrc/sweep.cu).  Each head keeps its own sampler stream, learning-rate
schedule, weight decay, alpha and early-stopping state; a stopped head is masked out of the launches.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from .. import _lib
from .._lib import SWEEP_MAX_HEADS, SweepArgs, check
from ..ops import UPDATE_KINDS

MAX_HEADS = SWEEP_MAX_HEADS


def group_blockers(models, optimizers, image_loaders, text_loaders) -> List[str]:
    """Why these runs cannot share a HeadGroup (empty list: they can)."""
    why = []
    m0, o0 = models[0], optimizers[0]
    if len(models) > MAX_HEADS:
        why.append(f"more than {MAX_HEADS} heads")
    for m in models:
        if m.img_proj is not None:
            why.append("adapter (img_proj) variants are not batched")
            break
    if any(getattr(m, "learnable_temp", False) for m in models):
        why.append("learnable temperatures are not batched")
    if any(m.num_classes != m0.num_classes or m.shared_dim != m0.shared_dim for m in models):
        why.append("heads differ in shape")
    if not why and any(tuple(float(s) for s in m.scales()) != tuple(float(s) for s in m0.scales()) for m in models):
        why.append("heads differ in logit scale")
    if any(o.name != o0.name or o.defaults["betas"] != o0.defaults["betas"] or o.defaults["eps"] != o0.defaults["eps"]
           or o.defaults["momentum"] != o0.defaults["momentum"] for o in optimizers):
        why.append("optimizers differ in kind / betas / eps / momentum")
    for loaders in (image_loaders, text_loaders):
        live = [l for l in loaders if l is not None]
        if live and len(live) != len(loaders):
            why.append("runs differ in modality")
        elif live:
            l0 = live[0]
            if any(l.bank is not l0.bank or l.batch_size != l0.batch_size or l.drop_last != l0.drop_last
                   or not l.shuffle or l.generator is not None or l.shard_of is not None or l.upload != "epoch"
                   for l in live):
                why.append("loaders must share bank, batch size and protocol (shuffled, upload='epoch', no generator)")
    return why


class HeadGroup:
    def __init__(self, models: Sequence, optimizers: Sequence, image_bank, text_bank, max_img_rows: int,
                 max_txt_rows: int, device, log_slots: int = 128):
        self.K = K = len(models)
        if not 1 <= K <= MAX_HEADS:
            raise ValueError(f"a HeadGroup holds 1..{MAX_HEADS} heads")
        self.models, self.opts = list(models), list(optimizers)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HeadGroup: heads must live on a CUDA device (no CPU path)")
        m0, o0 = models[0], optimizers[0]
        self.C, self.D = int(m0.num_classes), int(m0.shared_dim)
        self.kind = UPDATE_KINDS[o0.name]
        if self.kind == 0:
            raise ValueError("HeadGroup needs an optimizer")
        self.banks = (image_bank, text_bank)
        for b in self.banks:
            if b is not None and (b.features.dtype != torch.float32 or b.dim != self.D or not b.features.is_contiguous()):
                raise ValueError("HeadGroup: banks must be contiguous fp32 [N, shared_dim]")
        self.max_rows = int(max_img_rows) + int(max_txt_rows)
        n = self.C * self.D
        self.stride = (n + 3) // 4 * 4
        self.W = torch.empty((K, self.stride), device=self.device)
        self.m = torch.zeros((K, self.stride), device=self.device)
        self.v = torch.zeros((K, self.stride), device=self.device) if o0.name != "sgd" else None
        for k, (model, opt) in enumerate(zip(models, optimizers)):
            p = model.head.weight
            if p.device.type != "cuda":
                raise RuntimeError("HeadGroup: heads must live on a CUDA device (no CPU path)")
            view = self.W[k, :n].view(self.C, self.D)
            view.copy_(p.data)
            p.data = view  # the module keeps working (state_dict, validate) on its slice of the slab
            st = opt.slot(p)
            if st["step"] != 0:
                raise ValueError("HeadGroup: optimizers must be fresh")
            st["m"] = self.m[k, :n].view(self.C, self.D)
            if self.v is not None:
                st["v"] = self.v[k, :n].view(self.C, self.D)
        self.ldg = (self.C + 3) // 4 * 4
        self.G = torch.empty((K, self.max_rows, self.ldg), device=self.device)
        self.row_loss = torch.empty((K, self.max_rows), device=self.device)
        self.row_correct = torch.empty((K, self.max_rows), device=self.device, dtype=torch.int32)
        self.log_slots = int(log_slots)
        self.stats_log = torch.zeros((self.log_slots, K, 2, 4), device=self.device)
        self.launches = 0
        s_i, s_t = (float(s) for s in m0.scales())
        a = self.args = SweepArgs()
        a.n_heads, a.dim, a.n_classes, a.kind = K, self.D, self.C, self.kind
        for s, b in enumerate(self.banks):
            if b is not None:
                a.bank[s], a.bank_ld[s], a.labels[s] = b.features.data_ptr(), b.features.stride(0), b.labels.data_ptr()
                a.perm_len[s] = len(b)
        a.scale[0], a.scale[1] = s_i, s_t
        a.W, a.m, a.v = self.W.data_ptr(), self.m.data_ptr(), (self.v.data_ptr() if self.v is not None else None)
        a.head_stride = self.stride
        a.G, a.ldg, a.max_rows = self.G.data_ptr(), self.ldg, self.max_rows
        a.row_loss, a.row_correct = self.row_loss.data_ptr(), self.row_correct.data_ptr()
        a.beta1, a.beta2 = o0.defaults["betas"]
        a.eps, a.momentum = o0.defaults["eps"], o0.defaults["momentum"]
        for k, (model, opt) in enumerate(zip(models, optimizers)):
            a.weight_decay[k] = opt.group_of(model.head.weight)["weight_decay"]

    def run(self, perms_img: Optional[Sequence[torch.Tensor]], perms_txt: Optional[Sequence[torch.Tensor]], pos_img: int,
            pos_txt: int, rows: Sequence[Sequence[int]], lrs: Sequence[Sequence[float]], alphas: Sequence[float],
            active: Sequence[bool], slot0: int):
        """Enqueue ``len(rows)`` consecutive steps of every active head.  ``perms_*[k]``: head k's epoch permutation on
        the device; ``rows[i] = (image rows, text rows)`` of step i; ``lrs[i][k]``; stats go to log slots slot0...;
        all on the current stream, no synchronisation."""
        n, K, a = len(rows), self.K, self.args
        s0 = slot0 % self.log_slots
        if n == 0 or not any(active):
            return
        if s0 + n > self.log_slots:
            raise ValueError("HeadGroup.run: the steps of one call must fit the stats log without wrapping")
        for s, perms in enumerate((perms_img, perms_txt)):
            for k in range(K):
                if perms is None:
                    a.perm[s][k] = None
                    continue
                t = perms[k]
                if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous() or t.numel() < a.perm_len[s]:
                    raise ValueError("HeadGroup.run: permutations must be contiguous CUDA int64 tensors covering the bank")
                a.perm[s][k] = t.data_ptr()
        a.pos[0], a.pos[1] = int(pos_img), int(pos_txt)
        step = None
        for k in range(K):
            a.alpha[k] = float(alphas[k])
            a.active[k] = 1 if active[k] else 0
            if active[k]:
                st = self.opts[k].slot(self.models[k].head.weight)
                if step is None:
                    step = st["step"]
                elif st["step"] != step:
                    raise ValueError("HeadGroup.run: active heads must have taken the same number of steps")
        a.step = step + 1
        a.stats = self.stats_log[s0].data_ptr()
        rows_c = (C.c_int64 * (2 * n))(*[int(x) for r in rows for x in r])
        lr_c = (C.c_float * (n * K))(*[float(x) for l in lrs for x in l])
        l0 = _lib.load().uml_sweep_launch_count()
        check(_lib.load().uml_sweep_run(C.byref(a), n, rows_c, lr_c, torch.cuda.current_stream().cuda_stream))
        launched = (_lib.load().uml_sweep_launch_count() - l0) & 0x7FFFFFFF  # two to four kernels per step (csrc/sweep.cu)
        self.launches += launched
        _lib.LAUNCH_COUNT[0] += launched
        for k in range(K):
            if active[k]:
                self.opts[k].slot(self.models[k].head.weight)["step"] += n

    KERNELS = ("sweep_logits", "sweep_softmax_ce", "sweep_dw_update", "sweep_stats")

    def time_last_step(self, enable: bool = True):
        """Bracket the four launch sites of the LAST step of every following ``run`` with CUDA events (a site whose work was
        fused into its neighbour spans nothing)."""
        self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)] if enable else []
        for i in range(8):
            if enable:
                self._ev[i].record()  # forces creation of the underlying cudaEvent_t
                self.args.ev[i] = self._ev[i].cuda_event
            else:
                self.args.ev[i] = None

    def kernel_times_ms(self):
        """Device time of each launch of the last timed step (call after a synchronize)."""
        ev = getattr(self, "_ev", [])
        return {name: ev[2 * i].elapsed_time(ev[2 * i + 1]) for i, name in enumerate(self.KERNELS)} if ev else {}

    def read_log(self, slots: Sequence[int], has_img: bool, has_txt: bool):
        """Per-step stats of the given log slots as ``{name: [len(slots)][K] nested lists}`` for image_loss, text_loss,
        img_acc, text_acc (zeros for an absent modality); one synchronising D2H copy of the log."""
        raw = self.stats_log.cpu().numpy()
        ints = raw.view("int32")
        idx = [s % self.log_slots for s in slots]
        zeros = [[0.0] * self.K for _ in idx]
        out = {"image_loss": zeros, "text_loss": zeros, "img_acc": zeros, "text_acc": zeros}
        for name_l, name_a, s, present in (("image_loss", "img_acc", 0, has_img), ("text_loss", "text_acc", 1, has_txt)):
            if present:
                out[name_l] = raw[idx, :, s, 0].tolist()
                out[name_a] = (ints[idx, :, s, 2] / ints[idx, :, s, 3].clip(min=1)).tolist()
        return out
