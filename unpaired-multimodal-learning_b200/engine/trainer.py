"""The fused UML training step (L3 hot loop body) on device.

One call to ``StepEngine.step`` is one iteration of the reference loop body
(vision_language/finetune.py:163-195): fetch an image batch and an unpaired text batch, push both
through the shared head, ``loss = CE(img) + alpha * CE(txt)``, backward to the shared weights,
optimizer update.  No autograd graph, no ``.item()``: per-step losses / hit counts land in a device
log that the caller reads at evaluation cadence.

Two arithmetic paths (chosen per step from the row count unless forced):
  fp32  SIMT kernels, exact like the reference; rows are gathered inside the GEMMs by index and the
        optimizer update runs in the dW epilogue: 3 launches for the linear head.
  bf16  tcgen05 kernels: TMA gather+cast -> fused forward/CE/G -> split-K dW -> optimizer update that
        also sums the split-K partials and refreshes the bf16 weight shadow.
Data-parallel runs (``dist_group``): every rank takes its slice of the global batch, gradients are
summed with one NCCL all-reduce per step and every rank applies the same update.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from .datasets.utils import IndexBatch

BF16_MIN_ROWS = 1024  # below this the step is launch-bound and the exact fp32 path is used


class StepEngine:
    def __init__(self, model, optimizer, device, max_img_rows: int, max_txt_rows: int, log_slots: int = 128,
                 precision: str = "auto", dist_group=None, world_size: int = 1):
        if precision not in ("auto", "fp32", "bf16"):
            raise ValueError("precision must be auto, fp32 or bf16")
        self.model, self.opt, self.device = model, optimizer, torch.device(device)
        self.precision = precision
        self.C, self.D, self.Dv = model.num_classes, model.shared_dim, model.img_indim
        self.adapter = model.img_proj is not None
        self.max_img, self.max_txt = int(max_img_rows), int(max_txt_rows)
        self.max_rows = self.max_img + self.max_txt
        self.dist_group, self.world = dist_group, int(world_size)
        self.W = model.head.weight
        if self.W.device.type != "cuda":
            raise RuntimeError("StepEngine: the model must live on a CUDA device (no CPU path)")
        self.learnable = bool(getattr(model, "learnable_temp", False))
        self.ws32: Optional[ops.HeadWorkspace] = None
        self.ws16: Optional[ops.HeadWorkspace] = None
        self.log_slots = int(log_slots)
        self.stats_log = torch.zeros((self.log_slots, 2, 4), device=self.device, dtype=torch.float32)
        self.slot_modalities = [None] * self.log_slots
        self.dW = None       # fp32 gradient buffer (data-parallel / unfused paths)
        self.Z = self.dZ = None
        self.X16 = self.W16 = self.labels32 = self.partials = None
        self._w16_valid = False
        self.host_log = None
        self.profile = None  # set to {} to collect (start, end) CUDA events per kernel name

    # ------------------------------------------------------------------------------------ helpers
    def _timed(self, name, fn, *a, **k):
        if self.profile is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        self.profile.setdefault(name, []).append((e0, e1))
        return r

    def kernel_times_ms(self):
        """Mean device time per launch of every profiled kernel (call after a synchronize)."""
        return {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in (self.profile or {}).items()}

    def _scales(self):
        si, st = self.model.scales()
        if self.learnable:  # device scalars, read by the kernels; never synchronise to fetch them
            return 0.0, 0.0, si.data, st.data
        return float(si), float(st), None, None

    def _use_bf16(self, rows: int) -> bool:
        if self.precision == "fp32" or self.adapter:
            return False
        ok = self.D % 8 == 0 and self.C <= 2048
        if self.precision == "bf16":
            if not ok:
                raise RuntimeError("bf16 path needs dim % 8 == 0 and <= 2048 classes")
            return True
        return ok and rows >= BF16_MIN_ROWS

    @staticmethod
    def _view(b: IndexBatch):
        """(rows, labels, idx) for a batch; sequential batches become dense views."""
        if b.idx is not None:
            return b.bank.features, b.bank.labels, b.idx
        return b.bank.features[b.start:b.start + b.n], b.bank.labels[b.start:b.start + b.n], None

    def invalidate_shadow(self):
        """Call after the weights were changed outside the engine (e.g. load_state_dict)."""
        self._w16_valid = False

    # ------------------------------------------------------------------------------------ step
    def step(self, img: Optional[IndexBatch], txt: Optional[IndexBatch], alpha: float, slot: int,
             global_img_rows: Optional[int] = None, global_txt_rows: Optional[int] = None):
        if img is None and txt is None:
            raise ValueError("at least one modality per step")
        n_i, n_t = (img.n if img else 0), (txt.n if txt else 0)
        slot %= self.log_slots
        self.slot_modalities[slot] = (img is not None, txt is not None)
        # data parallel: the loss is a mean over the GLOBAL batch, so rescale the local 1/n
        wi = 1.0 if global_img_rows is None or n_i == 0 else n_i / float(global_img_rows)
        wt = alpha if global_txt_rows is None or n_t == 0 else alpha * n_t / float(global_txt_rows)
        if self._use_bf16(n_i + n_t):
            self._step_bf16(img, txt, n_i, n_t, wi, wt, slot)
        else:
            self._step_fp32(img, txt, n_i, n_t, wi, wt, slot)

    # ------------------------------------------------------------------------------------ fp32
    def _step_fp32(self, img, txt, n_i, n_t, wi, wt, slot):
        if self.ws32 is None:
            self.ws32 = ops.HeadWorkspace(self.max_rows, self.C, self.device)
        ws, W = self.ws32, self.W.data
        s_i, s_t, sd_i, sd_t = self._scales()
        runs = []
        if img is not None and n_i:
            rows, labels, idx = self._view(img)
            if self.adapter:
                if self.Z is None:
                    self.Z = torch.empty((self.max_img, self.D), device=self.device)
                    self.dZ = torch.empty((self.max_img, self.D), device=self.device)
                Wp = self.model.img_proj.weight.data
                ops.gemm_nt(rows, Wp, self.Z, a_row_idx=idx, m=n_i)
                runs.append(ops.Run(self.Z, labels, None, n_i, s_i, wi, label_idx=idx, scale_dev=sd_i))
            else:
                runs.append(ops.Run(rows, labels, idx, n_i, s_i, wi, scale_dev=sd_i))
        if txt is not None and n_t:
            rows, labels, idx = self._view(txt)
            runs.append(ops.Run(rows, labels, idx, n_t, s_t, wt, scale_dev=sd_t))
        stats = self.stats_log[slot]
        self._timed("head_fwd_ce_f32", ops.head_fwd_ce_f32, runs, W, ws, stats=stats)
        if self.adapter and img is not None and n_i:
            # dZ = G_img W uses the head BEFORE its update; image rows are the first n_i rows of G
            ops.gemm_nn(ws.G, W, self.dZ, m=n_i, k=self.C)
        if self.learnable:
            k = 0
            if img is not None and n_i:
                self._apply_scalar(self.model.img_scale, stats[k, 1:2])
                k += 1
            if txt is not None and n_t:
                self._apply_scalar(self.model.txt_scale, stats[k, 1:2])
        if self.world > 1:
            if self.dW is None:
                self.dW = torch.empty_like(W)
            ops.head_bwd_dw_f32(runs, W, ws, dW=self.dW)
            torch.distributed.all_reduce(self.dW, group=self.dist_group)
            self.opt.apply(self.W, self.dW)
        else:
            self._timed("head_bwd_dw_f32", ops.head_bwd_dw_f32, runs, W, ws, update=self.opt.update_struct(self.W))
        if self.adapter and img is not None and n_i:
            rows, _, idx = self._view(img)
            Wp = self.model.img_proj.weight
            if self.world > 1:
                raise NotImplementedError("data-parallel adapter training is not wired yet")
            ops.gemm_tn(self.dZ, rows, None, k=n_i, m=self.D, n=self.Dv, b_row_idx=idx, P=Wp.data,
                        update=self.opt.update_struct(Wp), ldc=Wp.data.stride(0))
        self._w16_valid = False

    def _apply_scalar(self, param, grad_view):
        if self.world > 1:
            g = grad_view.clone()
            torch.distributed.all_reduce(g, group=self.dist_group)
            grad_view = g
        self.opt.apply(param, grad_view)

    # ------------------------------------------------------------------------------------ bf16
    def _alloc_bf16(self):
        dev = self.device
        self.ws16 = ops.HeadWorkspace(self.max_rows, self.C, dev, bf16=True)
        self.X16 = torch.empty((self.max_rows, self.D), device=dev, dtype=torch.bfloat16)
        self.W16 = torch.empty((self.C, self.D), device=dev, dtype=torch.bfloat16)
        self.labels32 = torch.empty(self.max_rows, device=dev, dtype=torch.int32)
        self.max_splits = max(1, ops.tc_dw_splits(self.max_rows, self.D, self.C))
        self.partials = torch.empty((self.max_splits, self.C, self.D), device=dev)

    def _step_bf16(self, img, txt, n_i, n_t, wi, wt, slot):
        if self.ws16 is None:
            self._alloc_bf16()
        ws, W, n = self.ws16, self.W.data, n_i + n_t
        if not self._w16_valid:
            ops.cast_bf16(W, self.W16)
            self._w16_valid = True
        s_i, s_t, sd_i, sd_t = self._scales()
        rows_l, scales, weights, sdevs = [], [], [], []
        off = 0
        for b, cnt, s, w, sd in ((img, n_i, s_i, wi, sd_i), (txt, n_t, s_t, wt, sd_t)):
            if b is None or cnt == 0:
                continue
            feats, labels, idx = self._view(b)
            if idx is None:
                idx_arg = None
                self._timed("cast_bf16", ops.cast_bf16, feats, self.X16[off:off + cnt])
            else:
                idx_arg = idx
                self._timed("gather_bf16", ops.gather_rows, feats, idx, out=self.X16[off:off + cnt])
            ops.gather_labels(labels, idx_arg, cnt, self.labels32[off:off + cnt])
            rows_l.append(cnt); scales.append(s); weights.append(w); sdevs.append(sd)
            off += cnt
        segs = ops.tc_segments(rows_l, scales, weights, sdevs if self.learnable else None)
        self._timed("head_fwd_ce_bf16", ops.head_fwd_ce_bf16, self.X16, self.W16, self.labels32, segs, ws, ws.row_loss,
                    row_correct=ws.row_correct, row_dscale=ws.row_dscale if self.learnable else None, n_rows=n)
        stats = self.stats_log[slot]
        ops.reduce_seg_stats(ws.row_loss, ws.row_correct, ws.row_dscale if self.learnable else None, rows_l, stats)
        if self.learnable:
            k = 0
            if n_i:
                self._apply_scalar(self.model.img_scale, stats[k, 1:2]); k += 1
            if n_t:
                self._apply_scalar(self.model.txt_scale, stats[k, 1:2])
        splits = min(self.max_splits, max(1, ops.tc_dw_splits(n, self.D, self.C)))
        self._timed("head_bwd_dw_bf16", ops.head_bwd_dw_bf16, ws.G, ws.ldg, self.X16, n, self.C, self.partials, splits)
        g, st = self.opt.group_of(self.W), self.opt.slot(self.W)
        if self.world > 1 or self.opt.name == "sgd":
            if self.dW is None:
                self.dW = torch.empty_like(W)
            ops.sum_partials(self.partials, splits, W.numel(), self.dW)
            if self.world > 1:
                torch.distributed.all_reduce(self.dW, group=self.dist_group)
            self.opt.apply(self.W, self.dW, shadow=self.W16)
        else:
            st["step"] += 1
            self._timed("adamw_step_partials", ops.adamw_step_partials, W, self.partials, splits, st["m"], st["v"],
                        lr=g["lr"], step=st["step"], weight_decay=g["weight_decay"], betas=g["betas"], eps=g["eps"],
                        decoupled=(self.opt.name == "adamw"), shadow=self.W16)

    # ------------------------------------------------------------------------------------ readback
    def copy_slot_to_host(self, slot):
        """Asynchronous D2H of one step's stats record into a pinned host ring (no host sync)."""
        if self.host_log is None:
            self.host_log = torch.zeros((self.log_slots, 2, 4), dtype=torch.float32).pin_memory()
        slot %= self.log_slots
        self.host_log[slot].copy_(self.stats_log[slot], non_blocking=True)

    def read_log(self, slots, from_host_ring=False):
        """Per-step stats for the given slots: one synchronising D2H of the device log, or - when every
        step already pushed its record with ``copy_slot_to_host`` - a stream sync and a host read."""
        if from_host_ring and self.host_log is not None:
            torch.cuda.current_stream().synchronize()
            raw = self.host_log.clone()
        else:
            raw = self.stats_log.cpu()
        ints = raw.view(torch.int32)
        out = []
        for s in slots:
            s %= self.log_slots
            has_i, has_t = self.slot_modalities[s]
            rec = {"image_loss": 0.0, "text_loss": 0.0, "img_acc": 0.0, "text_acc": 0.0}
            k = 0
            if has_i:
                n = max(1, int(ints[s, k, 3]))
                rec["image_loss"], rec["img_acc"] = float(raw[s, k, 0]), int(ints[s, k, 2]) / n
                k += 1
            if has_t:
                n = max(1, int(ints[s, k, 3]))
                rec["text_loss"], rec["text_acc"] = float(raw[s, k, 0]), int(ints[s, k, 2]) / n
            out.append(rec)
        return out
