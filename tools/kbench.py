"""Kernel micro-benchmarks (CUDA events, L2 flushed between timed launches)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200  # noqa
from uml_b200 import ops

DEV = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    B, D, C = int(os.environ.get("KB_B", 16384)), int(os.environ.get("KB_D", 768)), int(os.environ.get("KB_C", 1000))
    N = 2 * B
    res = {}
    bank = torch.randn(1_281_167 if D == 768 else 200_000, D, device=DEV)
    idx = torch.randint(0, bank.shape[0], (N,), device=DEV)
    x16 = torch.empty(N, D, dtype=torch.bfloat16, device=DEV)
    med, best = timeit(lambda: ops.gather_rows(bank, idx, out=x16))
    res["gather_bf16"] = dict(ms=med, best_ms=best, GBs=N * (D * 4 + D * 2 + 8) / med / 1e6)
    x32 = torch.empty(N, D, device=DEV)
    med, best = timeit(lambda: ops.gather_rows(bank, idx, out=x32))
    res["gather_f32"] = dict(ms=med, best_ms=best, GBs=N * (D * 8 + 8) / med / 1e6)
    W = torch.randn(C, D, device=DEV)
    W = W / W.norm(dim=1, keepdim=True)
    w16 = ops.cast_bf16(W)
    labels = torch.randint(0, C, (N,), device=DEV, dtype=torch.int32)
    ws = ops.HeadWorkspace(N, C, DEV, bf16=True)
    segs = ops.tc_segments([B, B], [100.0, 100.0], [1.0, 0.5])
    med, best = timeit(lambda: ops.head_fwd_ce_bf16(x16, w16, labels, segs, ws, ws.row_loss, row_correct=ws.row_correct))
    res["fwd_tc"] = dict(ms=med, best_ms=best, TFs=2.0 * N * D * C / med / 1e9)
    med, best = timeit(lambda: ops.head_fwd_ce_bf16(x16, w16, labels, segs, None, ws.row_loss, row_correct=ws.row_correct))
    res["fwd_tc_eval"] = dict(ms=med, best_ms=best, TFs=2.0 * N * D * C / med / 1e9)
    splits = ops.tc_dw_splits(N, D, C)
    parts = torch.empty(splits, C, D, device=DEV)
    med, best = timeit(lambda: ops.head_bwd_dw_bf16(ws.G, ws.ldg, x16, N, C, parts, splits))
    res["dw_tc"] = dict(ms=med, best_ms=best, TFs=2.0 * N * D * C / med / 1e9, splits=splits)
    m, v = torch.zeros_like(W), torch.zeros_like(W)
    med, best = timeit(lambda: ops.adamw_step_partials(W, parts, splits, m, v, lr=1e-3, step=1, weight_decay=0.01, shadow=w16))
    res["adamw_partials"] = dict(ms=med, best_ms=best, GBs=(C * D * (splits * 4 + 12 + 12 + 2)) / med / 1e6)
    g = torch.randn_like(W)
    med, best = timeit(lambda: ops.adamw_step(W, g, m, v, lr=1e-3, step=1, weight_decay=0.01))
    res["adamw"] = dict(ms=med, best_ms=best, GBs=(C * D * 28) / med / 1e6)
    big = torch.randn(64 * 1024 * 1024, device=DEV)
    bm, bv, bg = torch.zeros_like(big), torch.zeros_like(big), torch.randn_like(big)
    med, best = timeit(lambda: ops.adamw_step(big, bg, bm, bv, lr=1e-3, step=1, weight_decay=0.01), iters=5)
    res["adamw_64M"] = dict(ms=med, best_ms=best, GBs=(big.numel() * 28) / med / 1e6)
    # torch references for scale
    a = torch.randn(N, D, device=DEV, dtype=torch.bfloat16)
    med, best = timeit(lambda: torch.matmul(a, w16.t()))
    res["torch_matmul_fwd_bf16"] = dict(ms=med, TFs=2.0 * N * D * C / med / 1e9)
    med, best = timeit(lambda: torch.matmul(ws.G[:, :C].t(), a))
    res["torch_matmul_dw_bf16"] = dict(ms=med, TFs=2.0 * N * D * C / med / 1e9)
    # whole step through the single C call (host dispatch cost included)
    from uml_b200.engine.datasets.utils import BankLoader, FeatureBank
    from uml_b200.engine.models.head import UMLClip
    from uml_b200.engine.optimizer.optim import build_optimizer
    from uml_b200.engine.trainer import StepEngine
    from uml_b200 import finetune as ft
    import time
    fb = FeatureBank.__new__(FeatureBank); fb.features = bank; fb.labels = torch.randint(0, C, (bank.shape[0],), device=DEV)
    model = UMLClip(f"s:{D}", C, logit_scale_init=4.60517).to(DEV)
    opt = build_optimizer(model.parameters(), "adamw", 1e-3, 0.01)
    eng = StepEngine(model, opt, DEV, B, B, precision="auto")
    il, tl = BankLoader(fb, B, shuffle=True), BankLoader(fb, B, shuffle=True)
    ii, ti = iter(il), iter(tl)
    def one(i):
        nonlocal ii, ti
        a, ii = ft.fetch_next(il, ii); b, ti = ft.fetch_next(tl, ti)
        eng.step(a, b, 0.5, slot=i)
    for i in range(5): one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50): one(i)
    host = (time.perf_counter() - t0) / 50 * 1e3
    e1.record(); torch.cuda.synchronize()
    res["step_single_call"] = dict(ms=e0.elapsed_time(e1) / 50, host_ms=host, TFs=4.0 * N * D * C / (e0.elapsed_time(e1) / 50) / 1e9)
    for k, val in res.items():
        print(k, json.dumps({a: (round(b, 4) if isinstance(b, float) else b) for a, b in val.items()}))


if __name__ == "__main__":
    main()
