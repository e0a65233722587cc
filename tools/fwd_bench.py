"""Forward-kernel timing at several batch sizes (CUDA events, L2 flushed between launches)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200  # noqa
from uml_b200 import ops

DEV = "cuda:0"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
D, C = 768, 1000
for N0, N1 in [tuple(int(v) for v in s.split("+")) for s in os.environ.get("FB_SHAPES", "34304+3584,33280+3584,37888+3584").split(",")]:
    N = N0 + N1
    x16 = (torch.randn(N, D, device=DEV) * 1.0).to(torch.bfloat16)
    W = torch.randn(C, D, device=DEV); W = W / W.norm(dim=1, keepdim=True)
    w16 = ops.cast_bf16(W)
    labels = torch.randint(0, C, (N,), device=DEV, dtype=torch.int32)
    ws = ops.HeadWorkspace(N, C, DEV, bf16=True)
    segs = ops.tc_segments([N0, N1], [100.0, 100.0], [1.0, 0.5])
    for mode, w in (("train", ws), ("eval", None)):
        ts = []
        for i in range(13):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.head_fwd_ce_bf16(x16, w16, labels, segs, w, ws.row_loss, row_correct=ws.row_correct); b.record()
            torch.cuda.synchronize()
            if i >= 3: ts.append(a.elapsed_time(b))
        ts.sort()
        print(f"rows {N} {mode}: median {ts[len(ts)//2]*1e3:.1f} us  best {ts[0]*1e3:.1f} us  -> {2.0*N*D*C/ts[len(ts)//2]/1e9:.0f} TFLOP/s")
