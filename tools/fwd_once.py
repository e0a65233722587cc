"""A few launches of the forward kernel at the bench shape (for ncu): FW_MODE=train|eval."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200  # noqa
from uml_b200 import ops
DEV = "cuda:0"
N0, N1 = (int(v) for v in os.environ.get("FW_SHAPE", "33280+3584").split("+"))
D, C, N = 768, 1000, N0 + N1
x16 = torch.randn(N, D, device=DEV).to(torch.bfloat16)
W = torch.randn(C, D, device=DEV); W = W / W.norm(dim=1, keepdim=True)
w16 = ops.cast_bf16(W)
labels = torch.randint(0, C, (N,), device=DEV, dtype=torch.int32)
ws = ops.HeadWorkspace(N, C, DEV, bf16=True)
segs = ops.tc_segments([N0, N1], [100.0, 100.0], [1.0, 0.5])
for mode in os.environ.get("FW_MODE", "train,eval").split(","):
    for _ in range(3):
        ops.head_fwd_ce_bf16(x16, w16, labels, segs, ws if mode == "train" else None, ws.row_loss, row_correct=ws.row_correct)
torch.cuda.synchronize()
print("ok")
