"""Debug: where the forward kernel's warps spend their cycles (needs a -DUML_FWD_TIMING build)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200
from uml_b200 import ops, _lib
lib = C.CDLL(_lib.LIB_PATH)
DEV = "cuda:0"
B, D, Cc = 16384, 768, 1000
N = 2 * B
x16 = torch.randn(N, D, device=DEV).to(torch.bfloat16)
W = torch.randn(Cc, D, device=DEV); W = W / W.norm(dim=1, keepdim=True)
w16 = ops.cast_bf16(W)
labels = torch.randint(0, Cc, (N,), device=DEV, dtype=torch.int32)
ws = ops.HeadWorkspace(N, Cc, DEV, bf16=True)
segs = ops.tc_segments([B, B], [100.0, 100.0], [1.0, 0.5])
names = ["prod_wait_empty", "prod_issue", "-", "-", "mma_wait_tempty", "mma_wait_full", "mma_issue", "-", "epi_wait_tfull", "epi_chunk_work", "epi_tile_tail", "-", "norm_wait", "norm_work", "-", "-"]
for mode, wsx in (("eval", None), ("train", ws)):
    for _ in range(3):
        ops.head_fwd_ce_bf16(x16, w16, labels, segs, wsx, ws.row_loss)
    torch.cuda.synchronize()
    lib.uml_debug_fwd_timing(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.head_fwd_ce_bf16(x16, w16, labels, segs, wsx, ws.row_loss); e1.record()
    torch.cuda.synchronize()
    buf = (C.c_longlong * (148 * 16))()
    lib.uml_debug_fwd_timing(buf, 0)
    t = torch.tensor(list(buf)).view(148, 16).float()
    print(mode, "kernel us", e0.elapsed_time(e1) * 1e3, "cg", os.environ.get("UML_TC_CTA_GROUP", "auto"))
    for i, n in enumerate(names):
        print(f"  {n:18s} mean {t[:, i].mean().item():10.0f}  max {t[:, i].max().item():10.0f} cycles (CTA0 {t[0, i].item():.0f}, CTA1 {t[1, i].item():.0f})")
