"""Debug: where the exchange forward kernel's warps spend their cycles (needs a -DUML_FWD_TIMING build:
UML_NVCC_FLAGS=-DUML_FWD_TIMING UML_OBJ_DIR=/tmp/objt UML_LIB_PATH=/path/libuml_timing.so python .../build.py --force)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200
from uml_b200 import ops, _lib
lib = C.CDLL(_lib.LIB_PATH)
DEV = "cuda:0"
N0, N1 = (int(v) for v in os.environ.get('FW_SHAPE', '34304+3584').split('+'))
D, Cc = 768, 1000
N = N0 + N1
x16 = torch.randn(N, D, device=DEV).to(torch.bfloat16)
W = torch.randn(Cc, D, device=DEV); W = W / W.norm(dim=1, keepdim=True)
w16 = ops.cast_bf16(W)
labels = torch.randint(0, Cc, (N,), device=DEV, dtype=torch.int32)
ws = ops.HeadWorkspace(N, Cc, DEV, bf16=True)
segs = ops.tc_segments([N0, N1], [100.0, 100.0], [1.0, 0.5])
names = ["epi_wait_tfull", "epi_first_ld", "epi_pass1+finish_prev", "epi_finish_rows", "epi_combine_publish", "epi_last_resolve", "epi_last_finish",
         "-", "-", "mma_wait_tempty", "mma_wait_full", "mma_issue", "tma_wait_empty", "tma_issue", "-", "-"]
for mode in ("train",):
    for _ in range(3):
        ops.head_fwd_ce_bf16(x16, w16, labels, segs, ws, None)
    torch.cuda.synchronize()
    lib.uml_debug_fwdx_timing(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.head_fwd_ce_bf16(x16, w16, labels, segs, ws, None); e1.record()
    torch.cuda.synchronize()
    buf = (C.c_longlong * (148 * 16))()
    lib.uml_debug_fwdx_timing(buf, 0)
    t = torch.tensor(list(buf)).view(148, 16).float()[:144]
    print(mode, "kernel us", e0.elapsed_time(e1) * 1e3)
    for i, n in enumerate(names):
        print(f"  {n:18s} mean {t[:, i].mean().item():10.0f}  max {t[:, i].max().item():10.0f} cycles (CTA0 {t[0, i].item():.0f}, CTA1 {t[1, i].item():.0f}, CTA2 {t[2, i].item():.0f})")
