"""Time to bring an image bank from disk (page cache warm) into HBM: v1 pickled dict (torch.load + .to) against the
v2 mapped file (features.load_bank_v2).  Usage: python tools/bank_load_bench.py [rows] [dim]"""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uml_b200  # noqa: F401,E402
from uml_b200 import features as F  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
d = tempfile.mkdtemp()
feats, labels = torch.randn(rows, dim), torch.randint(0, 1000, (rows,))
v1 = os.path.join(d, "test.pth")
F.write_image_bank(v1, test=(feats, labels), lab2cname=None)
(v2,) = F.convert_bank(v1)
torch.cuda.init()
torch.zeros(1, device="cuda")


def t_v1():
    t0 = time.perf_counter()
    x = torch.load(v1, map_location="cpu")
    f, l = x["features"].to("cuda"), x["labels"].to("cuda")
    b = f.to(torch.bfloat16)  # the shadow the tensor-core path needs
    torch.cuda.synchronize()
    return time.perf_counter() - t0, f


def t_v2():
    t0 = time.perf_counter()
    t, _, _ = F.load_bank_v2(v2, "cuda", sections=("features", "features_bf16", "labels"))
    torch.cuda.synchronize()
    return time.perf_counter() - t0, t["features"]


for name, fn in (("v1 torch.load + to(cuda) + bf16 cast", t_v1), ("v2 mapped file -> pinned -> HBM (fp32 + bf16 + labels)", t_v2)):
    fn()
    best, ref = min((fn() for _ in range(3)), key=lambda r: r[0])
    assert torch.equal(ref.cpu(), feats)
    gb = rows * dim * 4 / 1e9
    print(f"{name}: {best * 1e3:.0f} ms for a {gb:.2f} GB fp32 bank ({rows} x {dim})")
