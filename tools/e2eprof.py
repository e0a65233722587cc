"""Repeat bench.py's end-to-end arm (public train() call, per-step H2D/D2H) and profile the host side."""
import contextlib
import cProfile
import io
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import uml_b200  # noqa: F401,E402
from uml_b200 import finetune as ft  # noqa: E402
from uml_b200.engine.datasets.utils import BankLoader  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    dev = torch.device("cuda", 0)
    wl = bench.WORKLOADS["cfg3"]
    img_bank, txt_bank, val_bank, _ = bench.build_banks(wl, dev)
    B = wl["batch"]
    for r in range(reps):
        m2, o2, s2 = bench.make_model(wl, dev, txt_bank)
        il2 = BankLoader(img_bank, B, shuffle=True, upload="step")
        tl2 = BankLoader(txt_bank, wl["batch_txt"], shuffle=True, upload="step")
        vl2 = BankLoader(val_bank, 512, shuffle=False)
        torch.manual_seed(2)
        tr = {"timing": {"warmup": 5}, "indices": False}
        pr = cProfile.Profile() if r == reps - 1 else None
        with contextlib.redirect_stdout(io.StringIO()):
            if pr:
                pr.enable()
            ft.train(m2, il2, tl2, vl2, None, o2, s2, device=dev, max_iters=5 + K, alpha=0.5, eval_freq=10 ** 9,
                     patience=5, stats_to_host="step", trace=tr)
            if pr:
                pr.disable()
        t = tr["timing"]
        print(f"rep {r}: {t['seconds'] * 1e3:.2f} ms for {t['iters']} iters -> {t['rows'] / t['seconds'] / 1e6:.1f} M/s "
              f"({t['seconds'] / t['iters'] * 1e3:.4f} ms/step)")
        if pr:
            pstats.Stats(pr).sort_stats("tottime").print_stats(18)


if __name__ == "__main__":
    main()
