"""Host-side cost of enqueueing UML steps (cProfile over the bench loop at a batch small enough that the GPU
never back-pressures the launch queue).  Usage: python tools/hostprof.py [B] [steps] [chunk]"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import uml_b200  # noqa: F401,E402
from uml_b200 import finetune as ft  # noqa: E402
from uml_b200.engine.datasets.utils import BankLoader  # noqa: E402
from uml_b200.engine.trainer import StepEngine  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda", 0)
    wl = dict(bench.WORKLOADS["cfg3"], batch=B, batch_txt=B, n_img=int(os.environ.get("N_IMG", 200_000)))
    img_bank, txt_bank, _, _ = bench.build_banks(wl, dev)
    model, opt, sch = bench.make_model(wl, dev, txt_bank)
    engine = StepEngine(model, opt, dev, B, B, log_slots=64, precision="bf16")
    il = BankLoader(img_bank, B, shuffle=True, upload=os.environ.get("UPLOAD", "epoch"))
    tl = BankLoader(txt_bank, B, shuffle=True, upload=os.environ.get("UPLOAD", "epoch"))
    torch.manual_seed(2)
    its = [iter(il), iter(tl)]

    def run(n_steps):
        i = 0
        while i < n_steps:
            batches, lrs = [], []
            for _ in range(chunk):
                img, its[0] = ft.fetch_next(il, its[0])
                txt, its[1] = ft.fetch_next(tl, its[1])
                batches.append((img, txt))
                lrs.append(sch.get_last_lr()[0])
                sch.step()
            engine.run(batches, 0.5, lrs, slot0=i)
            i += chunk

    run(50)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(f"B={B} chunk={chunk}: host {host / steps * 1e6:.1f} us/step, wall {total / steps * 1e6:.1f} us/step")
    pr = cProfile.Profile()
    pr.enable()
    run(steps)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
