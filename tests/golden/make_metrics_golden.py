"""Records tests/golden/metrics.npz from the UNMODIFIED reference metrics module (Gaussian_experiment/metrics.py, identical
to vision_language/metrics.py).  Build container only:

    python tests/golden/make_metrics_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("UML_REFERENCE_ROOT", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "Gaussian_experiment"))
from metrics import AlignmentMetrics  # noqa: E402  (the reference's own module)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics.npz")

CASES = {  # name: (rows, width of A, width of B, correlation of B with A)
    "small": (64, 10, 10, 0.7),
    "widths_differ": (200, 7, 12, 0.3),
    "independent": (150, 10, 10, 0.0),
    "identical": (100, 10, 10, 1.0),
    "gaussian_val": (2000, 10, 10, 0.5),   # the Gaussian experiment's validation set size and latent width
}


def make(seed, n, da, db, rho):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, da, generator=g)
    mix = torch.randn(da, db, generator=g)
    b = rho * (a @ mix) + (1.0 - rho) * torch.randn(n, db, generator=g) if rho < 1.0 else a.clone()
    return a, b


def main():
    out = {"torch_version": np.array(torch.__version__)}
    for i, (name, (n, da, db, rho)) in enumerate(CASES.items()):
        a, b = make(100 + i, n, da, db, rho)
        out[f"{name}/a"], out[f"{name}/b"] = a.numpy(), b.numpy()
        out[f"{name}/cka"] = np.float64(AlignmentMetrics.measure("cka", a, b, kernel_metric="ip"))
        out[f"{name}/mknn"] = np.float64(AlignmentMetrics.measure("mutual_knn", a, b, topk=10))
    np.savez_compressed(OUT, **out)
    print({k: float(v) for k, v in out.items() if k.endswith(("cka", "mknn"))})


if __name__ == "__main__":
    main()
