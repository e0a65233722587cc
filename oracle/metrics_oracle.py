"""TEST INFRASTRUCTURE - CPU restatement of the two alignment probes the reference logs next to its training loops
(not part of the optimisation result):

  * linear CKA with the biased HSIC estimator - ``AlignmentMetrics.cka(kernel_metric='ip', unbiased=False)``,
    ``metrics.py:96-119`` with ``hsic_biased`` ``:252-255`` (identical files under ``vision_language/`` and
    ``Gaussian_experiment/``); called from ``Gaussian_experiment/main.py:21-23,79`` and ``vision_language/finetune.py:112-114,232``;
  * mutual k-nearest-neighbour accuracy - ``AlignmentMetrics.mutual_knn``, ``metrics.py:55-86`` with
    ``compute_nearest_neighbors`` ``:272-285``; ``topk=10`` at both call sites.

Pinned by ``tests/golden/metrics.npz`` (recorded from the unmodified reference by ``tests/golden/make_metrics_golden.py``).
Only ``tests/`` may import this module.  The CUDA kernels for this row are not built yet; ``cka_linear_features`` states the
O(N d^2) form they will use, and the test checks it against the reference's O(N^3) form.
"""
from __future__ import annotations

import torch


def hsic_biased(K: torch.Tensor, L: torch.Tensor) -> torch.Tensor:
    """trace(K H L H) with H = I - 1/n  (metrics.py:252-255)."""
    n = K.shape[0]
    H = torch.eye(n, dtype=K.dtype) - 1.0 / n
    return torch.trace(K @ H @ L @ H)


def cka_linear(feats_a: torch.Tensor, feats_b: torch.Tensor) -> float:
    """The reference's formula, kernel matrices and all (metrics.py:100-119): hsic_kl / (sqrt(hsic_kk hsic_ll) + 1e-6)."""
    K, L = feats_a @ feats_a.T, feats_b @ feats_b.T
    kk, ll, kl = hsic_biased(K, K), hsic_biased(L, L), hsic_biased(K, L)
    return float(kl / (torch.sqrt(kk * ll) + 1e-6))


def cka_linear_features(feats_a: torch.Tensor, feats_b: torch.Tensor) -> float:
    """The same quantity without n x n matrices: for linear kernels trace(K H L H) = ||A_c^T B_c||_F^2 with
    column-centred features, so the three HSIC terms cost O(n d^2).  Accumulated in float64."""
    a = feats_a.double() - feats_a.double().mean(0, keepdim=True)
    b = feats_b.double() - feats_b.double().mean(0, keepdim=True)
    kl = (a.T @ b).pow(2).sum()
    kk = (a.T @ a).pow(2).sum()
    ll = (b.T @ b).pow(2).sum()
    return float(kl / (torch.sqrt(kk * ll) + 1e-6))


def nearest_neighbors(feats: torch.Tensor, topk: int) -> torch.Tensor:
    """Indices of the ``topk`` largest inner products per row, self excluded (metrics.py:272-285: the diagonal is set to
    -1e8 and the row argsorted in descending order)."""
    sim = (feats @ feats.T).fill_diagonal_(-1e8)
    return sim.argsort(dim=1, descending=True)[:, :topk]


def mutual_knn(feats_a: torch.Tensor, feats_b: torch.Tensor, topk: int = 10) -> float:
    """Mean over rows of |kNN_A(i) & kNN_B(i)| / topk  (metrics.py:55-86)."""
    ka, kb = nearest_neighbors(feats_a, topk), nearest_neighbors(feats_b, topk)
    n, k = ka.shape
    ma, mb = torch.zeros(n, n), torch.zeros(n, n)
    rows = torch.arange(n).unsqueeze(1)
    ma[rows, ka] = 1.0
    mb[rows, kb] = 1.0
    return float(((ma * mb).sum(1) / k).mean())
