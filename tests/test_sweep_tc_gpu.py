"""The sweep kernels' tensor-core forms (tcgen05 kind::tf32, every operand split into two tf32 terms) against their FFMA
forms: the same group of heads run in two fresh processes, one with UML_SWEEP_TC=0.  The library reads the switch once per
process, hence the subprocesses.  What is asserted is the accuracy claim of csrc/sweep.cu - the three-term product stays at
fp32 level, i.e. the two trajectories differ by summation-order noise, not by tf32's 2^-11."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
import uml_b200  # noqa: F401
from uml_b200.engine.datasets.utils import FeatureBank
from uml_b200.engine.models.head import UMLClip
from uml_b200.engine.optimizer.optim import build_optimizer
from uml_b200.engine.sweep import HeadGroup
dev, K, C, D, B, steps = "cuda:0", 4, 1000, 512, 32, 6
g = torch.Generator().manual_seed(5)
ib = FeatureBank(torch.randn(4000, D, generator=g), torch.randint(0, C, (4000,), generator=g), dev)
tb = FeatureBank(torch.randn(3000, D, generator=g), torch.arange(3000) %% C, dev)
W0 = torch.randn(C, D, generator=g) * 0.02
models, opts, pi, pt = [], [], [], []
for k in range(K):
    m = UMLClip(f"synthetic:{D}", C, logit_scale_init=3.0)
    m.load_state_dict({"head.weight": W0.clone()})
    m.to(dev)
    models.append(m)
    # SGD: the update is linear in the gradient, so the weights show the kernels' accuracy (Adam's ~lr * sign(g) steps would
    # turn any rounding difference of a nearly cancelling gradient into a visible step)
    opts.append(build_optimizer(m.parameters(), "sgd", [1e-3, 1e-4, 1e-2, 1e-3][k], [0.0, 0.01, 0.001, 0.0][k]))
    pi.append(torch.randperm(4000, generator=g).to(dev))
    pt.append(torch.randperm(3000, generator=g).to(dev))
group = HeadGroup(models, opts, ib, tb, B, B, dev, log_slots=steps)
lrs = [[[1e-3, 1e-4, 1e-2, 1e-3][k] for k in range(K)] for _ in range(steps)]
group.run(pi, pt, 0, 0, [(B, B)] * (steps - 1) + [(B, 17)], lrs, [0.5, 1.0, 0.2, 1.5], [True] * K, slot0=0)
torch.cuda.synchronize()
log = group.read_log(list(range(steps)), True, True)
np.savez(sys.argv[1], W0=W0.numpy(), W=torch.stack([m.head.weight.detach().cpu() for m in models]).numpy(),
         il=np.array(log["image_loss"]), tl=np.array(log["text_loss"]), ia=np.array(log["img_acc"]), ta=np.array(log["text_acc"]))
'''


@pytest.mark.gpu
def test_tensor_core_sweep_kernels_stay_at_fp32_level(tmp_path):
    outs = {}
    for tc in ("1", "0"):
        out = str(tmp_path / f"tc{tc}.npz")
        env = dict(os.environ, UML_SWEEP_TC=tc)
        r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}, out], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tc] = np.load(out)
    a, b = outs["1"], outs["0"]
    np.testing.assert_allclose(a["il"], b["il"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(a["tl"], b["tl"], rtol=2e-5, atol=2e-5)
    assert np.array_equal(a["ia"], b["ia"]) and np.array_equal(a["ta"], b["ta"])
    # the accumulated SGD updates of the two runs, relative to the largest update of the head: fp32-level agreement
    # (one tf32 term per operand would sit at ~1e-3)
    for k in range(a["W"].shape[0]):
        ua, ub = a["W"][k] - a["W0"], b["W"][k] - b["W0"]
        d = np.abs(ua - ub) / np.abs(ub).max()
        assert d.max() < 2e-5 and d.mean() < 1e-6, (k, d.max(), d.mean())
    assert not np.array_equal(a["W"], b["W"]), "the two runs must have taken different kernels"
