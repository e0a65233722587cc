"""a-14: the Gaussian_experiment step on the GPU (csrc/gauss.cu through uml_b200.gaussian) against the golden
trace of the UNMODIFIED reference (tests/golden/misc.npz, recorded by make_golden.py) and against the oracle at
the reference's real sizes (train.yaml: dim_obs 50, dim_common 128, dim_latent 10, batch 512)."""
import ast
import os
import types

import numpy as np
import pytest
import torch

from oracle import uml_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"

if torch.cuda.is_available():
    import uml_b200  # noqa: F401
    from uml_b200 import gaussian as G


def _args(mode, ax, ay):
    return types.SimpleNamespace(mode=mode, alpha_x=ax, alpha_y=ay)


@pytest.mark.parametrize("mode", ["xy", "x"])
def test_gaussian_step_matches_reference_trace(mode):
    m = np.load(os.path.join(GOLDEN, "misc.npz"), allow_pickle=False)
    cfg = ast.literal_eval(str(m["gauss/cfg"]))
    base = dict(seed=cfg["seed"], num_samples=cfg["num_samples"], dim_c=cfg["dim_c"], dim_x=cfg["dim_x"], dim_y=cfg["dim_y"],
                dim_obs=cfg["dim_obs"], noise_std=cfg["noise_std"], attenuate_x=True, attenuation=cfg["attenuation"])
    d1 = G.generate_data(dict(base, shared_latent_distribution_type="gaussian"))
    d2 = G.generate_data(dict(base, seed=44, shared_latent_distribution_type="laplace"))
    dx, dy = (d1["x"][:32], d1["y"][:32]) if mode == "xy" else (d1["x"], d2["y"][:40])
    torch.manual_seed(3)
    model = G.SharedAutoencoder(12, 16, 6, device=DEV)
    for k, v in model.state_dict().items():  # same init draws as the reference constructor
        np.testing.assert_array_equal(v.cpu().numpy(), m[f"gauss/{mode}/init/{k}"], err_msg=k)
    g = torch.Generator()
    g.manual_seed(42)
    loader = G.unpaired_loader(G.UnpairedDataset(dx, dy, DEV), 16, generator=g)
    opt = G.Adam(model, lr=1e-3)
    hist = G.train_model_steps(model, loader, opt, 9, args=_args(mode, 1.0, 0.5))
    np.testing.assert_allclose(hist["loss_x"], m[f"gauss/{mode}/loss_x"], rtol=2e-5)
    np.testing.assert_allclose(hist["loss_y"], m[f"gauss/{mode}/loss_y"], rtol=2e-5)
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), m[f"gauss/{mode}/final/{k}"], rtol=2e-4, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("mode", ["xy", "x"])
def test_gaussian_step_matches_oracle_at_reference_size(mode):
    kw = dict(seed=42, num_samples=3000, dim_c=10, dim_x=5, dim_y=5, dim_obs=50, noise_std=0.09, attenuate_x=True,
              attenuation=0.05)
    d = G.generate_data(dict(kw, shared_latent_distribution_type="gaussian"))
    dx, dy = d["x"][:1500], d["y"][:1400]          # different lengths: the index wrap of UnpairedDataset matters
    torch.manual_seed(0)
    model = G.SharedAutoencoder(50, 128, 10, device=DEV)
    p = {k: v.cpu().clone() for k, v in model.state_dict().items()}
    hist_o = O.gaussian_train(p, dx, dy, num_steps=25, batch_size=512, lr=1e-3, mode=mode, alpha_x=1.0, alpha_y=0.7)
    g = torch.Generator()
    g.manual_seed(42)
    loader = G.unpaired_loader(G.UnpairedDataset(dx, dy, DEV), 512, generator=g)
    trace = {}
    hist = G.train_model_steps(model, loader, G.Adam(model, lr=1e-3), 25, val_data_x=d["x"][2000:2600], val_data_y=d["y"][2000:2600],
                               args=_args(mode, 1.0, 0.7), eval_every=25, trace=trace)
    for s in range(25):  # bit-exact sampler stream (explicit generator, drop_last)
        assert np.array_equal(trace["idx"][s].numpy(), hist_o[s][2])
    np.testing.assert_allclose(hist["loss_x"], [h[0] for h in hist_o], rtol=2e-5)
    np.testing.assert_allclose(hist["loss_y"], [h[1] for h in hist_o], rtol=2e-5)
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), p[k].numpy(), rtol=5e-4, atol=5e-6, err_msg=k)
    if mode == "x":  # the y heads were never touched
        torch.manual_seed(0)
        fresh = G.SharedAutoencoder(50, 128, 10, device=DEV).state_dict()
        for k in ("in_head_y.weight", "in_head_y.bias", "out_head_y.weight", "out_head_y.bias"):
            assert torch.equal(model.state_dict()[k], fresh[k])
    # validation forward against the oracle's forward on the trained weights
    (step, vx, vy), = hist["val"]
    lx, _, _, _ = O._gauss_branch(p, d["x"][2000:2600], "x")
    ly, _, _, _ = O._gauss_branch(p, d["y"][2000:2600], "y")
    assert abs(vx - float(lx)) <= 1e-4 * abs(float(lx)) and abs(vy - float(ly)) <= 1e-4 * abs(float(ly))
