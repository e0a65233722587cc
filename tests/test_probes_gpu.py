"""the alignment probes on the device (csrc/probes.cu) against the golden vectors recorded from the UNMODIFIED
reference's metrics module (tests/golden/metrics.npz, make_metrics_golden.py) and against the oracle restatement
(oracle/metrics_oracle.py) on shapes the goldens do not cover; the Gaussian autoencoder's embedding forward against torch."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics.npz")

if torch.cuda.is_available():
    import uml_b200  # noqa: F401
    from uml_b200 import gaussian as G, ops


@pytest.mark.parametrize("case", ["small", "widths_differ", "independent", "identical", "gaussian_val"])
def test_probes_match_the_reference_goldens(case):
    fx = np.load(GOLDEN, allow_pickle=False)
    a = torch.from_numpy(fx[f"{case}/a"]).cuda()
    b = torch.from_numpy(fx[f"{case}/b"]).cuda()
    cka = float(ops.cka_linear(a, b).item())
    mknn = float(ops.mutual_knn(a, b, 10).item())
    # CKA: fp64 accumulation of fp32 products against the reference's fp32 n x n matrix products: 2e-4 relative
    assert abs(cka - float(fx[f"{case}/cka"])) <= 2e-4 * max(1.0, abs(float(fx[f"{case}/cka"]))), (cka, float(fx[f"{case}/cka"]))
    # mutual kNN: a neighbour list can differ where two inner products tie to fp32 rounding: at most a few rows of n
    n = a.shape[0]
    assert abs(mknn - float(fx[f"{case}/mknn"])) <= 3.0 / (n * 10) + 1e-7, (mknn, float(fx[f"{case}/mknn"]))


def test_probes_on_wide_features_against_the_oracle():
    """The vision-language call site compares class-mean image features with text features at the encoders' width."""
    g = torch.Generator().manual_seed(2)
    n, da, db = 397, 512, 384
    base = torch.randn(n, 64, generator=g)
    a = base @ torch.randn(64, da, generator=g) + 0.3 * torch.randn(n, da, generator=g)
    b = base @ torch.randn(64, db, generator=g) + 0.3 * torch.randn(n, db, generator=g) + 2.0   # a non-zero mean as well
    cka = float(ops.cka_linear(a.cuda(), b.cuda()).item())
    want = M.cka_linear_features(a, b)
    assert abs(cka - want) <= 1e-4 * max(1.0, abs(want)), (cka, want)
    mknn = float(ops.mutual_knn(a.cuda(), b.cuda(), 10).item())
    assert abs(mknn - M.mutual_knn(a, b, 10)) <= 3.0 / (n * 10) + 1e-7
    # strided views (row pitch larger than the width) are taken as they are
    wide = torch.randn(n, da + 32, generator=g).cuda()
    assert float(ops.cka_linear(wide[:, :da], wide[:, :da]).item()) == pytest.approx(1.0, abs=1e-4)


def test_gaussian_embeddings_and_alignment_through_the_training_call():
    torch.manual_seed(0)
    model = G.SharedAutoencoder(50, 128, 10, device="cuda")
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(777, 50, generator=g), torch.randn(777, 50, generator=g)
    ex, ey = G.get_embeddings(model, x, y)

    def ref(v, head):
        z = v @ sd[f"{head}.weight"].T + sd[f"{head}.bias"]
        h = torch.relu(z @ sd["shared_encoder.0.weight"].T + sd["shared_encoder.0.bias"])
        return h @ sd["shared_encoder.2.weight"].T + sd["shared_encoder.2.bias"]

    torch.testing.assert_close(ex.cpu(), ref(x, "in_head_x"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ey.cpu(), ref(y, "in_head_y"), rtol=1e-4, atol=1e-5)
    only_x, none_y = G.get_embeddings(model, x, None)
    assert none_y is None and torch.equal(only_x, ex)
    # the training call logs the probes at every evaluation point (main.py:67-84)
    data = G.generate_data(dict(seed=42, num_samples=2000, dim_c=10, dim_x=5, dim_y=5, dim_obs=50, noise_std=0.09, attenuate_x=True,
                                attenuation=0.05, shared_latent_distribution_type="gaussian"))
    ds = G.UnpairedDataset(data["x"][:1500], data["y"][:1500], "cuda")
    loader = G.unpaired_loader(ds, 128, generator=torch.Generator().manual_seed(3))
    import types
    out = G.train_model_steps(model, loader, G.Adam(model, 1e-3), 40, val_data_x=data["x"][1500:], val_data_y=data["y"][1500:],
                              args=types.SimpleNamespace(mode="xy", alpha_x=1.0, alpha_y=1.0), eval_every=20)
    assert [s for s, _, _ in out["align"]] == [19, 39] == [s for s, _, _ in out["val"]]
    exv, eyv = G.get_embeddings(model, data["x"][1500:], data["y"][1500:])
    want_cka, want_mknn = M.cka_linear_features(exv.cpu(), eyv.cpu()), M.mutual_knn(exv.cpu(), eyv.cpu(), 10)
    assert abs(out["align"][-1][1] - want_cka) <= 2e-4 and abs(out["align"][-1][2] - want_mknn) <= 3.0 / 5000 + 1e-7


def test_alignment_probe_of_the_vision_language_loop():
    """finetune.alignment_probe (the reference's feature probes, finetune.py:209-233) against the oracle's formulas."""
    from uml_b200 import finetune as ft
    from uml_b200.engine.models.head import UML

    torch.manual_seed(0)
    C, Dv, D = 23, 48, 32
    model = UML(f"synthetic:{Dv}", D, C)
    model.to("cuda")
    g = torch.Generator().manual_seed(6)
    xi, yi = torch.randn(C * 4, Dv, generator=g), torch.arange(C * 4) % C
    text_c, text_n = torch.randn(C, D, generator=g), torch.randn(C * 4, D, generator=g)
    feats = (xi @ model.img_proj.weight.detach().cpu().T)
    means = torch.stack([feats[yi == c].mean(0) for c in range(C)])
    got = ft.alignment_probe(model, xi.cuda(), yi.cuda(), text_c.cuda(), C)
    assert set(got) == {"inclass_distance", "cka"}
    assert abs(got["cka"] - M.cka_linear_features(means, text_c)) <= 2e-4
    want_d = float(torch.stack([(feats[yi == c] - means[c]).norm(dim=1).mean() for c in range(C)]).mean())
    assert abs(got["inclass_distance"] - want_d) <= 1e-4 * want_d
    got = ft.alignment_probe(model, xi.cuda(), yi.cuda(), text_n.cuda(), C)
    assert set(got) == {"inclass_distance", "mknn"}
    assert abs(got["mknn"] - M.mutual_knn(feats, text_n, 10)) <= 3.0 / (C * 40) + 1e-7
