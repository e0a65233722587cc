"""finetune.main(args, alphas=[...]): the alpha list of the reference's sweep YAML (configs/finetune.yaml:17) trained as ONE
batched job over banks loaded once - result files per alpha where one main() per alpha would have written them, and
every alpha's numbers identical to its own batched run."""
import os

import pytest
import torch

from oracle.synth import synth_banks

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import uml_b200  # noqa: F401
    from uml_b200 import features as F, finetune as ft
    from uml_b200.engine.config import parser
    from uml_b200.engine.optimizer.default import HYPER_DICT


def test_main_trains_the_alpha_list_in_one_group(tmp_path):
    C, D = 20, 64
    xi, yi, xt, yt, xv, yv = synth_banks(3, C, D, D, 16 * C, 6, 4 * C)
    g = torch.Generator().manual_seed(5)
    proto = torch.cat([xi[yi == c].mean(0, keepdim=True) for c in range(C)])
    yte = torch.randint(0, C, (500,), generator=g)
    xte = proto[yte] + 0.8 * torch.randn(500, D, generator=g)
    fdir = str(tmp_path / "features")
    lab2cname = {c: f"class_{c}" for c in range(C)}
    F.write_text_bank(F.text_outdir(fdir, "ViT-B/16", "synthset", "cupl"), xt, yt, lab2cname=lab2cname)
    F.write_image_bank(F.img_outdir(fdir, "ViT-B/16", "synthset", "crop", 16, 1, "train"), train=(xi, yi), val=(xv, yv),
                       lab2cname=lab2cname)
    F.write_image_bank(F.img_outdir(fdir, "ViT-B/16", "synthset", "crop", 16, 1, "test"), test=(xte, yte), lab2cname=lab2cname)
    HYPER_DICT["unit_test_a"] = dict(HYPER_DICT["clip_linear"], lr=[1e-3, 1e-4], weight_decay=[0.0], max_iter=[200], patience=[3])

    def args_for(rdir, alpha):
        return parser.parse_args(["--dataset", "synthset", "--train-shot", "16", "--seed", "1", "--clip-encoder", "ViT-B/16",
                                  "--modality", "crossmodal", "--text_type", "cupl", "--hyperparams", "unit_test_a",
                                  "--alpha", str(alpha), "--feature_dir", fdir, "--result_dir", str(tmp_path / rdir),
                                  "--num-workers", "0", "--sweep-batched"])

    outs = ft.main(args_for("exp_all", 0.2), alphas=[0.2, 1.0])
    assert len(outs) == 2 and all(len(o[0]["val_acc"]) == 2 for o in outs)
    found = [os.path.join(r, f) for r, _, fs in os.walk(str(tmp_path / "exp_all")) for f in fs]
    assert sum(f.endswith("test_result.pth") for f in found) == 4 and sum(f.endswith("results.pth") for f in found) == 2
    assert sum("alpha_0.2" in f and f.endswith("log.txt") for f in found) == 1
    assert sum("alpha_1.0" in f and f.endswith("log.txt") for f in found) == 1
    res1, _, _ = ft.main(args_for("exp_one", 1.0))
    assert res1["val_acc"] == outs[1][0]["val_acc"] and res1["test_acc"] == outs[1][0]["test_acc"]


def test_full_size_group_properties():
    """cfg2 shapes (16 000 x 512 image bank, 29 940-row text bank, 1000 classes, 32 + 32 rows per head and step): the
    oracle is too slow for a trajectory at this size, so check what must hold whatever the size - a head's result does
    not depend on its slot in the group or on its neighbours, twin heads stay bit-identical, a masked head is never
    touched, and the logged losses are finite and start at the zero-shot level."""
    from uml_b200.engine.datasets.utils import FeatureBank
    from uml_b200.engine.models.head import UMLClip
    from uml_b200.engine.optimizer.optim import build_optimizer
    from uml_b200.engine.sweep import HeadGroup
    dev = "cuda:0"
    C, D, B, steps = 1000, 512, 32, 12
    g = torch.Generator().manual_seed(7)
    ib = FeatureBank(torch.randn(16000, D, generator=g), torch.randint(0, C, (16000,), generator=g), dev)
    tb = FeatureBank(torch.randn(29940, D, generator=g), torch.arange(29940) % C, dev)
    W0 = torch.randn(C, D, generator=g) * 0.02
    hyper = [(1e-3, 0.0, 0.5), (1e-4, 0.01, 1.0), (1e-3, 0.001, 0.2), (1e-2, 0.0, 1.5), (1e-3, 0.0, 0.5), (1e-4, 0.0, 0.7)]
    perm_seed = [11, 12, 13, 14, 11, 16]   # heads 0 and 4 are twins: same hyper-parameters, same sampler stream
    active = [True, True, False, True, True, True]

    def run(order):
        models, opts, pi, pt = [], [], [], []
        for k in order:
            m = UMLClip(f"synthetic:{D}", C, logit_scale_init=3.0)
            m.load_state_dict({"head.weight": W0.clone()})
            m.to(dev)
            models.append(m)
            opts.append(build_optimizer(m.parameters(), "adamw", hyper[k][0], hyper[k][1]))
            gk = torch.Generator().manual_seed(perm_seed[k])
            pi.append(torch.randperm(16000, generator=gk).to(dev))
            pt.append(torch.randperm(29940, generator=gk).to(dev))
        group = HeadGroup(models, opts, ib, tb, B, B, dev, log_slots=steps)
        lrs = [[hyper[k][0] * (j + 1) / steps for k in order] for j in range(steps)]
        group.run(pi, pt, 0, 0, [(B, B)] * steps, lrs, [hyper[k][2] for k in order], [active[k] for k in order], slot0=0)
        torch.cuda.synchronize()
        log = group.read_log(list(range(steps)), True, True)
        return ({k: models[i].head.weight.detach().cpu().clone() for i, k in enumerate(order)},
                {k: [log["image_loss"][j][i] for j in range(steps)] for i, k in enumerate(order)})

    w_a, l_a = run([0, 1, 2, 3, 4, 5])
    w_b, l_b = run([5, 3, 4, 0, 2, 1])
    for k in range(6):
        assert torch.equal(w_a[k], w_b[k]), f"head {k} depends on its slot"
        if active[k]:
            assert l_a[k] == l_b[k]
            assert all(torch.isfinite(torch.tensor(l_a[k])))
            assert not torch.equal(w_a[k], W0)
    assert torch.equal(w_a[0], w_a[4]) and l_a[0] == l_a[4]
    assert torch.equal(w_a[2], W0)
    assert not torch.equal(w_a[0], w_a[1])


@pytest.mark.parametrize("bi,bt", [(48, 48), (65, 65), (64, 8)])
def test_sweep_run_more_than_one_row_tile(bi, bt):
    """Batches beyond 64 rows per head and step (preset full_ds_full_model_finetune uses 64 + 64): the logits launch gets a
    second row tile and the dW launch a second pass over the rows.  Against the oracle's step + optimizer."""
    from oracle import uml_oracle as O
    from uml_b200.engine.datasets.utils import FeatureBank
    from uml_b200.engine.models.head import UMLClip
    from uml_b200.engine.optimizer.optim import build_optimizer
    from uml_b200.engine.sweep import HeadGroup
    dev, K, C, D = "cuda:0", 3, 70, 96
    g = torch.Generator().manual_seed(bi * 100 + bt)
    xi, yi = torch.randn(400, D, generator=g), torch.randint(0, C, (400,), generator=g)
    xt, yt = torch.randn(300, D, generator=g), torch.randint(0, C, (300,), generator=g)
    W0 = [torch.randn(C, D, generator=g) * 0.05 for _ in range(K)]
    pi = [torch.randperm(400, generator=g) for _ in range(K)]
    pt = [torch.randperm(300, generator=g) for _ in range(K)]
    lr, wd, alpha = [1e-3, 1e-2, 1e-4], [0.0, 0.01, 0.1], [1.0, 0.5, 1.5]
    rows = [(bi, bt), (bi, bt - 3), (bi - 5, bt)]
    models, opts = [], []
    for k in range(K):
        m = UMLClip(f"synthetic:{D}", C, logit_scale_init=2.0)
        m.load_state_dict({"head.weight": W0[k].clone()})
        m.to(dev)
        models.append(m)
        opts.append(build_optimizer(m.parameters(), "adamw", lr[k], wd[k]))
    group = HeadGroup(models, opts, FeatureBank(xi, yi, dev), FeatureBank(xt, yt, dev), bi, bt, dev, log_slots=4)
    lrs = [[lr[k] for k in range(K)] for _ in rows]
    group.run([p.to(dev) for p in pi], [p.to(dev) for p in pt], 5, 2, rows, lrs, alpha, [True] * K, slot0=0)
    torch.cuda.synchronize()
    got = group.read_log([0, 1, 2], True, True)
    s = float(models[0].scales()[0])
    for k in range(K):
        st = O.HeadState(head=W0[k].clone(), img_scale=s, txt_scale=s)
        opt = O.OracleOptimizer(st.param_dict(), "adamw", lr[k], wd[k])
        a, b = 5, 2
        for i, (n_i, n_t) in enumerate(rows):
            ii, it = pi[k][a:a + n_i], pt[k][b:b + n_t]
            stats, grads = O.uml_step_grads(st, xi[ii], yi[ii], xt[it], yt[it], alpha[k])
            opt.step(grads, lr[k])
            a, b = a + n_i, b + n_t
            assert abs(got["image_loss"][i][k] - stats["image_loss"]) <= 1e-4 * max(1.0, abs(stats["image_loss"]))
            assert abs(got["text_loss"][i][k] - stats["text_loss"]) <= 1e-4 * max(1.0, abs(stats["text_loss"]))
            assert abs(got["img_acc"][i][k] - stats["img_acc"]) < 1e-6 and abs(got["text_acc"][i][k] - stats["text_acc"]) < 1e-6
        diff = (models[k].head.weight.detach().cpu() - st.head).abs() / st.head.abs().max()
        assert float(diff.max()) < 1e-3 and float(diff.mean()) < 5e-6, (k, float(diff.max()), float(diff.mean()))


def test_train_group_heads_of_different_lengths():
    """Same body as test_sweep_gpu.py::test_train_group_matches_reference_and_oracle with heads that end at different
    iterations, the num_workers > 0 sampler protocol and an evaluation interval of 7 (chunks cut at odd places)."""
    from sweep_case import run_group_case
    run_group_case(ft, "cuda:0", max_iters=[60, 33, 47], num_workers=2, eval_freq=7)
