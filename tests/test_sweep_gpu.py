"""Sweep-level batching on the GPU: uml_sweep_run through the C ABI against the oracle / a torch fp32
reference, and finetune.train_group against the reference's golden trace (same body as the CPU host-logic test, here
with the real kernels)."""
import numpy as np
import pytest
import torch

from oracle import uml_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

if torch.cuda.is_available():
    import uml_b200  # noqa: F401
    from uml_b200 import finetune as ft
    from uml_b200.engine.datasets.utils import FeatureBank
    from uml_b200.engine.models.head import UMLClip
    from uml_b200.engine.optimizer.optim import build_optimizer
    from uml_b200.engine.sweep import HeadGroup

from sweep_case import run_group_case


def test_train_group_matches_reference_and_oracle():
    outs, traces, stopped = run_group_case(ft, DEV)
    assert stopped[2] and not stopped[0]
    assert traces[0]["engine"].launches > 0


def test_train_group_sgd_image_only():
    run_group_case(ft, DEV, optim="sgd", modality="image")


def test_train_group_adam_text_only():
    run_group_case(ft, DEV, optim="adam", modality="text")


@pytest.mark.parametrize("C,D,optim", [(37, 50, "adamw"), (64, 128, "adamw"), (100, 36, "sgd"), (1000, 512, "adam"),
                                       (300, 32, "sgd"), (650, 64, "adamw"), (800, 64, "adamw"),   # clusters of 3, 6 and 7 class tiles
                                       (1100, 32, "adamw")])                                        # 9 tiles: separate softmax launch
def test_sweep_run_against_oracle(C, D, optim):
    """K heads (one of them switched off) x 3 steps with ragged batches through uml_sweep_run; every head against the
    oracle's step + optimizer on the same rows.  Shapes cover the vector (dim % 4 == 0) and scalar epilogues and tiles
    that overhang every edge."""
    K, n_img, n_txt = 5, 300, 120
    g = torch.Generator().manual_seed(C * 1000 + D)
    xi, yi = torch.randn(n_img, D, generator=g), torch.randint(0, C, (n_img,), generator=g)
    xt, yt = torch.randn(n_txt, D, generator=g), torch.randint(0, C, (n_txt,), generator=g)
    W0 = [torch.randn(C, D, generator=g) * 0.05 for _ in range(K)]
    perms_i = [torch.randperm(n_img, generator=g) for _ in range(K)]
    perms_t = [torch.randperm(n_txt, generator=g) for _ in range(K)]
    lrs_base = [1e-3, 3e-3, 1e-2, 1e-4, 5e-3]
    wds = [0.0, 0.01, 0.001, 0.1, 0.0]
    alphas = [1.0, 0.5, 0.2, 1.5, 0.7]
    active = [True, True, False, True, True]
    rows = [(32, 32), (32, 17), (5, 32)]
    pos_i, pos_t = 11, 3
    scale = 30.0

    models, opts = [], []
    for k in range(K):
        m = UMLClip(f"synthetic:{D}", C, logit_scale_init=float(np.log(scale)))
        m.load_state_dict({"head.weight": W0[k].clone()})
        m.to(DEV)
        models.append(m)
        opts.append(build_optimizer(m.parameters(), optim, lrs_base[k], wds[k]))
    group = HeadGroup(models, opts, FeatureBank(xi, yi, DEV), FeatureBank(xt, yt, DEV), 32, 32, DEV, log_slots=8)
    lrs = [[lrs_base[k] * (1.0 - 0.1 * i) for k in range(K)] for i in range(len(rows))]
    group.run([p.to(DEV) for p in perms_i], [p.to(DEV) for p in perms_t], pos_i, pos_t, rows, lrs, alphas, active, slot0=2)
    torch.cuda.synchronize()
    got = group.read_log([2, 3, 4], True, True)
    s_model = float(models[0].scales()[0])
    for k in range(K):
        w = models[k].head.weight.detach().cpu()
        if not active[k]:
            assert torch.equal(w, W0[k]), "an inactive head must not be touched"
            continue
        st = O.HeadState(head=W0[k].clone(), img_scale=s_model, txt_scale=s_model)
        opt = O.OracleOptimizer(st.param_dict(), optim, lrs_base[k], wds[k])
        pi, pt = pos_i, pos_t
        for i, (n_i, n_t) in enumerate(rows):
            ii, it = perms_i[k][pi:pi + n_i], perms_t[k][pt:pt + n_t]
            stats, grads = O.uml_step_grads(st, xi[ii], yi[ii], xt[it], yt[it], alphas[k])
            opt.step(grads, lrs[i][k])
            pi, pt = pi + n_i, pt + n_t
            assert abs(got["image_loss"][i][k] - stats["image_loss"]) <= 1e-4 * max(1.0, abs(stats["image_loss"]))
            assert abs(got["text_loss"][i][k] - stats["text_loss"]) <= 1e-4 * max(1.0, abs(stats["text_loss"]))
            assert abs(got["img_acc"][i][k] - stats["img_acc"]) < 1e-6
            assert abs(got["text_acc"][i][k] - stats["text_acc"]) < 1e-6
        # Adam divides by |g| + eps: the few elements whose gradient (plus L2 term) nearly cancels amplify fp32
        # summation-order noise, so the maximum gets the stated 1e-3 and the mean a much tighter bound
        diff = (w - st.head).abs() / st.head.abs().max()
        assert float(diff.max()) < 1e-3 and float(diff.mean()) < 5e-6, (k, float(diff.max()), float(diff.mean()))
        assert opts[k].slot(models[k].head.weight)["step"] == len(rows)


def test_sweep_run_matches_single_head_engine():
    """One head of a group follows the same trajectory as the single-head fp32 engine (finetune.train) on the same
    sampler stream: losses to summation-order noise, identical evaluation schedule and best iteration."""
    from uml_b200.engine.datasets.utils import BankLoader
    from uml_b200.engine.optimizer.scheduler import build_lr_scheduler
    from oracle.synth import synth_banks
    C, D = 40, 64
    xi, yi, xt, yt, xv, yv = synth_banks(3, C, D, D, 400, 5, 200)
    ib, tb, vb = FeatureBank(xi, yi, DEV), FeatureBank(xt, yt, DEV), FeatureBank(xv, yv, DEV)
    W0 = O.zero_shot_weights(xt, yt, C)
    res = {}
    for mode in ("single", "group"):
        ms, os_, ss, il, tl, vl = [], [], [], [], [], []
        for k, (lr, seed) in enumerate(((1e-3, 21), (1e-2, 22))):
            m = UMLClip(f"synthetic:{D}", C, logit_scale_init=3.0)
            m.precision = "fp32"
            m.load_state_dict({"head.weight": W0.clone()})
            m.to(DEV)
            o = build_optimizer(m.parameters(), "adamw", lr, 0.01)
            rng = torch.Generator().manual_seed(seed)
            ms.append(m); os_.append(o)
            ss.append(build_lr_scheduler(o, "cosine", 5, 80, warmup_type="linear", warmup_lr=1e-5))
            il.append(BankLoader(ib, 32, shuffle=True, rng=rng)); tl.append(BankLoader(tb, 32, shuffle=True, rng=rng))
            vl.append(BankLoader(vb, 32, shuffle=False, rng=rng))
        if mode == "group":
            trs = [{}, {}]
            outs = ft.train_group(ms, il, tl, vl, None, os_, ss, device=DEV, max_iters=80, alphas=[0.5, 0.5], eval_freq=10,
                                  patience=5, traces=trs)
        else:
            outs, trs = [], []
            for k in range(2):
                tr = {}
                outs.append(ft.train(ms[k], il[k], tl[k], vl[k], None, os_[k], ss[k], device=DEV, max_iters=80, alpha=0.5,
                                     eval_freq=10, patience=5, trace=tr))
                trs.append(tr)
        res[mode] = (outs, trs)
    for k in range(2):
        (o1, t1), (o2, t2) = (res["single"][0][k], res["single"][1][k]), (res["group"][0][k], res["group"][1][k])
        assert len(t1["stats"]) == len(t2["stats"])
        for a, b in zip(t1["img_idx"], t2["img_idx"]):
            assert torch.equal(a, b)
        for a, b in zip(t1["txt_idx"], t2["txt_idx"]):
            assert torch.equal(a, b)
        np.testing.assert_allclose([s["image_loss"] for s in t1["stats"]], [s["image_loss"] for s in t2["stats"]], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose([s["text_loss"] for s in t1["stats"]], [s["text_loss"] for s in t2["stats"]], rtol=1e-4, atol=1e-5)
        assert [e[0] for e in t1["evals"]] == [e[0] for e in t2["evals"]]
        np.testing.assert_allclose([e[2] for e in t1["evals"]], [e[2] for e in t2["evals"]], atol=1e-6)
        assert o1["iter"] == o2["iter"]
        a, b = o1["model"]["head.weight"], o2["model"]["head.weight"]
        assert float((a - b).abs().max() / a.abs().max()) < 1e-4


def test_batched_group_evaluation_matches_per_head_validate():
    """validate_group_enqueue (one logits + argmax launch and one reduction launch for all running heads of a group over
    the same bank) against validate_enqueue head by head: identical losses and hit counts (same kernel, same order)."""
    import torch
    from uml_b200 import finetune as ft, ops
    from uml_b200.engine.datasets.utils import BankLoader, FeatureBank

    class _Head:
        def __init__(self, w, s):
            self.head = type("H", (), {})()
            self.head.weight = torch.nn.Parameter(w, requires_grad=False)
            self.img_proj = None
            self.precision = "fp32"
            self._s = s

        def scales(self):
            return self._s, self._s

    class _Group:
        pass

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(9)
    n, D, C, K = 1234, 96, 37, 5
    bank = FeatureBank(torch.randn(n, D, generator=g), torch.randint(0, C, (n,), generator=g), dev)
    stride = (C * D + 3) // 4 * 4
    grp = _Group()
    grp.W = torch.randn(K, stride, generator=g).to(dev)
    grp.C, grp.D = C, D
    models = [_Head(grp.W[k, :C * D].view(C, D), 3.0 + k) for k in range(K)]
    loaders = [BankLoader(bank, 100, shuffle=False) for _ in range(K)]
    heads = [0, 2, 3]
    got = ft.validate_group_enqueue(grp, models, heads, lambda k: loaders[k])
    for (loss, hits, rows), k in zip(got, heads):
        l1, h1, n1 = ft.validate_enqueue(models[k], loaders[k])
        assert rows == n1 == n
        assert float(loss.item()) == float(l1.item())
        assert int(hits.item()) == int(h1.item())
