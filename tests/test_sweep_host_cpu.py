"""Host logic of sweep batching (finetune.train_group) without a GPU: the four kernels of uml_sweep_run and the eval
kernels are replaced by a plain torch fp32 stand-in built on the oracle's step, so that everything around them - the
per-head sampler streams, chunking at epoch ends and evaluation points, lr schedules, stats log, early stopping of
single heads, best-state bookkeeping - is checked against the reference's golden trace and the oracle here on the CPU.
The real kernels run the same body in tests/test_sweep_gpu.py."""
import torch

import uml_b200  # noqa: F401
from oracle import uml_oracle as O
from uml_b200 import finetune as ft
from uml_b200.engine import sweep as sweep_mod

from sweep_case import run_group_case


class TorchGroup(sweep_mod.HeadGroup):
    """HeadGroup with the launches replaced by torch ops (test double; the product class refuses CPU tensors)."""

    def __init__(self, models, optimizers, image_bank, text_bank, max_img_rows, max_txt_rows, device, log_slots=128):
        self.K = len(models)
        self.models, self.opts = list(models), list(optimizers)
        self.C, self.D = models[0].num_classes, models[0].shared_dim
        self.banks = (image_bank, text_bank)
        self.log_slots = log_slots
        self.stats_log = torch.zeros((log_slots, self.K, 2, 4))
        self.W = torch.stack([m.head.weight.data.reshape(-1) for m in models])
        for k, m in enumerate(models):
            m.head.weight.data = self.W[k].view(self.C, self.D)
        o0 = optimizers[0]
        self.oracle_opts = [O.OracleOptimizer({"head.weight": m.head.weight.data}, o.name, o.defaults["lr"],
                                              o.group_of(m.head.weight)["weight_decay"])
                            for m, o in zip(models, optimizers)]
        self.scales = tuple(float(s) for s in models[0].scales())
        assert o0.name in ("adamw", "adam", "sgd")

    def run(self, perms_img, perms_txt, pos_img, pos_txt, rows, lrs, alphas, active, slot0):
        s0 = slot0 % self.log_slots
        assert s0 + len(rows) <= self.log_slots
        ints = self.stats_log.view(torch.int32)
        for i, (n_i, n_t) in enumerate(rows):
            for k in range(self.K):
                if not active[k]:
                    continue
                st = O.HeadState(head=self.models[k].head.weight.data, img_scale=self.scales[0], txt_scale=self.scales[1])
                xi = yi = xt = yt = None
                if n_i:
                    idx = perms_img[k][pos_img:pos_img + n_i]
                    xi, yi = self.banks[0].features[idx], self.banks[0].labels[idx]
                if n_t:
                    idx = perms_txt[k][pos_txt:pos_txt + n_t]
                    xt, yt = self.banks[1].features[idx], self.banks[1].labels[idx]
                stats, grads = O.uml_step_grads(st, xi, yi, xt, yt, alphas[k])
                self.oracle_opts[k].step(grads, lrs[i][k])
                self.opts[k].slot(self.models[k].head.weight)["step"] += 1
                for s, (l, a, n) in enumerate((("image_loss", "img_acc", n_i), ("text_loss", "text_acc", n_t))):
                    self.stats_log[s0 + i, k, s, 0] = stats[l] if n else 0.0
                    ints[s0 + i, k, s, 2] = int(round(stats[a] * n)) if n else 0
                    ints[s0 + i, k, s, 3] = n
            pos_img += n_i
            pos_txt += n_t


def torch_validate_enqueue(model, val_loader):
    iter(val_loader)
    bank = val_loader.bank
    st = O.HeadState(head=model.head.weight.data, img_scale=float(model.scales()[0]))
    loss, acc = O.validate(st, bank.features, bank.labels, val_loader.batch_size, loader_protocol=False)
    return torch.tensor([loss]), torch.tensor([int(round(acc * len(bank)))], dtype=torch.int32), len(bank)


def _patched(monkeypatch):
    monkeypatch.setattr(sweep_mod, "HeadGroup", TorchGroup)
    monkeypatch.setattr(ft, "validate_enqueue", torch_validate_enqueue)


def test_group_of_three_heads_matches_reference_and_oracle(monkeypatch):
    _patched(monkeypatch)
    outs, traces, stopped = run_group_case(ft, "cpu")
    assert stopped[2] and not stopped[0], "the case must cover a head that stops early while others continue"


def test_group_sgd_image_only(monkeypatch):
    _patched(monkeypatch)
    run_group_case(ft, "cpu", optim="sgd", modality="image")


def test_group_text_only(monkeypatch):
    _patched(monkeypatch)
    run_group_case(ft, "cpu", optim="adam", modality="text")


def test_group_rejects_mixed_runs():
    from uml_b200.engine.datasets.utils import BankLoader, FeatureBank
    from uml_b200.engine.models.head import UML, UMLClip
    from uml_b200.engine.optimizer.optim import build_optimizer
    bank = FeatureBank(torch.zeros(10, 8), torch.zeros(10, dtype=torch.int64), "cpu")
    a, b = UMLClip("synthetic:8", 4), UML("synthetic:8", 8, 4)
    oa, ob = build_optimizer(a.parameters(), "adamw", 1e-3, 0.0), build_optimizer(b.parameters(), "adamw", 1e-3, 0.0)
    la = [BankLoader(bank, 4, shuffle=True), BankLoader(bank, 4, shuffle=True)]
    assert sweep_mod.group_blockers([a, b], [oa, ob], la, [None, None])  # adapter variant
    c = UMLClip("synthetic:8", 4)
    oc = build_optimizer(c.parameters(), "sgd", 1e-3, 0.0)
    assert sweep_mod.group_blockers([a, c], [oa, oc], la, [None, None])  # optimizer kinds differ
    od = build_optimizer(c.parameters(), "adamw", 1e-2, 0.01)
    assert sweep_mod.group_blockers([a, c], [oa, od], la, [None, None]) == []  # lr / wd may differ
    assert sweep_mod.group_blockers([a, c], [oa, od], [la[0], BankLoader(bank, 8, shuffle=True)], [None, None])


def _sweep_args(tmp_path, **over):
    import types
    a = types.SimpleNamespace(device="cpu", savepath=str(tmp_path), use_clip=True, clip_encoder="synthetic", img_indim=16,
                              nclasses=6, logit=3.0, hyperparams="clip_linear", modality="crossmodal", classifier_init="zeroshot",
                              common_dim=0, text_indim=16, num_workers=0, alpha=0.5, eval_test=False, seed=3, overwrite=False,
                              sweep_batched=True, dataset="synth", precision="fp32", vision_model="", dp_sampler="global")
    a.__dict__.update(over)
    return a


def _sweep_datasets():
    from oracle.synth import synth_banks
    from uml_b200.engine.datasets.utils import FeatureBank, TextTensorDataset
    xi, yi, xt, yt, xv, yv = synth_banks(2, 6, 16, 16, 60, 4, 30)
    tds = TextTensorDataset(xt, yt, torch.zeros(len(yt), dtype=torch.int64))
    return {"text_ds": tds, "text_bank": FeatureBank.from_text_dataset(tds, "cpu"), "img_tr_bank": FeatureBank(xi, yi, "cpu"),
            "img_val_bank": FeatureBank(xv, yv, "cpu"), "img_te_bank": FeatureBank(xv, yv, "cpu")}


def test_batched_sweep_orchestration(monkeypatch, tmp_path):
    """finetune.sweep with --sweep-batched (setup_group): all four combinations of a 2 x 2 grid train as one group, write
    their result files, and a second call finds them; the results are those of each combination trained alone with
    its derived seed."""
    import os
    _patched(monkeypatch)
    hp = dict(optim="adamw", lr=[1e-2, 1e-3], weight_decay=[0.0, 0.01], lr_scheduler="cosine", batch_size=8, max_iter=40,
              warmup_iter=5, warmup_type="linear", warmup_min_lr=1e-5, dropout=0.0, learnable_temp=False, patience=3)
    calls = []
    real = ft.train_group
    monkeypatch.setattr(ft, "train_group", lambda *a, **k: (calls.append(len(a[0])), real(*a, **k))[1])
    monkeypatch.setattr(ft, "EVAL_FREQ", 10)
    res, best_val, best_test = ft.sweep(_sweep_datasets(), hp, _sweep_args(tmp_path))
    assert calls == [4] and len(res["val_acc"]) == 4 and best_val > 1.0 / 6
    files = [os.path.join(r, f) for r, _, fs in os.walk(str(tmp_path)) for f in fs if f == "test_result.pth"]
    assert len(files) == 4
    # second call: nothing left to train
    res2, _, _ = ft.sweep(_sweep_datasets(), hp, _sweep_args(tmp_path))
    assert calls == [4] and res2["val_acc"] == res["val_acc"]
    # combination n alone, same derived seed -> same result (the group does not couple its members)
    tmp2 = tmp_path / "alone"
    tmp2.mkdir()
    hp1 = dict(hp, lr=[1e-3], weight_decay=[0.0])  # combination index 2 of the grid above
    args = _sweep_args(tmp2, seed=3)
    import uml_b200.finetune as F2
    monkeypatch.setattr(F2, "setup_group", lambda ds, combos, a: real_setup_group_with_offset(ds, combos, a, 2))

    def real_setup_group_with_offset(ds, combos, a, offset):
        ctx = F2._prepare(ds, combos[0], a, rng=torch.Generator().manual_seed(a.seed * 1000003 + offset))
        out = real([ctx["model"]], [ctx["image_loader"]], [ctx["text_loader"]], [ctx["val_loader"]], None, [ctx["optimizer"]],
                   [ctx["scheduler"]], device=a.device, max_iters=combos[0]["max_iter"], alphas=a.alpha, eval_freq=10,
                   patience=combos[0]["patience"])
        return [F2._finish(ctx, out[0], a)]
    res3, _, _ = ft.sweep(_sweep_datasets(), hp1, args)
    assert res3["val_acc"][0] == res["val_acc"][2] and res3["test_acc"][0] == res["test_acc"][2]


def test_setup_group_routes_unbatchable_runs_to_the_plain_loop(monkeypatch, tmp_path):
    """Learnable-temperature combinations cannot share a HeadGroup: they go through train() one by one, the others
    still train together."""
    _patched(monkeypatch)
    hp = dict(optim="adamw", lr=[1e-2, 1e-3], weight_decay=0.0, lr_scheduler="cosine", batch_size=8, max_iter=20,
              warmup_iter=5, warmup_type="linear", warmup_min_lr=1e-5, dropout=0.0, learnable_temp=[False, True], patience=3)
    groups, solos = [], []
    real = ft.train_group
    monkeypatch.setattr(ft, "train_group", lambda *a, **k: (groups.append(len(a[0])), real(*a, **k))[1])

    def fake_train(model, *a, **k):
        solos.append(bool(model.learnable_temp))
        return {"iter": 0, "val_acc": 0.5, "val_loss": 1.0, "model": {k_: v.cpu() for k_, v in model.state_dict().items()}}
    monkeypatch.setattr(ft, "train", fake_train)
    monkeypatch.setattr(ft, "EVAL_FREQ", 10)
    # image-only UML without adapter (common_dim 0): the head is batchable unless its temperatures are learnable
    args = _sweep_args(tmp_path, use_clip=False, vision_model="synthetic", common_dim=0, modality="image",
                       classifier_init="random")
    res, _, _ = ft.sweep(_sweep_datasets(), hp, args)
    assert groups == [2] and solos == [True, True] and len(res["val_acc"]) == 4
