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


def torch_validate_group_enqueue(group, models, heads, val_loader_of):
    """Stand-in of the batched evaluation launch: head by head through the oracle (same draw order per head)."""
    return [torch_validate_enqueue(models[k], val_loader_of(k)) for k in heads]


def _patched(monkeypatch):
    monkeypatch.setattr(sweep_mod, "HeadGroup", TorchGroup)
    monkeypatch.setattr(ft, "validate_enqueue", torch_validate_enqueue)
    monkeypatch.setattr(ft, "validate_group_enqueue", torch_validate_group_enqueue)


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


def test_group_heads_with_different_lengths_and_worker_protocol(monkeypatch):
    """Heads that end at different iterations (one in the middle of an evaluation interval), the num_workers > 0 sampler
    protocol (seed drawn when the iterator is built) and an evaluation interval that does not divide the epochs."""
    _patched(monkeypatch)
    run_group_case(ft, "cpu", max_iters=[60, 33, 47], num_workers=2, eval_freq=7)


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
    # combination 2 of the grid trained alone with its derived seed: the same weights, bit for bit (the group does not
    # couple its members)
    tmp2 = tmp_path / "alone"
    tmp2.mkdir()
    a2 = _sweep_args(tmp2)
    hp2 = dict(hp, lr=1e-3, weight_decay=0.0)
    ctx = ft._prepare(_sweep_datasets(), hp2, a2, rng=torch.Generator().manual_seed(ft.run_seed(a2, 2)))
    out = real([ctx["model"]], [ctx["image_loader"]], [ctx["text_loader"]], [ctx["val_loader"]], None, [ctx["optimizer"]],
               [ctx["scheduler"]], device="cpu", max_iters=hp2["max_iter"], alphas=a2.alpha, eval_freq=10, patience=hp2["patience"])
    alone = ft._finish(ctx, out[0], a2)
    grouped = torch.load([f for f in files if "lr_0.001-wd_0.0-" in f][0])
    assert grouped["iter"] == alone["iter"] and grouped["val_acc"] == alone["val_acc"]
    assert torch.equal(grouped["model"]["head.weight"], alone["model"]["head.weight"])


def test_sweep_alphas_trains_every_alpha_in_one_group(monkeypatch, tmp_path):
    """sweep_alphas: alpha x grid combinations in lock step, result files per alpha where one main() per alpha would have
    written them, and each alpha's results identical to its own batched sweep."""
    import os
    _patched(monkeypatch)
    monkeypatch.setattr(ft, "EVAL_FREQ", 10)
    hp = dict(optim="adamw", lr=[1e-2, 1e-3], weight_decay=0.0, lr_scheduler="cosine", batch_size=8, max_iter=30,
              warmup_iter=5, warmup_type="linear", warmup_min_lr=1e-5, dropout=0.0, learnable_temp=False, patience=3)
    sizes = []
    real = ft.train_group
    monkeypatch.setattr(ft, "train_group", lambda *a, **k: (sizes.append(len(a[0])), real(*a, **k))[1])
    per_alpha = []
    for al in (0.2, 1.0, 1.5):
        d = tmp_path / f"alpha_{al}"
        d.mkdir()
        per_alpha.append(_sweep_args(d, alpha=al))
    outs = ft.sweep_alphas(_sweep_datasets(), hp, per_alpha)
    assert sizes == [6] and len(outs) == 3
    for a, (res, best_val, best_test) in zip(per_alpha, outs):
        assert len(res["val_acc"]) == 2
        names = [f for _, _, fs in os.walk(a.savepath) for f in fs]
        assert names.count("test_result.pth") == 2 and "results.pth" in names
    # alpha = 1.0 on its own (fresh directory): same numbers and weights
    d = tmp_path / "single"
    d.mkdir()
    res1, _, _ = ft.sweep(_sweep_datasets(), hp, _sweep_args(d, alpha=1.0))
    assert res1["val_acc"] == outs[1][0]["val_acc"] and res1["test_acc"] == outs[1][0]["test_acc"]
    pick = lambda root: sorted(os.path.join(r, f) for r, _, fs in os.walk(str(root)) for f in fs if f == "test_result.pth")
    for fa, fb in zip(pick(per_alpha[1].savepath), pick(d)):
        assert torch.equal(torch.load(fa)["model"]["head.weight"], torch.load(fb)["model"]["head.weight"])


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


def test_cli_groups_alpha_jobs_only_when_batched(monkeypatch, tmp_path):
    """The YAML's alpha list becomes ONE job per remaining combination when sweep_batched is set, and stays one job per
    alpha (the reference's behaviour, finetune.py:541-560) otherwise."""
    import yaml
    calls = []
    monkeypatch.setattr(ft, "main", lambda args, alphas=None: calls.append((args.dataset, args.alpha, alphas)))
    cfg = {"dataset": ["sun397", "food101"], "alpha": [0.2, 0.5, 1.0], "modality": "crossmodal", "hyperparams": "clip_linear"}
    path = tmp_path / "sweep.yaml"
    path.write_text(yaml.safe_dump(cfg, sort_keys=False))
    ft.cli(["-c", str(path)])
    assert [(d, a) for d, a, _ in calls] == [(d, a) for d in ("sun397", "food101") for a in (0.2, 0.5, 1.0)]
    assert all(al is None for _, _, al in calls)
    calls.clear()
    path.write_text(yaml.safe_dump(dict(cfg, sweep_batched=True), sort_keys=False))
    ft.cli(["-c", str(path)])
    assert calls == [("sun397", 0.2, [0.2, 0.5, 1.0]), ("food101", 0.2, [0.2, 0.5, 1.0])]
