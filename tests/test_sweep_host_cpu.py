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
