"""BASELINE config 5: the text:image conversion-ratio sweep of ``configs/ratio_sweep_sun397.yaml`` driven through the
public ``finetune.cli`` on synthetic banks with the real class counts (SUN397: 397, Food101: 101).

The YAML is the repo's own file with the lists cut down for test time (one train_shot, two text_shot values, a short
test preset) and ``sweep_batched`` switched on.  Checked:
  * the job list the YAML expands to runs end to end at C = 397 and C = 101 (ratio 0 = image only included) and writes its
    result files where ``collect_results.py`` of the reference looks for them;
  * ``text_shot`` selects min(text_shot, rows of the class) text rows per class (reference engine/datasets/utils.py:55-98);
  * every hyper-parameter combination of the crossmodal SUN397 points equals the CPU oracle's ``train`` run alone with
    the combination's seed (best iteration, validation accuracy, best weights <= 1e-3 of the tensor's norm)."""
import os

import numpy as np
import pytest
import torch
import yaml

from oracle import uml_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if torch.cuda.is_available():
    import uml_b200  # noqa: F401
    from uml_b200 import features as F, finetune as ft
    from uml_b200.engine.datasets.utils import TextTensorDataset
    from uml_b200.engine.optimizer.default import HYPER_DICT


def _banks(seed, C, D, shots, tpc, n_val_per_class, n_test):
    g = torch.Generator().manual_seed(seed)
    proto = torch.randn(C, D, generator=g)
    proto = proto / proto.norm(dim=1, keepdim=True)

    def rows(labels, noise):
        x = proto[labels] + noise * torch.randn(labels.numel(), D, generator=g) / D ** 0.5
        return x / x.norm(dim=1, keepdim=True)

    yi = torch.arange(C * shots) % C
    yt = torch.arange(C * tpc) % C
    yt = yt[torch.randperm(yt.numel(), generator=g)]          # text rows of a class are not contiguous in the file
    yv = torch.arange(C * n_val_per_class) % C
    yte = torch.randint(0, C, (n_test,), generator=g)
    return rows(yi, 1.5), yi, rows(yt, 1.0), yt, rows(yv, 1.5), yv, rows(yte, 1.5), yte


def test_ratio_yaml_runs_through_cli_and_matches_the_oracle(tmp_path):
    with open(os.path.join(ROOT, "configs", "ratio_sweep_sun397.yaml")) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    assert cfg["dataset"] == ["sun397", "food101"] and cfg["modality"] == ["crossmodal", "image"]
    assert cfg["text_shot"] == [1, 2, 4, 8, 16, 30] and cfg["train_shot"] == [1, 2, 4, 8, 16]
    fdir, rdir = str(tmp_path / "features"), str(tmp_path / "experiments")
    D, shots, tpc = 64, 4, 6
    HYPER_DICT["unit_test_ratio"] = dict(HYPER_DICT["clip_linear"], lr=[1e-2, 1e-3], weight_decay=[0.0], max_iter=[60], patience=[2])
    cfg.update(train_shot=[shots], text_shot=[2, 16], hyperparams=["unit_test_ratio"], num_workers=[0], feature_dir=[fdir],
               result_dir=[rdir], sweep_batched=[True], text_type=["cupl"])
    data = {}
    for name, C in (("sun397", 397), ("food101", 101)):
        xi, yi, xt, yt, xv, yv, xte, yte = _banks(C, C, D, shots, tpc, 2, 600)
        data[name] = (C, xi, yi, xt, yt, xv, yv)
        lab2cname = {c: f"class_{c}" for c in range(C)}
        F.write_text_bank(F.text_outdir(fdir, "ViT-B/16", name, "cupl"), xt, yt, lab2cname=lab2cname)
        F.write_image_bank(F.img_outdir(fdir, "ViT-B/16", name, "crop", shots, 1, "train"), train=(xi, yi), val=(xv, yv),
                           lab2cname=lab2cname)
        F.write_image_bank(F.img_outdir(fdir, "ViT-B/16", name, "crop", shots, 1, "test"), test=(xte, yte), lab2cname=lab2cname)
    # one of the two text banks also as a v2 file: the class-sorted row index then comes from the file
    F.convert_bank(F.text_outdir(fdir, "ViT-B/16", "sun397", "cupl"))
    ypath = str(tmp_path / "ratio.yaml")
    with open(ypath, "w") as f:
        yaml.dump(cfg, f)
    import unittest.mock as mock
    with mock.patch.object(ft, "EVAL_FREQ", 10):
        ft.cli(["-c", ypath])

    found = [os.path.join(r, f) for r, _, fs in os.walk(rdir) for f in fs]
    # datasets x (two crossmodal points + ONE image-only directory: its path does not depend on text_shot, so the second
    # image-only job finds the first one's results and skips, as in the reference)
    n_dirs = 2 * (2 + 1)
    assert sum(f.endswith("results.pth") for f in found) == n_dirs
    assert sum(f.endswith("test_result.pth") for f in found) == n_dirs * 2   # two combinations per directory
    assert sum(f.endswith("log.txt") for f in found) == n_dirs

    # the crossmodal SUN397 points against the oracle, combination by combination
    C, xi, yi, xt, yt, xv, yv = data["sun397"]
    scale = float(torch.tensor(4.60517).exp())
    args_like = type("A", (), {"seed": 1, "alpha": 1.0})()
    for text_shot in (2, 16):
        job_dir = ft.savedir(rdir, "sun397", "ViT-B/16", shots, 1, "cupl", text_shot, "crop", "crossmodal", "zeroshot", 1.0, 0, "")
        assert os.path.isfile(os.path.join(job_dir, "results.pth")), job_dir
        torch.manual_seed(1)   # main(): set_random_seed(args.seed), then the text selection is the first thing to draw
        tds = TextTensorDataset(xt, yt, torch.zeros(yt.numel(), dtype=torch.int64), n_shots=text_shot)
        assert tds.label_tensor.numel() == C * min(text_shot, tpc)
        W0 = O.zero_shot_weights(tds.input_tensor, tds.label_tensor, C)
        for n, (lr, wd) in enumerate([(1e-2, 0.0), (1e-3, 0.0)]):
            st = O.HeadState(head=W0.clone(), img_scale=scale, txt_scale=scale)
            torch.manual_seed(ft.run_seed(args_like, n))
            want, _ = O.train(st, (xi, yi), (tds.input_tensor, tds.label_tensor), (xv, yv), batch_size=32, optim="adamw", lr=lr,
                              weight_decay=wd, warmup_iter=50, sched_max_iter=60, max_iters=60, alpha=1.0, eval_freq=10,
                              patience=2, num_workers=0)
            got = torch.load(os.path.join(job_dir, ft.hparam_str("adamw", lr, wd, 32, 60, 0.0, False), "test_result.pth"),
                             map_location="cpu")
            assert got["iter"] == want["iter"], (text_shot, n, got["iter"], want["iter"])
            assert abs(got["val_acc"] - want["val_acc"]) < 1e-6
            a, b = got["model"]["head.weight"].numpy(), want["model"]["head.weight"].numpy()
            # 1e-3 relative for the tensor; single elements may be off by a few 1e-3 of the largest weight after 60 AdamW
            # steps at lr 1e-2 (the update m / sqrt(v) amplifies fp32 summation-order noise of near-zero gradients)
            assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-3
            assert np.abs(a - b).max() / np.abs(b).max() < 1e-2
