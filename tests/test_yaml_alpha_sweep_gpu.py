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
