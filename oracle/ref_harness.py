"""TEST INFRASTRUCTURE — drives the *unmodified* reference on CPU.

Usable where the reference tree exists: ``/root/reference`` (the build container) or its byte-for-byte copy
``oracle/_ref`` made by ``oracle/build_ref.py`` (git-ignored; travels to the GPU box with the repo snapshot).
Nothing in the product imports this module.  It is the tool that *pins* ``oracle/uml_oracle.py``:
``tests/golden/make_golden.py`` calls it to produce the committed golden traces, and ``bench.py --impl reference``
times the reference's own ``finetune.train`` through it.

Recipe: the reference imports ``timm`` and ``ftfy`` which are not in
this image, so both are replaced by stub modules *before* the reference is imported;
``timm.models.create_model`` returns an identity backbone exposing ``num_features``
so that ``engine.models.head.UML`` (reference ``vision_language/engine/models/head.py:39-98``)
builds its own ``img_proj`` / ``head`` / scales and consumes pre-extracted feature rows.
``finetune.train`` / ``finetune.validate`` (``vision_language/finetune.py:120-315``) are
then called as shipped.
"""
from __future__ import annotations

import os
import sys
import types

def _default_root():
    if os.path.isfile("/root/reference/vision_language/finetune.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = os.environ.get("UML_REFERENCE_ROOT") or _default_root()
_VL = os.path.join(REFERENCE_ROOT, "vision_language")
_GAUSS = os.path.join(REFERENCE_ROOT, "Gaussian_experiment")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_VL, "finetune.py"))


class _IdentityBackbone:  # built lazily so torch is imported by the caller first
    pass


def _install_stubs():
    import torch

    os.environ.setdefault("WANDB_MODE", "disabled")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    # transformers probes timm.__spec__, so it must be imported before the stub exists
    import transformers  # noqa: F401
    import wandb  # noqa: F401

    if "timm" not in sys.modules or not hasattr(sys.modules["timm"], "_uml_stub"):

        class Identity(torch.nn.Module):
            """Stands in for a frozen timm backbone: features in, features out."""

            def __init__(self, dim):
                super().__init__()
                self.num_features = int(dim)
                self.num_classes = 0
                self.embed_dim = int(dim)

            def forward(self, x):
                return x

            def encode_image(self, x, **_):
                return x

        def create_model(name, **_kw):
            # names look like "ident:768"
            return Identity(int(str(name).split(":")[1]))

        timm = types.ModuleType("timm")
        timm._uml_stub = True
        timm.__spec__ = None
        models = types.ModuleType("timm.models")
        models.create_model = create_model
        timm.models = models
        timm.create_model = create_model
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["_uml_identity"] = types.ModuleType("_uml_identity")
        sys.modules["_uml_identity"].Identity = Identity
    for name in ("ftfy",):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt


_VL_MODS = None


def load_vision_language():
    """Import the reference's vision_language modules; returns a namespace."""
    global _VL_MODS
    if _VL_MODS is not None:
        return _VL_MODS
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if _VL not in sys.path:
        sys.path.insert(0, _VL)
    # guard against name clashes with modules of the same name from the Gaussian tree
    for clash in ("metrics", "utils", "model", "dataset", "data", "main"):
        sys.modules.pop(clash, None)
    import finetune  # noqa
    from engine.models import head  # noqa
    from engine.optimizer import optim, scheduler, default  # noqa
    from engine.datasets import utils as ds_utils  # noqa
    import features  # noqa

    ns = types.SimpleNamespace(
        finetune=finetune, head=head, optim=optim, scheduler=scheduler,
        default=default, ds_utils=ds_utils, features=features,
    )
    _VL_MODS = ns
    return ns


def load_gaussian():
    """Import the reference's Gaussian_experiment modules under private names."""
    import importlib.util

    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    out = {}
    saved = {k: sys.modules.get(k) for k in ("utils", "metrics", "model", "dataset", "data")}
    saved_path = list(sys.path)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        sys.path.insert(0, _GAUSS)
        for name in ("utils", "model", "dataset", "data"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(_GAUSS, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            out[name] = mod
    finally:
        sys.path[:] = saved_path
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return types.SimpleNamespace(**out)


# --------------------------------------------------------------------------------------
# Feature-row datasets in the shapes the reference loaders yield
# --------------------------------------------------------------------------------------

def make_image_rows_dataset(features, labels):
    """Dict-style dataset like the reference DatasetWrapper output
    (``engine/datasets/utils.py:162-174``) but over pre-extracted rows."""
    import torch

    class Rows(torch.utils.data.Dataset):
        def __len__(self):
            return features.shape[0]

        def __getitem__(self, i):
            return {"img": features[i], "label": labels[i]}

    return Rows()


def build_reference_model(kind, img_dim, text_indim, num_classes, learnable_temp=False,
                          logit=4.60517):
    """kind='uml'  -> reference UML (optional img_proj adapter when text_indim>0)
    kind='clip' -> reference UMLClip subclassed to add the extract_features method the
                   reference forgot (finetune.py:183 calls it; head.py:101-141 lacks it)."""
    import torch

    ns = load_vision_language()
    if kind == "uml":
        return ns.head.UML(f"ident:{img_dim}", text_indim, num_classes, bias=False,
                           learnable_temp=learnable_temp, freeze_backbone=True)
    Identity = sys.modules["_uml_identity"].Identity

    class _Clip:
        @staticmethod
        def load(name, jit=False):
            return Identity(img_dim), None

    orig = ns.head.clip
    ns.head.clip = _Clip
    try:
        class UMLClipX(ns.head.UMLClip):
            def extract_features(self, images):
                return self.vision_model.encode_image(images)

        m = UMLClipX("ident", num_classes, logit_scale_init=logit, bias=False,
                     learnable_temp=learnable_temp, freeze_backbone=True)
    finally:
        ns.head.clip = orig
    return m


class StepRecorder:
    """Optimizer/scheduler shim that records what the reference loop does each step
    without changing it: wraps optimizer.step to snapshot head.weight afterwards."""

    def __init__(self, model, optimizer, every=1):
        self.model = model
        self.opt = optimizer
        self.weights = []
        self.every = every
        self._n = 0
        orig = optimizer.step

        def step(*a, **k):
            r = orig(*a, **k)
            if self._n % self.every == 0:
                self.weights.append({k: v.detach().clone() for k, v in model.state_dict().items()})
            self._n += 1
            return r

        optimizer.step = step


class LossLogger:
    """Passed as ``logger=`` to reference train(); captures the per-step scalars the
    reference logs (finetune.py:236-244).  cka/mknn names are injected by the harness
    because the reference only defines them when capture_features_during_training."""

    def __init__(self):
        self.rows = []
        self.evals = []

    def log(self, d):
        if "train/image_loss" in d:
            self.rows.append(dict(d))
        else:
            self.evals.append(dict(d))
