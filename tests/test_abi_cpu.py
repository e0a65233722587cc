"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol
that include/uml_b200.h declares, and the product path refuses to run without CUDA."""
import os
import re

import pytest
import torch

import uml_b200
from uml_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "uml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uml_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in uml_b200.h but not exported"
    assert lib.uml_abi_version() == 1
    # every exported entry point has a ctypes prototype (or is the error getter)
    assert set(names) - {"uml_last_error"} == set(_lib.PROTOTYPES)


def test_no_cpu_fallback():
    from uml_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.gather_rows(torch.zeros(4, 8), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.adamw_step(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), lr=1e-3, step=1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "unpaired-multimodal-learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", text, flags=re.S).replace("# oracle", ""), f
