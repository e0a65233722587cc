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


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof and every field offset of the ctypes mirrors equal what a C compiler makes of include/uml_b200.h."""
    import ctypes
    import shutil
    import subprocess
    pairs = {"uml_segment": _lib.Segment, "uml_seg_stats": _lib.SegStats, "uml_update": _lib.Update,
             "uml_tc_segments": _lib.TcSegments, "uml_linear_step_args": _lib.LinearStepArgs,
             "uml_run_step": _lib.RunStep, "uml_sweep_args": _lib.SweepArgs}
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "uml_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_sweep_run_validates_arguments_before_touching_the_device():
    import ctypes as C
    lib = _lib.load()
    a = _lib.SweepArgs()
    rows, lr = (C.c_int64 * 2)(4, 4), (C.c_float * 1)(1e-3)
    a.n_heads = 0
    assert lib.uml_sweep_run(C.byref(a), 1, rows, lr, None) != 0
    assert b"heads" in lib.uml_last_error()
    a.n_heads, a.dim, a.n_classes, a.ldg, a.kind = 1, 8, 4, 4, 7
    assert lib.uml_sweep_run(C.byref(a), 1, rows, lr, None) != 0
    assert b"optimizer kind" in lib.uml_last_error()
    a.kind = 1
    assert lib.uml_sweep_run(C.byref(a), 1, rows, lr, None) != 0  # null buffers
    assert b"null buffer" in lib.uml_last_error()
