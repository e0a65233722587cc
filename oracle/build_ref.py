"""TEST / BENCH INFRASTRUCTURE - makes the UNMODIFIED reference importable where /root/reference does not exist.

    python oracle/build_ref.py            # copy /root/reference/{vision_language,Gaussian_experiment}/**.py -> oracle/_ref/

The reference is pure Python (no build system): "building" it is copying its source files, byte for byte, into the
git-ignored directory ``oracle/_ref/`` - which is NOT gpurun-ignored, so it travels to the GPU box exactly like the
compiled ``lib/libuml_b200.so``.  Nothing is edited: the stubs for the two absent third-party imports (timm, ftfy;
SURVEY.md section 8c) live in ``oracle/ref_harness.py`` and are installed into ``sys.modules`` before the import.
Only ``bench.py --impl reference`` (the CPU arm) and the golden-vector scripts use the copy; the product never does.
``__graft_entry__.build()`` calls this when /root/reference is present.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

SRC = os.environ.get("UML_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
TREES = ("vision_language", "Gaussian_experiment")
# python sources + the small text assets the modules open at import time (the CLIP tokenizer's vocabulary)
KEEP = (".py", ".gz", ".yaml", ".yml")


def build(verbose: bool = False) -> str:
    if not os.path.isdir(os.path.join(SRC, "vision_language")):
        raise RuntimeError(f"reference tree not found at {SRC}")
    manifest = []
    for tree in TREES:
        for root, dirs, files in os.walk(os.path.join(SRC, tree)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__", "descriptions", "assets")]
            for f in sorted(files):
                if not f.endswith(KEEP):
                    continue
                s = os.path.join(root, f)
                rel = os.path.relpath(s, SRC)
                d = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest.append((rel, hashlib.sha256(open(d, "rb").read()).hexdigest()))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write(f"# byte-for-byte copies of {SRC} made by oracle/build_ref.py\n")
        for rel, h in sorted(manifest):
            fh.write(f"{h}  {rel}\n")
    if verbose:
        print(f"{len(manifest)} files -> {DST}")
    return DST


if __name__ == "__main__":
    build(verbose=True)
    sys.exit(0)
