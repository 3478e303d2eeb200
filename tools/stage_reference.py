#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules the benchmark's reference arms import (build container only).

/root/reference does not exist on the GPU box, and the reference is a plain source tree (no setup.py / pyproject: nothing to
pip-install), so the four files its model code consists of are copied verbatim into the git-ignored, gpurun-shipped
``baseline/_ref/``:  models/generator.py, models/conformer.py, models/discriminator.py, utils.py.
bench.py imports them from there (``--impl reference`` on the host cores, ``gpu_eager_reference`` on the B200) and falls back
to the oracle port when the directory is absent.  Nothing under baseline/_ref is tracked, and the product never imports it.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/generator.py", "models/conformer.py", "models/discriminator.py", "utils.py"]


def stage() -> bool:
    if not os.path.isdir(SRC):
        return False
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), d)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write(SRC + "\n" + "\n".join(FILES) + "\n")
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference tree not present: nothing staged", file=sys.stderr)
