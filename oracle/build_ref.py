"""Build recipe for oracle/_ref: the UNMODIFIED reference MAS kernel, compiled where it lies.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package (torch_tts_b200); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it.

What it does
------------
The reference ships its hot loop as Cython: /root/reference/vits2/monotonic_align/core.pyx
(built by vits2/monotonic_align/setup.py:1-9 with no OpenMP flag, so the
`prange` at core.pyx:41 is serial).  This script

  1. runs `cython` on /root/reference/vits2/monotonic_align/core.pyx with the
     generated C written to a temporary directory (never into this repo),
  2. compiles that C with gcc -O2 (the distutils default optimisation level the
     reference's setup.py gets), no -ffast-math, no -fopenmp,
  3. places ONLY the resulting binary at oracle/_ref/core<ext-suffix>.so.

oracle/_ref/ is git-ignored (binary artefact) but not gpurun-ignored, so the
compiled reference travels to the GPU box, where /root/reference does not exist.
No reference source is copied into the repository.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/vits2/monotonic_align/core.pyx"
OUT_DIR = os.path.join(HERE, "_ref")


def ref_so_path() -> str:
    return os.path.join(OUT_DIR, "core" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Compile the reference core.pyx into oracle/_ref/.  Returns the .so path,
    or None when /root/reference is absent (GPU box: the prebuilt file is used)."""
    so = ref_so_path()
    if not os.path.exists(REF_PYX):
        return so if os.path.exists(so) else None
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(REF_PYX):
        return so
    import numpy  # the reference's setup.py adds numpy's include dir

    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="mas_ref_build_") as tmp:
        c_file = os.path.join(tmp, "core.c")
        cmd = [sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file]
        subprocess.run(cmd, check=True, capture_output=not verbose)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        cmd = [
            cc, "-shared", "-fPIC", "-O2", "-fwrapv", "-fno-strict-aliasing",
            "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
            c_file, "-o", os.path.join(tmp, "core.so"),
        ]
        subprocess.run(cmd, check=True, capture_output=not verbose)
        shutil.copyfile(os.path.join(tmp, "core.so"), so)
    return so


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p if p else "reference sources absent and no prebuilt oracle/_ref")
