"""In-tree build of libmas_b200.so for sm_100a (nvcc cross-compiles without a GPU).
Every .cu is compiled to an object in parallel, then linked into one shared library."""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libmas_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _compile(args):
    nvcc, src, obj, verbose, extra = args
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True builds libmas_b200_trace.so with -DMAS_TRACE (device timestamps, debug switches, the single-CTA
    contraction) for the timeline tools under tools/; select it with MAS_LIB_PATH.  The product library has none
    of that code."""
    obj_dir, out = (OBJ + "_trace", OUT.replace(".so", "_trace.so")) if trace else (OBJ, OUT)
    extra = ["-DMAS_TRACE"] if trace else []
    srcs = sources()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "mas_b200.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(obj_dir, exist_ok=True)
    jobs, objs = [], []
    for src in srcs:
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)
        if stale:
            jobs.append((nvcc, src, obj, verbose, extra))
    if not jobs and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(o) for o in objs):
        return out
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        for src, rc, log in ex.map(_compile, jobs):
            if verbose or rc:
                sys.stderr.write(log)
            if rc:
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
    r = subprocess.run([nvcc, "-shared", "-o", out] + objs, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed linking {os.path.basename(out)}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
