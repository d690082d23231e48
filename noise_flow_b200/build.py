"""In-tree build of ``libnoiseflow_b200.so`` (hand-written sm_100a CUDA + C-ABI) with nvcc.

``python -m noise_flow_b200.build`` or ``noise_flow_b200.build.build()``.  The shared library is placed
next to this file so it travels with the repository snapshot to the GPU box; nothing is written outside
the repository.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libnoiseflow_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# -use_fast_math only for the device code: ex2/rcp/lg2/rsq approximations are part of the kernel design
# (DESIGN.md "transcendentals"); the host-side folding in nf_api.cu stays IEEE double.
UNITS = [
    ("nf_kernels.cu", ["-use_fast_math"]),
    ("nf_stream.cu", ["-use_fast_math"]),
    ("nf_tc.cu", ["-use_fast_math"]),
    ("nf_hybrid.cu", ["-use_fast_math"]),
    ("nf_wino.cu", ["-use_fast_math"]),
    ("nf_wide.cu", ["-use_fast_math"]),
    ("nf_wide_cond.cu", ["-use_fast_math"]),
    ("nf_wide_tc.cu", ["-use_fast_math"]),
    ("nf_wide_tcs.cu", ["-use_fast_math"]),
    ("nf_train.cu", []),
    ("nf_train_wide.cu", []),
    ("nf_trainer.cu", []),
    ("nf_api.cu", []),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".h", ".cuh"))]
    out.append(os.path.join(ROOT, "include", "noiseflow_b200.h"))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit for sm_100a and link the shared library. Returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        deps_newer = (not os.path.exists(obj)) or force or any(
            os.path.getmtime(s) > os.path.getmtime(obj) for s in _sources() if s.endswith((".h", ".cuh")) or s == path)
        if not deps_newer:
            continue
        cmd = [nvcc] + ARCH + COMMON + extra + os.environ.get("NF_EXTRA_NVCC_FLAGS", "").split() + \
            (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]      # NF_EXTRA_NVCC_FLAGS: experiment switches (-DNF_...)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: %s" % " ".join(cmd))
    cmd = [nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed: %s" % " ".join(cmd))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
