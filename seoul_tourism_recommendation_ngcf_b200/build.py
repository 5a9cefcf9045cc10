"""In-tree nvcc build of libngcf_b200.so (sm_100a only).  `python -m seoul_tourism_recommendation_ngcf_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libngcf_b200.so")
SOURCES = ["csr_misc.cu", "spmm.cu", "dense.cu", "dense_tc.cu", "bpr.cu", "topk.cu", "topk_tc.cu", "adam.cu", "evalsample.cu"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the NGCF B200 library cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ngcf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every .cu under csrc/ for sm_100a into one C-ABI shared library, in-tree."""
    if not force and not _stale():
        return LIB
    cmd = [nvcc_path(), "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
