"""In-tree nvcc build of libngcf_b200.so (sm_100a only).  `python -m seoul_tourism_recommendation_ngcf_b200.build`.

Every .cu under csrc/ is compiled to an object file (in parallel) and linked into ONE C-ABI shared library next to the
package.  Objects are cached under _build/ keyed on a hash of the source, every header and the flags, so an unchanged
file is not recompiled; ``force=True`` (what ``__graft_entry__.build()`` passes) recompiles everything."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libngcf_b200.so")
FLAGS = ["-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a"]


def sources() -> list:
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the NGCF B200 library cannot be built")


def _headers_digest() -> bytes:
    h = hashlib.sha256()
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "ngcf_b200.h"))
    for d in deps:
        h.update(open(d, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.digest()


def _key(src: str, hdr: bytes) -> str:
    return hashlib.sha256(hdr + open(os.path.join(CSRC, src), "rb").read()).hexdigest()


def source_hash() -> str:
    """Hash of everything the library is built from (sources, headers, flags)."""
    hdr = _headers_digest()
    return hashlib.sha256("".join(_key(s, hdr) for s in sources()).encode()).hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every .cu under csrc/ for sm_100a into one C-ABI shared library, in-tree."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc, hdr = nvcc_path(), _headers_digest()
    stamp = os.path.join(OBJ, "libngcf_b200.hash")
    want = source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return LIB
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]

    def compile_one(src: str):
        obj, keyf = os.path.join(OBJ, src[:-3] + ".o"), os.path.join(OBJ, src[:-3] + ".key")
        key = _key(src, hdr)
        if not force and os.path.exists(obj) and os.path.exists(keyf) and open(keyf).read().strip() == key:
            return src, 0, ""
        cmd = [nvcc, "-c"] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + inc + [os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            open(keyf, "w").write(key)
        return src, r.returncode, " ".join(cmd) + "\n" + r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(len(sources()), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, sources()))
    for src, rc, out in results:
        if rc != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared"] + FLAGS + [os.path.join(OBJ, s[:-3] + ".o") for s in sources()] + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    open(stamp, "w").write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
