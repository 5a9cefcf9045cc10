"""Installs the UNMODIFIED reference modules under baseline/_ref/ (git-ignored; it travels to the GPU box like the
built .so files).

The reference (haesungpyun/seoul_tourism_recommendation_NGCF) is a flat directory of scripts without setup.py /
pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to install; the equivalent is a
byte-for-byte copy of the six modules of the path (model/{NGCF,bprloss,experiment,matrix,utils,parsers}.py).  They are
used, unmodified, (a) by `bench.py --impl reference` (the reference's own CPU torch.sparse path, kind "reference")
and (b) by tests/test_dropin_boundary.py, which runs the reference's Experiment over the drop-in NGCF / BPR.
Nothing under baseline/_ref is product code and none of it is committed."""
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/model"
DST = os.path.join(ROOT, "baseline", "_ref", "model")
FILES = ("NGCF.py", "bprloss.py", "experiment.py", "matrix.py", "utils.py", "parsers.py")


def install(verbose: bool = False) -> str | None:
    """Copies the reference modules when /root/reference exists (the build container); returns the install dir, or
    None if neither the reference nor an earlier install is present (e.g. a GPU box that received no baseline/_ref)."""
    if os.path.isdir(SRC):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        with open(os.path.join(DST, "SHA256SUMS"), "w") as out:
            for f in FILES:
                out.write(f"{hashlib.sha256(open(os.path.join(DST, f), 'rb').read()).hexdigest()}  {f}\n")
        if verbose:
            print(f"reference modules installed under {DST}")
    return DST if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


def ref_dir() -> str | None:
    return DST if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None


if __name__ == "__main__":
    print(install(verbose=True))
