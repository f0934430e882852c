"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.

Packs the UNMODIFIED reference package (/root/reference/neural_lam/*.py, read where it lies)
into the git-ignored archive oracle/_ref/neural_lam_ref.zip so that the reference's own CPU
path can be timed on the GPU box's host cores next to the CUDA path (`bench.py --impl
reference`, `cpu_baseline.kind = "reference"`): /root/reference does not exist there, the
archive travels with the repo snapshot like the built .so files do.  Nothing is copied into
git history (oracle/_ref/ is in .gitignore); the archive is imported with zipimport behind
oracle/ref_stubs.py, byte for byte the reference's code.

    python oracle/stage_ref.py        # (re)build the archive; called by __graft_entry__.build()
"""
import hashlib
import os
import sys
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST_DIR = os.path.join(ROOT, "oracle", "_ref")
DST = os.path.join(DST_DIR, "neural_lam_ref.zip")


def _files():
    out = []
    pkg = os.path.join(SRC, "neural_lam")
    for d, _, names in sorted(os.walk(pkg)):
        for n in sorted(names):
            if n.endswith(".py"):
                p = os.path.join(d, n)
                out.append((p, os.path.relpath(p, SRC)))
    return out


def digest(files):
    h = hashlib.sha256()
    for p, rel in files:
        h.update(rel.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def stage(force=False):
    """Returns the archive path, or None when /root/reference is absent (GPU box: the
    prebuilt archive, if any, is used as it is)."""
    if not os.path.isdir(os.path.join(SRC, "neural_lam")):
        return DST if os.path.exists(DST) else None
    files = _files()
    dg = digest(files)
    stamp = os.path.join(DST_DIR, "neural_lam_ref.sha256")
    if not force and os.path.exists(DST) and os.path.exists(stamp) and open(stamp).read().strip() == dg:
        return DST
    os.makedirs(DST_DIR, exist_ok=True)
    tmp = DST + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for p, rel in files:
            z.write(p, rel)
    os.replace(tmp, DST)
    with open(stamp, "w") as f:
        f.write(dg + "\n")
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
