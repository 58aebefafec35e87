"""Stage the reference's own arch files under oracle/_ref/ (git-ignored, but shipped to the GPU box with the snapshot).

    python oracle/stage_ref.py          # build container only: needs /root/reference

The reference is pure Python, so "building" it is copying the four files of the hot path
(basicsr/models/archs/{FDN_arch,fdnlol24_arch,mar_arch,LPNet_arch}.py) unmodified next to the oracle; nothing under
oracle/_ref/ is tracked by git and nothing in the product package imports it.  With the files staged,
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference module itself on the box's host cores
(``kind: "reference"``) instead of the oracle port, and oracle/ref_loader.py finds the reference on the GPU box, where
/root/reference does not exist.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("FDN_REFERENCE_ROOT", "/root/reference"), "basicsr", "models", "archs")
DST = os.path.join(HERE, "_ref", "basicsr", "models", "archs")
FILES = ("FDN_arch.py", "fdnlol24_arch.py", "mar_arch.py", "LPNet_arch.py")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("reference not mounted at %s: nothing staged" % SRC)
        return False
    os.makedirs(DST, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        with open(os.path.join(DST, f), "rb") as fh:
            lines.append("%s  %s" % (hashlib.sha256(fh.read()).hexdigest(), f))
    with open(os.path.join(HERE, "_ref", "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print("staged %d reference files under %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
