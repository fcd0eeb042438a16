"""Recipe for oracle/_ref/: the reference's OWN implementation of the gated-attention pool, taken from the sources where
they lie under /root/reference at build time (never committed: oracle/_ref/ is git-ignored but travels to the GPU box
with the snapshot).  The reference is pure Python, so "building" it is copying the one file the CPU arm executes:

    /root/reference/model/dim1/ABMIL.py  ->  oracle/_ref/ABMIL.py      (unmodified; sha256 recorded next to it)

`bench.py --impl reference` and the `cpu_baseline` leg import that file by path and time `ABMIL.forward` + autograd on the
host cores (`cpu_baseline.kind = "reference"`); without it (no /root/reference at build time) they fall back to the
oracle's restatement (`kind = "port"`).  TEST / MEASUREMENT INFRASTRUCTURE ONLY — the product never imports it.
Run: python oracle/make_ref.py   (also called by __graft_entry__.build())."""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/model/dim1/ABMIL.py"
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "ABMIL.py")


def make():
    if not os.path.exists(SRC):
        return None
    os.makedirs(DST_DIR, exist_ok=True)
    shutil.copyfile(SRC, DST)
    digest = hashlib.sha256(open(DST, "rb").read()).hexdigest()
    with open(os.path.join(DST_DIR, "ABMIL.py.sha256"), "w") as f:
        f.write(f"{digest}  model/dim1/ABMIL.py (upstream, unmodified)\n")
    return DST


def load_reference_abmil():
    """The reference's ABMIL class from oracle/_ref/ABMIL.py, or None when the recipe has not run."""
    if not os.path.exists(DST):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("milb200_ref_abmil", DST)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ABMIL


if __name__ == "__main__":
    out = make()
    print(out if out else f"{SRC} not present: nothing copied (the CPU arm will use the oracle port)")
    sys.exit(0)
