"""Copy the UNMODIFIED reference sources into baseline/_ref/ (git-ignored; it travels to the GPU box with gpurun).

The reference is an un-packaged tree of Python scripts (no setup.py / pyproject), so `pip install --target baseline/_ref
/root/reference` has nothing to build; a plain copy of its *.py files is the equivalent install.  bench.py's
`--impl reference` arm and `cpu_baseline` leg import `features` from there when it exists (kind "reference"), and fall
back to the oracle port (kind "port") when it does not.  Test infrastructure: nothing under the product package reads it.

    python oracle/install_reference.py        # run in the CPU container, where /root/reference exists
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("DSP_REF_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install():
    if not os.path.isdir(os.path.join(SRC, "features")):
        return None
    os.makedirs(DST, exist_ok=True)
    n = 0
    for base, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if not d.startswith(".")]
        rel = os.path.relpath(base, SRC)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(base, f), os.path.join(DST, rel, f))
                n += 1
    return DST, n


if __name__ == "__main__":
    r = install()
    print("reference not found at " + SRC if r is None else "installed %d files into %s" % (r[1], r[0]))
    sys.exit(0)
