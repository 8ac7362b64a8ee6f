"""Stage the UNMODIFIED reference files of the hot path into ``oracle/_ref/`` so that they travel to the GPU box.

    python -m oracle.stage_reference            (also run by __graft_entry__.build() when /root/reference is present)

TEST / BASELINE INFRASTRUCTURE.  The reference is pure Python without a ``setup.py`` / ``pyproject.toml`` (a
``pip install --target baseline/_ref /root/reference`` therefore has nothing to install), and ``/root/reference`` does not exist
on the GPU box.  This recipe is the Python counterpart of compiling a C reference into ``oracle/_ref/``: it copies, byte for
byte, the package directories ``nets/Achelous.py:49-53`` imports (``nets/ backbone/ neck/ head/``) and
``utils/utils_bbox.py`` into ``oracle/_ref/`` - a git-ignored build artefact that is never committed and never imported by the
product - so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's OWN forward on the GPU box's
host cores (``kind: "reference"``), and ``tests/test_oracle_vs_reference.py`` can pin the oracle against it there as well.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ACHELOUS_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("nets", "backbone", "neck", "head")
FILES = ("utils/__init__.py", "utils/utils_bbox.py")


def stage(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "nets", "Achelous.py")):
        if verbose:
            print(f"[stage_reference] {SRC} not present: nothing staged (the GPU box uses the copy that travelled)")
        return False
    os.makedirs(DST, exist_ok=True)
    for pkg in PACKAGES:
        dst = os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, pkg), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.onnx"))
    for rel in FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
    n = 0
    for root, dirs, files in os.walk(DST):      # byte-for-byte: the staged copy IS the reference
        dirs[:] = [d for d in dirs if d != "__pycache__"]     # (left behind by importing the staged copy)
        for f in files:
            if f == "STAGED_FROM" or f.endswith(".pyc"):      # this recipe's own marker from an earlier run
                continue
            rel = os.path.relpath(os.path.join(root, f), DST)
            assert filecmp.cmp(os.path.join(DST, rel), os.path.join(SRC, rel), shallow=False), rel
            n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{SRC}\n{n} files, unmodified (oracle/stage_reference.py)\n")
    if verbose:
        print(f"[stage_reference] {n} reference files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
