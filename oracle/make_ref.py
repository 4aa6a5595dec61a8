"""Recipe for oracle/_ref/: a verbatim copy of the reference package, taken where it lies (/root/reference, or
$PYSVM_REF), so that `bench.py --impl reference` can time the UNMODIFIED reference on the GPU box's host cores
(that box has no /root/reference; oracle/_ref/ is git-ignored -- it never enters the history -- but travels with the
snapshot like the built .so files).  Test infrastructure: only bench.py's reference arm imports it.

    python oracle/make_ref.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def make_ref(src_root=None, quiet=False) -> bool:
    src_root = src_root or os.environ.get("PYSVM_REF", "/root/reference")
    src = os.path.join(src_root, "pysvm")
    dst = os.path.join(HERE, "_ref", "pysvm")
    if not os.path.isdir(src):
        if not quiet:
            print(f"{src} not found: oracle/_ref left as it is")
        return os.path.isdir(dst)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
    if not quiet:
        print(f"copied {src} -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
