"""Build recipe for the C oracle (test infrastructure).  Output: oracle/_build/liboracle_knn.so

The reference is pure Python, so there is nothing under /root/reference to compile into
oracle/_ref; the Python reference itself is imported in the dev container by oracle/make_golden.py
to generate tests/golden/*.npz.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "liboracle_knn.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "knn_oracle.c")
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force) and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(src):
        return SO
    # -ffp-contract=off: the restated fvec_L2sqr rounds the product and the add separately
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", SO, src, "-lm"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
