"""Compile the plain-C restatement (oracle/tree_oracle.c) into oracle/_build/libtree_oracle.so.
TEST INFRASTRUCTURE ONLY (see the header of tree_oracle.c).  Called by __graft_entry__.build() and lazily by
oracle/tree_oracle_c.py."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tree_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libtree_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    base = ["gcc", "-O2", "-std=c99", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    res = subprocess.run(base[:1] + ["-fopenmp"] + base[1:], capture_output=True, text=True)
    if res.returncode != 0:  # no OpenMP runtime: scalar build
        res = subprocess.run(base, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed for oracle/tree_oracle.c:\n" + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
