"""Build libnbdt_b200.so (hand-written sm_100a CUDA + the C ABI of include/nbdt_b200.h) in-tree.

    python neural-backed-decision-trees_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The library is written to
neural-backed-decision-trees_b200/lib/ (git-ignored; it travels to the GPU box with the snapshot).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libnbdt_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".h", ".cuh")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(INCLUDE, "nbdt_b200.h"), "rb").read())
    h.update(open(path, "rb").read())
    h.update(" ".join(FLAGS + ARCH).encode())
    return h.hexdigest()


def _generate():
    """Hierarchy-specialised tree kernels (csrc/tree_spec_gen.cu) from the shipped hierarchies."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_tree_kernels", os.path.join(CSRC, "gen_tree_kernels.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def build(force=False, verbose=False):
    _generate()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    objs, rebuilt = [], False
    for src in _sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest(path)
        if force or not os.path.exists(obj) or not os.path.exists(stamp) or open(stamp).read() != dig:
            cmd = [NVCC] + ARCH + FLAGS + ["-c", path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed for %s" % src)
            with open(obj + ".log", "w") as f:
                f.write(res.stdout + res.stderr)
            with open(stamp, "w") as f:
                f.write(dig)
            rebuilt = True
        objs.append(obj)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
