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


def _shipped_descs():
    """(name, C, child_ptr, child_node, root) of every shipped hierarchy, numbered like nbdt/_b200/tables.py."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_tree_kernels", os.path.join(CSRC, "gen_tree_kernels.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    nbdt = os.path.join(HERE, "nbdt")
    out, seen = [], set()
    for dataset in sorted(os.listdir(os.path.join(nbdt, "hierarchies"))):
        for fname in sorted(os.listdir(os.path.join(nbdt, "hierarchies", dataset))):
            if not fname.endswith(".json"):
                continue
            pw = os.path.join(nbdt, "wnids", dataset + ".txt")
            wnids = [l.strip() for l in open(pw)] if os.path.exists(pw) else None
            C, children, root = mod.flatten(os.path.join(nbdt, "hierarchies", dataset, fname), wnids)
            h = mod.fnv1a(C, children, root)
            if h in seen:
                continue
            seen.add(h)
            ptr = [0]
            for ch in children:
                ptr.append(ptr[-1] + len(ch))
            out.append((f"{dataset}/{fname}", C, ptr, [c for ch in children for c in ch], root))
    return out


def _precompile_one(job):
    import ctypes

    name, C, ptr, nodes, root, mode, dtype = job
    lib = ctypes.CDLL(LIB)

    class Desc(ctypes.Structure):
        _fields_ = [("num_classes", ctypes.c_int32), ("num_inodes", ctypes.c_int32), ("root", ctypes.c_int32),
                    ("child_ptr", ctypes.POINTER(ctypes.c_int32)), ("child_node", ctypes.POINTER(ctypes.c_int32))]

    a = (ctypes.c_int32 * len(ptr))(*ptr)
    b = (ctypes.c_int32 * len(nodes))(*nodes)
    d = Desc(C, len(ptr) - 1, root, a, b)
    lib.nbdt_tree_jit_precompile.argtypes = [ctypes.POINTER(Desc), ctypes.c_int, ctypes.c_int]
    rc = lib.nbdt_tree_jit_precompile(ctypes.byref(d), mode, dtype)
    lib.nbdt_last_error.restype = ctypes.c_char_p
    return name, mode, dtype, rc, (lib.nbdt_last_error() or b"").decode() if rc else ""


def _precompile_jobs():
    return [(name, C, ptr, nodes, root, mode, dtype) for name, C, ptr, nodes, root in _shipped_descs() if 2 <= C <= 1536
            for mode in (0, 1, 2) for dtype in (0, 1, 2)]


def _precompile_shard(index, count, verbose=False):
    """Compile jobs[index::count] in this process; the number that failed is the exit status of the worker."""
    failed = 0
    for job in _precompile_jobs()[index::count]:
        name, mode, dtype, rc, err = _precompile_one(job)
        if rc != 0:
            failed += 1
            sys.stderr.write(f"tree JIT precompile failed for {name} mode {mode} dtype {dtype}: {err}\n")
        elif verbose:
            sys.stderr.write(f"tree JIT precompiled {name} mode {mode} dtype {dtype}\n")
    return failed


def precompile_tree_kernels(verbose=False):
    """Populate lib/tree_cache/ with the run-time-specialised tree kernels (csrc/tree_jit.cu) of every shipped hierarchy with
    at most 1536 leaves: soft loss, hard loss, soft forward x fp32, bf16, fp16.  NVRTC needs no GPU.  Any other hierarchy is
    compiled on first use at run time; this only moves the few seconds per kernel to build time.  The work is sharded over
    worker processes started from this file by path, so it does not matter how the caller imported it."""
    total = len(_precompile_jobs())
    count = max(1, min(8, os.cpu_count() or 1))
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--precompile-shard", str(i), str(count)]
                              + (["--verbose"] if verbose else [])) for i in range(count)]
    failed = 0
    for p in procs:
        rc = p.wait()
        failed += rc if 0 <= rc < 250 else total // count  # a crashed worker counts as its whole shard
    if failed:
        sys.stderr.write(f"{failed} of {total} tree kernels were not precompiled (they compile at first use instead)\n")
    return total - failed


def build(force=False, verbose=False, precompile=True):
    lib = _build_lib(force, verbose)
    if precompile:
        precompile_tree_kernels(verbose)
    return lib


def _build_lib(force=False, verbose=False):
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
    if "--precompile-shard" in sys.argv:
        at = sys.argv.index("--precompile-shard")
        sys.exit(min(249, _precompile_shard(int(sys.argv[at + 1]), int(sys.argv[at + 2]), "--verbose" in sys.argv)))
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, precompile="--no-precompile" not in sys.argv))
