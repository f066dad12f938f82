"""Compile (here, without a GPU) the tree-kernel cubins that `scripts/tune_tree.py <knobs...>` will ask for, so that the GPU box
finds them in lib/tree_cache/ instead of spending its minutes in NVRTC.
usage: python scripts/precompile_knobs.py 168,0,8,0,24,8,0,1 168,0,8,0,24,8,0,2 ..."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-backed-decision-trees_b200")
KNOBS = ["NBDT_B200_KA_REGCAP", "NBDT_B200_KA_MINB", "NBDT_B200_KA_UNROLL", "NBDT_B200_KA_PREFETCH", "NBDT_B200_KA_MAXLD", "NBDT_B200_KA_FENCE",
         "NBDT_B200_KA_WARPS", "NBDT_B200_KA_EXPMUFU", "NBDT_B200_KA_ROWS", "NBDT_B200_KA_TEAMS", "NBDT_B200_KA_LOCKSTEP"]
WANT = ("CIFAR100", "TinyImagenet200", "Imagenet1000")   # hierarchy="induced", as scripts/tune_tree.py asks

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import ctypes

    sys.path.insert(0, PKG)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import jit_host
    from nbdt._b200 import lib as L
    from nbdt.tree import Tree

    lib = L.load()
    desc, keep = jit_host.tree_desc(Tree(sys.argv[2], hierarchy="induced"))
    for dtype in (0, 1):
        rc = lib.nbdt_tree_jit_precompile(ctypes.byref(desc), 0, dtype)
        print(os.environ.get("KNOBS"), sys.argv[2], "dtype", dtype, "rc", rc, (lib.nbdt_last_error() or b"").decode()[:300] if rc else "", flush=True)
    sys.exit(0)

procs = []
for arg in sys.argv[1:]:
    if arg.startswith("--"):
        continue
    vals = arg.split(",")
    env = dict(os.environ, KNOBS=arg)
    for k, v in zip(KNOBS, vals):
        env[k] = v
    for name in WANT:
        if "--small" in sys.argv and name.startswith("Imagenet"):
            continue
        if os.environ.get("TUNE_CASES") and {"CIFAR100": "100", "TinyImagenet200": "200", "Imagenet1000": "1000"}[name] not in os.environ["TUNE_CASES"].split(","):
            continue
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--one", name], env=env))
for p in procs:
    p.wait()
