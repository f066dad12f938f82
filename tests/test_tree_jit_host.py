"""CPU checks of the run-time tree-kernel generator (csrc/tree_jit.cu): the CUDA source it emits for a hierarchy compiles, with
-DKA_HOST, as plain C++ (tests/jit_host.py) and must reproduce the pinned oracle -- soft probabilities, both losses and their
gradients -- for binary and k-ary, shipped and random hierarchies, every dtype, single-warp kernels and warp teams (the team's
warps run as host threads that meet at real barriers).  No GPU: what the GPU tests add is the device arithmetic (approximate
ex2 / rcp), the launch geometry and the cubin cache.
"""
import os

import numpy as np
import pytest

import jit_host as J
from conftest import write_random_hierarchy
from oracle import tree_oracle as O

SOFT, HARD, FWD = 0, 1, 2


def vec_for(C, dtype):
    b = C * (4 if dtype == 0 else 2)
    for v in (16, 8, 4, 2):
        if b % v == 0 and v >= (4 if dtype == 0 else 2):
            return v
    raise AssertionError


def check(tree, ot, C, dtype=0, rows=24, seed=0, warps=None, hw=0):
    """hw > 0: the class-major variant (vec = 0), the rows being hw pixels of rows / hw images."""
    env = os.environ.pop("NBDT_B200_KA_WARPS", None)
    if warps:
        os.environ["NBDT_B200_KA_WARPS"] = str(warps)
    try:
        rng = np.random.default_rng(seed)
        x = (3 * rng.standard_normal((rows, C))).astype(np.float32)
        y = rng.integers(0, C, rows)
        vec = 0 if hw else vec_for(C, dtype)
        _, xr = J.to_dtype_bytes(x, dtype)
        xd = xr.astype(np.float64)
        lp = dtype != 0
        cp = J.run_soft_forward(tree, dtype, vec, x, hw=hw)
        np.testing.assert_allclose(cp, O.soft_probs(ot, xd), rtol=5e-6, atol=1e-12)
        n_in = len(ot.inodes)
        for mode, fn, wt in ((SOFT, O.soft_tree_sup_loss, 1.3), (HARD, O.hard_tree_sup_loss, 1.3 * 2.0 * 2.0 / n_in)):
            kw = {"tree_supervision_weight": 2.0} if mode == HARD else {}
            want_l, want_g = fn(ot, xd, y, xent_weight=0.7, tree_weight=1.3, **kw)
            loss, grad = J.run_loss(tree, mode, dtype, vec, x, y, 0.7, wt, hw=hw)
            assert abs(loss - want_l) <= 2e-6 * abs(want_l)
            scale = np.abs(want_g).max()
            np.testing.assert_allclose(grad, want_g, rtol=1e-2 if lp else 2e-5, atol=(5e-3 if lp else 1e-6) * scale)
            # forward only
            loss2, _ = J.run_loss(tree, mode, dtype, vec, x, y, 0.7, wt, want_grad=False, hw=hw)
            assert abs(loss2 - want_l) <= 2e-6 * abs(want_l)
    finally:
        os.environ.pop("NBDT_B200_KA_WARPS", None)
        if env is not None:
            os.environ["NBDT_B200_KA_WARPS"] = env


def shipped(dataset, hierarchy):
    from nbdt.tree import Tree
    from nbdt.utils import dataset_to_default_path_wnids, hierarchy_to_path_graph

    tree = Tree(dataset, hierarchy=hierarchy)
    ot = O.OracleTree(hierarchy_to_path_graph(dataset, hierarchy), dataset_to_default_path_wnids(dataset))
    return tree, ot, len(tree.wnids_leaves)


@pytest.mark.parametrize("dataset,hierarchy,dtype,warps", [
    ("CIFAR10", "induced", 0, None), ("CIFAR10", "wordnet", 1, None), ("CIFAR10", "induced", 0, 2),
    ("CIFAR100", "induced", 0, None), ("CIFAR100", "wordnet", 2, None), ("CIFAR100", "wordnet", 0, 3),
    ("PascalContext", "induced-HRNet-w48-cls59", 0, None), ("PascalContext", "induced-HRNet-w48-cls59", 1, None),
    ("TinyImagenet200", "induced", 0, None), ("TinyImagenet200", "wordnet", 1, None), ("TinyImagenet200", "induced", 0, 1),
])
def test_generated_code_matches_oracle_on_shipped_hierarchies(dataset, hierarchy, dtype, warps):
    tree, ot, C = shipped(dataset, hierarchy)
    check(tree, ot, C, dtype=dtype, warps=warps)


@pytest.mark.parametrize("dataset,hierarchy,dtype,warps", [("PascalContext", "induced-HRNet-w48-cls59", 0, None), ("ADE20K", "induced-HRNet-w48", 1, None),
                                                          ("CIFAR100", "wordnet", 2, 3), ("LookIntoPerson", "induced-HRNet-w48-cls20", 0, 2)])
def test_class_major_variant_matches_oracle(dataset, hierarchy, dtype, warps):
    """(N, C, H, W) tensors addressed in place (the segmentation wrappers): same results as the row-major kernels."""
    tree, ot, C = shipped(dataset, hierarchy)
    check(tree, ot, C, dtype=dtype, warps=warps, rows=24, hw=8)


@pytest.mark.parametrize("env", [{"NBDT_B200_KA_ROWS": "2"}, {"NBDT_B200_KA_TEAMS": "4", "NBDT_B200_KA_LOCKSTEP": "8"},
                                 {"NBDT_B200_KA_ROWS": "2", "NBDT_B200_KA_WARPS": "2"}, {"NBDT_B200_KA_FENCE": "8", "NBDT_B200_KA_EXPMUFU": "0"}])
def test_generator_knob_variants_match_oracle(env, monkeypatch):
    """The measured-and-rejected code shapes stay correct: two rows per thread on packed pairs (with and without teams), several
    teams per CTA with lock-step barriers (no-ops on the host), and the round-1 fence / polynomial settings."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    tree, ot, C = shipped("CIFAR100", "wordnet")
    check(tree, ot, C, dtype=0, rows=9, warps=int(env["NBDT_B200_KA_WARPS"]) if "NBDT_B200_KA_WARPS" in env else None)


def test_generated_code_imagenet_team_of_16():
    tree, ot, C = shipped("Imagenet1000", "induced")
    check(tree, ot, C, dtype=0, rows=6)


@pytest.mark.parametrize("leaves,kmax,chain,warps", [(2, 2, 0.0, None), (3, 3, 0.0, None), (17, 4, 0.5, None), (17, 4, 0.5, 2),
                                                     (64, 13, 0.2, None), (64, 2, 0.9, 4), (150, 5, 0.3, None), (257, 2, 0.95, None)])
def test_generated_code_matches_oracle_on_random_hierarchies(leaves, kmax, chain, warps, tmp_path):
    from nbdt.tree import Tree

    pg, pw = write_random_hierarchy(tmp_path, leaves, kmax, seed=leaves * 31 + kmax, chain_bias=chain)
    tree = Tree(None, path_graph=pg, path_wnids=pw, classes=[f"c{i}" for i in range(leaves)])
    check(tree, O.OracleTree(pg, pw), leaves, rows=16, warps=warps)


def test_out_of_range_targets_are_ignored_rows():
    tree, ot, C = shipped("CIFAR100", "induced")
    rng = np.random.default_rng(3)
    x = (3 * rng.standard_normal((8, C))).astype(np.float32)
    y = rng.integers(0, C, 8)
    bad = y.copy()
    bad[[1, 5]] = [-100, C]
    keep = np.array([0, 2, 3, 4, 6, 7])
    for mode, fn, wt in ((SOFT, O.soft_tree_sup_loss, 1.0), (HARD, O.hard_tree_sup_loss, 2.0 / len(ot.inodes))):
        want_l, want_g = fn(ot, x[keep].astype(np.float64), y[keep])
        loss, grad = J.run_loss(tree, mode, 0, 16, x, bad, 1.0, wt, batch_rows=len(keep))
        assert abs(loss - want_l) <= 2e-6 * abs(want_l)
        assert (grad[[1, 5]] == 0).all()
        np.testing.assert_allclose(grad[keep], want_g, rtol=2e-5, atol=1e-6 * np.abs(want_g).max())


def test_source_is_refused_outside_the_limits():
    from nbdt._b200 import lib as L

    lib = L.load()
    ptr = np.zeros(2, dtype=np.int32)
    desc = L.TreeDesc(5000, 1, 0, ptr.ctypes.data_as(L.ctypes.POINTER(L.c_i32)), ptr.ctypes.data_as(L.ctypes.POINTER(L.c_i32)))
    assert lib.nbdt_tree_jit_source(L.ctypes.byref(desc), 0, 0, 16, None, 0) == -1
