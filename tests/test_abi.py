"""The C-ABI library loads and exports every symbol include/nbdt_b200.h declares; argument
validation that needs no GPU.  (No compute calls here.)"""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbdt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbdt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from nbdt._b200 import lib as L

    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in nbdt_b200.h but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature in nbdt/_b200/lib.py"
    assert sorted(L.SIGNATURES) == declared
    assert lib.nbdt_version() == 105


def _desc(L, C, ptr, child, root=0):
    ptr = np.asarray(ptr, dtype=np.int32)
    child = np.asarray(child, dtype=np.int32)
    d = L.TreeDesc(C, len(ptr) - 1, root, ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                   child.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    d._keep = (ptr, child)
    return d


@pytest.mark.parametrize("ptr,child,msg", [
    ([0, 2, 4], [0, 1, 2, 1], "more than one parent"),      # leaf 1 twice
    ([0, 2, 3], [0, 1, 2], "not a tree"),                    # slot count != nodes-1
    ([0, 2, 4], [0, 1, 2, 9], "out of range"),
    ([0, 2, 4], [0, 1, 2, 3], "more than one parent"),       # root (id 3) as a child
])
def test_tree_create_rejects_malformed_hierarchies(ptr, child, msg):
    from nbdt._b200 import lib as L

    lib = L.load()
    handle = ctypes.c_void_p()
    d = _desc(L, 3, ptr, child)
    rc = lib.nbdt_tree_create(ctypes.byref(d), 0, ctypes.byref(handle))
    assert rc in (1, 3) and not handle.value
    assert msg in lib.nbdt_last_error().decode()


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    from nbdt._b200 import lib as L

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    assert lib.nbdt_device_count() == 0
    handle = ctypes.c_void_p()
    d = _desc(L, 3, [0, 2, 4], [0, 4, 1, 2])  # root(3) -> leaf0, inode1(4) -> leaves 1,2
    rc = lib.nbdt_tree_create(ctypes.byref(d), 0, ctypes.byref(handle))
    assert rc == 2 and not handle.value  # NBDT_ECUDA: no silent CPU path
    assert lib.nbdt_last_error()
    assert lib.nbdt_soft_forward(None, None, 0, 1, 3, None, None) != 0


def test_cpu_tensors_raise():
    import torch
    import torch.nn as nn
    from nbdt._b200.lib import NBDTB200Error
    from nbdt.loss import HardTreeSupLoss, SoftTreeSupLoss
    from nbdt.model import HardEmbeddedDecisionRules, SoftEmbeddedDecisionRules

    x, y = torch.randn(2, 10), torch.tensor([1, 2])
    for rules in (SoftEmbeddedDecisionRules(dataset="CIFAR10", hierarchy="induced"),
                  HardEmbeddedDecisionRules(dataset="CIFAR10", hierarchy="induced")):
        with pytest.raises(NBDTB200Error):
            rules(x)
    for cls in (SoftTreeSupLoss, HardTreeSupLoss):
        with pytest.raises(NBDTB200Error):
            cls(dataset="CIFAR10", criterion=nn.CrossEntropyLoss(), hierarchy="induced")(x, y)
