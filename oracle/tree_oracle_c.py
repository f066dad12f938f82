"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper over the plain-C restatement oracle/tree_oracle.c (same functions as
oracle/tree_oracle.py, float32 logits in, float64 out).  Used by tests/ (pinned against the numpy oracle and the golden
vectors of the executing reference) and by bench_tree.py's cpu_baseline leg; never by the product path."""
import ctypes

import numpy as np

from . import build_oracle

_lib = None


class _OTree(ctypes.Structure):
    _fields_ = [("C", ctypes.c_int), ("N", ctypes.c_int), ("S", ctypes.c_int),
                ("child_ptr", ctypes.c_void_p), ("leaf_ptr", ctypes.c_void_p), ("leaf_idx", ctypes.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle.build())
        _lib.nbdt_oracle_threads.restype = ctypes.c_int
        _lib.nbdt_oracle_soft_loss.restype = ctypes.c_double
        _lib.nbdt_oracle_hard_loss.restype = ctypes.c_double
        _lib.nbdt_oracle_soft_probs.restype = None
    return _lib


class CTree:
    """Flat tables of an oracle.tree_oracle.OracleTree for the C functions."""

    def __init__(self, tree):
        self.C = tree.num_classes
        self.N = len(tree.inodes)
        child_ptr, leaf_ptr, leaf_idx = [0], [0], []
        coc = -np.ones((self.N, self.C), dtype=np.int32)
        for i, w in enumerate(tree.inodes):
            for k, idx in enumerate(tree.child_classes[w]):
                leaf_idx.extend(idx)
                leaf_ptr.append(len(leaf_idx))
                coc[i, idx] = k
            child_ptr.append(len(leaf_ptr) - 1)
        self.child_ptr = np.asarray(child_ptr, dtype=np.int32)
        self.leaf_ptr = np.asarray(leaf_ptr, dtype=np.int32)
        self.leaf_idx = np.asarray(leaf_idx, dtype=np.int32)
        self.child_of_class = np.ascontiguousarray(coc)
        self.S = len(leaf_ptr) - 1
        self.struct = _OTree(self.C, self.N, self.S, self.child_ptr.ctypes.data, self.leaf_ptr.ctypes.data,
                             self.leaf_idx.ctypes.data)


def _prep(x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float32)
    if y is None:
        return x
    return x, np.ascontiguousarray(y, dtype=np.int64)


def threads():
    return int(lib().nbdt_oracle_threads())


def soft_probs(ct, x):
    x = _prep(x)
    out = np.empty((x.shape[0], ct.C), dtype=np.float64)
    lib().nbdt_oracle_soft_probs(ctypes.byref(ct.struct), x.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(x.shape[0]),
                                 ctypes.c_longlong(x.shape[1]), out.ctypes.data_as(ctypes.c_void_p))
    return out


def soft_tree_sup_loss(ct, x, y, xent_weight=1.0, tree_weight=1.0, want_grad=True):
    x, y = _prep(x, y)
    g = np.empty((x.shape[0], ct.C), dtype=np.float64) if want_grad else None
    loss = lib().nbdt_oracle_soft_loss(ctypes.byref(ct.struct), x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.c_longlong(x.shape[0]), ctypes.c_longlong(x.shape[1]), ctypes.c_double(xent_weight),
                                       ctypes.c_double(tree_weight), g.ctypes.data_as(ctypes.c_void_p) if want_grad else None)
    return float(loss), g


def hard_tree_sup_loss(ct, x, y, xent_weight=1.0, tree_weight=1.0, tree_supervision_weight=1.0, want_grad=True):
    x, y = _prep(x, y)
    g = np.empty((x.shape[0], ct.C), dtype=np.float64) if want_grad else None
    loss = lib().nbdt_oracle_hard_loss(ctypes.byref(ct.struct), ct.child_of_class.ctypes.data_as(ctypes.c_void_p),
                                       x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.c_longlong(x.shape[0]), ctypes.c_longlong(x.shape[1]), ctypes.c_double(xent_weight),
                                       ctypes.c_double(tree_weight), ctypes.c_double(tree_supervision_weight),
                                       g.ctypes.data_as(ctypes.c_void_p) if want_grad else None)
    return float(loss), g
