"""Flatten an nbdt.tree.Tree into the tables of nbdt_tree_desc and own the device copies."""
import ctypes
import weakref

import numpy as np

from . import lib as L


def flatten(tree):
    """(child_ptr[N+1], child_node[S], root) with node ids: leaf c -> c, inode i -> C + i.

    inode order = tree.inodes (sorted by wnid, reference nbdt/tree.py:172-173); child order =
    Node.children (edge order of the graph JSON, reference nbdt/tree.py:88-94)."""
    C = len(tree.wnids_leaves)
    inode_index = {node.wnid: i for i, node in enumerate(tree.inodes)}
    child_ptr = [0]
    child_node = []
    for node in tree.inodes:
        for wnid in node.succ:
            if wnid in inode_index:
                child_node.append(C + inode_index[wnid])
            else:
                child_node.append(tree.wnid_to_class_index[wnid])
        child_ptr.append(len(child_node))
    root = inode_index[tree.root.wnid]
    return np.asarray(child_ptr, dtype=np.int32), np.asarray(child_node, dtype=np.int32), root


class DeviceTree:
    """RAII wrapper around nbdt_tree* for one CUDA device."""

    def __init__(self, tree, device_index):
        lib = L.load()
        self.child_ptr, self.child_node, self.root = flatten(tree)
        self.num_classes = len(tree.wnids_leaves)
        self.num_inodes = len(tree.inodes)
        desc = L.TreeDesc(
            self.num_classes,
            self.num_inodes,
            self.root,
            self.child_ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            self.child_node.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        )
        handle = ctypes.c_void_p()
        L.check(lib.nbdt_tree_create(ctypes.byref(desc), int(device_index), ctypes.byref(handle)))
        self.handle = handle
        self.device_index = int(device_index)
        self.max_depth = lib.nbdt_tree_max_depth(handle)
        self.num_slots = lib.nbdt_tree_num_slots(handle)
        self._finalizer = weakref.finalize(self, lib.nbdt_tree_destroy, handle)


def device_tree(tree, device):
    """Device tables of `tree` on `device`, built on first use and cached on the Tree."""
    import torch

    index = device.index if device.index is not None else torch.cuda.current_device()
    cache = tree.__dict__.setdefault("_b200_device_trees", {})
    dt = cache.get(index)
    if dt is None:
        dt = cache[index] = DeviceTree(tree, index)
    return dt
