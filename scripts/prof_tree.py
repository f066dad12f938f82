"""Launch the fused tree-loss kernel a few times at a saturating size (for ncu captures).
usage: python scripts/prof_tree.py [dataset hierarchy rows dtype]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-backed-decision-trees_b200"))
import torch
from nbdt.tree import Tree
from nbdt._b200 import ops

dataset = sys.argv[1] if len(sys.argv) > 1 else "CIFAR10"
hierarchy = sys.argv[2] if len(sys.argv) > 2 else "induced"
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 22
dtype = getattr(torch, sys.argv[4]) if len(sys.argv) > 4 else torch.float32
tree = Tree(dataset, hierarchy=hierarchy)
C = len(tree.leaves)
x = (3 * torch.randn((rows, C), device="cuda:0")).to(dtype)
y = torch.randint(C, (rows,), device="cuda:0")
for kind in ("soft", "hard", "soft"):
    ops.tree_loss(tree, x, y, kind, 1.0, 1.0)
torch.cuda.synchronize()
