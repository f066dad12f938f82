"""nbdt -- B200-native drop-in for the hot path of alvinwan/neural-backed-decision-trees.

Same import surface as the reference package (`nbdt.model`, `nbdt.loss`, `nbdt.tree`,
`nbdt.models`, ...); the tree head, the tree-supervision losses and the backbones run as
hand-written sm_100a CUDA kernels behind the C ABI in include/nbdt_b200.h
(neural-backed-decision-trees_b200/csrc).  There is no CPU implementation.
"""
__version__ = "0.0.4+b200"
