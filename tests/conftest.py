import json
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-backed-decision-trees_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


TREE_CASES = [
    "cifar10_induced", "cifar10_induced_resnet18", "cifar10_wordnet", "cifar100_induced", "cifar100_wordnet",
    "tiny200_induced_resnet18", "tiny200_wordnet", "imagenet1000_induced", "ade20k_induced", "lip_induced",
]


@pytest.fixture(scope="session")
def structure():
    with open(os.path.join(GOLDEN, "structure.json")) as f:
        return json.load(f)


def load_golden(case):
    return np.load(os.path.join(GOLDEN, f"tree_{case}.npz"))


def case_paths(structure, case):
    from nbdt.utils import dataset_to_default_path_wnids, hierarchy_to_path_graph

    s = structure[case]
    return hierarchy_to_path_graph(s["dataset"], s["hierarchy"]), dataset_to_default_path_wnids(s["dataset"])


def write_random_hierarchy(tmpdir, num_leaves, max_children, seed, chain_bias=0.0):
    """Random rooted tree with `num_leaves` leaves and 2..max_children children per inner node,
    written in the reference's node-link JSON + wnids format.  Returns (path_graph, path_wnids)."""
    rng = random.Random(seed)
    leaves = [f"f{i:08d}" for i in range(num_leaves)]
    order = leaves[:]
    rng.shuffle(order)
    pool = order  # current forest roots
    links, nodes = [], list(leaves)
    nid = 0
    while len(pool) > 1:
        k = min(len(pool), rng.randint(2, max_children))
        if rng.random() < chain_bias:  # deep, unbalanced trees: merge the newest root with fresh ones
            picked = [pool.pop()] + [pool.pop(0) for _ in range(k - 1)]
        else:
            picked = [pool.pop(rng.randrange(len(pool))) for _ in range(k)]
        parent = f"n{nid:08d}"
        nid += 1
        nodes.append(parent)
        for c in picked:
            links.append({"source": parent, "target": c})
        pool.append(parent)
    rng.shuffle(links)  # edge order defines child order; inode order is sorted(wnid)
    # keep sibling order stable per parent but interleave parents
    data = {"directed": True, "multigraph": False, "graph": {},
            "nodes": [{"id": n} for n in nodes], "links": links}
    pg = os.path.join(str(tmpdir), f"graph-random-{num_leaves}-{max_children}-{seed}.json")
    pw = os.path.join(str(tmpdir), f"wnids-{num_leaves}-{seed}.txt")
    with open(pg, "w") as f:
        json.dump(data, f)
    with open(pw, "w") as f:
        f.write("\n".join(leaves))
    return pg, pw
