"""Graph helpers over networkx DiGraphs (hierarchy JSON in node-link format).

Counterpart of the reference's nbdt/thirdparty/nx.py (read_graph :69-75, get_leaves :21-25,
get_roots :28-31, get_leaf_to_path :49-60).  Unlike the reference, traversal order is
deterministic (child order = edge order), not Python-set order.
"""
import json
import os
from pathlib import Path

import networkx as nx
from networkx.readwrite.json_graph import node_link_data, node_link_graph

from nbdt.utils import fwd, makeparentdirs


def is_leaf(G, node):
    return len(G.succ[node]) == 0


def get_leaves(G, root=None):
    """Leaves of G (below `root` if given), depth-first in child order."""
    starts = [root] if root is not None else list(get_roots(G))
    seen = set()
    for start in starts:
        stack = [start]
        while stack:
            node = stack.pop()
            if node in seen:
                continue
            seen.add(node)
            children = list(G.succ[node])
            if not children:
                yield node
            stack.extend(reversed(children))


def get_roots(G):
    for node in G.nodes:
        if len(G.pred[node]) == 0:
            yield node


def get_root(G):
    roots = list(get_roots(G))
    assert len(roots) == 1, f"Multiple ({len(roots)}) found"
    return roots[0]


def get_depth(G):
    depth = {}
    for node in reversed(list(nx.topological_sort(G))):
        depth[node] = 1 + max((depth[c] for c in G.succ[node]), default=0)
    return max(depth[r] for r in get_roots(G))


def get_leaf_to_path(G):
    """leaf -> [(child index taken to reach node, node), ...] from the root (index 0 at the root)."""
    leaf_to_path = {}
    for root in get_roots(G):
        frontier = [(root, [(0, root)])]
        while frontier:
            node, path = frontier.pop()
            children = list(G.succ[node])
            if not children:
                leaf_to_path[node] = path
            for i, child in enumerate(children):
                frontier.append((child, path + [(i, child)]))
    return leaf_to_path


def _resolve(path):
    if os.path.exists(path):
        return path
    parent = Path(fwd()).parent
    print(f"No such file or directory: {path}. Looking in {parent}")
    return parent / path


def write_graph(G, path):
    makeparentdirs(path)
    with open(str(path), "w") as f:
        json.dump(node_link_data(G, edges="links"), f)


def read_graph(path):
    with open(_resolve(path)) as f:
        return node_link_graph(json.load(f), edges="links")
