"""Hierarchy runtime: `Tree` / `Node` over an induced (or WordNet) hierarchy.

API-compatible with the reference's nbdt/tree.py (Node :38-143, Tree :145-241): same
constructor arguments, same attributes (`inodes` sorted by wnid, `leaves` in wnids-file
order, `Node.child_index_to_class_index`, ...).  Differences by design:

  * leaf sets are computed once per tree in a single post-order pass (the reference walks
    `nx.descendants` per child per node) and are ordered (the reference's order is Python
    set order, i.e. PYTHONHASHSEED dependent);
  * `Tree.device_tree(device)` flattens the hierarchy into the int32 tables the CUDA
    kernels consume (see _b200/tables.py, include/nbdt_b200.h).
"""
from collections import defaultdict

from nbdt.thirdparty.nx import get_leaf_to_path, read_graph
from nbdt.thirdparty.wn import FakeSynset, get_wnids, wnid_to_name, wnid_to_synset
from nbdt.utils import (
    DATASET_TO_NUM_CLASSES,
    DATASETS,
    dataset_to_default_path_graph,
    dataset_to_default_path_wnids,
    hierarchy_to_path_graph,
)


def dataset_to_dummy_classes(dataset):
    assert dataset in DATASETS
    return [FakeSynset.create_from_offset(i).wnid for i in range(DATASET_TO_NUM_CLASSES[dataset])]


def add_arguments(parser):
    parser.add_argument(
        "--hierarchy",
        help="Hierarchy to use. If supplied, will be used to generate --path-graph. --path-graph takes precedence.",
    )
    parser.add_argument("--path-graph", help="Path to graph-*.json file.")
    parser.add_argument("--path-wnids", help="Path to wnids.txt file.")


class Node:
    """One node of the hierarchy. For an inner node the decision is over its children:
    `child_index_to_class_index[k]` lists the class indices (leaves) under child k and
    `class_index_to_child_index[c]` is `[k]` when class c lies under child k (else `[]`)."""

    def __init__(self, tree, wnid, other_class=False):
        self.tree = tree
        self.wnid = wnid
        self.synset = wnid_to_synset(wnid)
        self.name = self.synset.name().split(".")[0]
        self.original_classes = tree.classes
        self.num_original_classes = len(tree.wnids_leaves)

        self.num_children = len(self.succ)
        self.has_other = bool(other_class) and not (self.is_root() or self.is_leaf())
        self.num_classes = self.num_children + int(self.has_other)

        self.class_index_to_child_index, self.child_index_to_class_index = self.build_class_mappings()
        self.classes = self.build_classes()
        assert len(self.classes) == self.num_classes, (
            f"Number of classes {self.num_classes} does not equal number of class names found "
            f"({len(self.classes)}): {self.classes}"
        )
        self.leaves = list(self.get_leaves())
        self.num_leaves = len(self.leaves)

    # ---- graph neighbourhood -------------------------------------------------------------
    @property
    def succ(self):
        return self.tree.G.succ[self.wnid]

    @property
    def pred(self):
        return self.tree.G.pred[self.wnid]

    @property
    def children(self):
        return [self.tree.wnid_to_node[w] for w in self.succ]

    @property
    def parents(self):
        return [self.tree.wnid_to_node[w] for w in self.pred]

    @property
    def parent(self):
        parents = self.parents
        return parents[0] if parents else None

    def is_leaf(self):
        return len(self.succ) == 0

    def is_root(self):
        return len(self.pred) == 0

    def get_leaves(self):
        """wnids of the leaves at or below this node."""
        return iter(self.tree._leaf_wnids_under(self.wnid))

    # ---- index maps ------------------------------------------------------------------------
    def wnid_to_class_index(self, wnid):
        return self.tree.wnid_to_class_index[wnid]

    def wnid_to_child_index(self, wnid):
        return list(self.succ).index(wnid)

    def build_class_mappings(self):
        old_to_new, new_to_old = defaultdict(list), defaultdict(list)
        if self.is_leaf():
            return {}, {}
        for new_index, child in enumerate(self.succ):
            for leaf in self.tree._leaf_wnids_under(child):
                old_index = self.tree.wnid_to_class_index[leaf]
                old_to_new[old_index].append(new_index)
                new_to_old[new_index].append(old_index)
        if self.has_other:
            other = self.num_children
            for old_index in range(self.num_original_classes):
                if old_index not in old_to_new:
                    old_to_new[old_index].append(other)
                    new_to_old[other].append(old_index)
        return old_to_new, new_to_old

    def build_classes(self):
        return [
            ",".join(self.original_classes[old] for old in self.child_index_to_class_index[new])
            for new in sorted(self.child_index_to_class_index)
        ]

    @property
    def class_counts(self):
        """Number of original classes in each child."""
        return [len(self.child_index_to_class_index[new]) for new in sorted(self.child_index_to_class_index)]

    @staticmethod
    def dim(nodes):
        return sum(node.num_classes for node in nodes)


class Tree:
    def __init__(self, dataset, path_graph=None, path_wnids=None, classes=None, hierarchy=None):
        if dataset and hierarchy and not path_graph:
            path_graph = hierarchy_to_path_graph(dataset, hierarchy)
        if dataset and not path_graph:
            path_graph = dataset_to_default_path_graph(dataset)
        if dataset and not path_wnids:
            path_wnids = dataset_to_default_path_wnids(dataset)
        if dataset and not classes:
            classes = dataset_to_dummy_classes(dataset)
        self.load_hierarchy(dataset, path_graph, path_wnids, classes)

    def load_hierarchy(self, dataset, path_graph, path_wnids, classes):
        self.dataset = dataset
        self.path_graph = path_graph
        self.path_wnids = path_wnids
        self.classes = classes
        self.G = read_graph(path_graph)
        self.wnids_leaves = get_wnids(path_wnids)
        self.wnid_to_class = dict(zip(self.wnids_leaves, self.classes))
        self.wnid_to_class_index = {wnid: i for i, wnid in enumerate(self.wnids_leaves)}
        self._leaves_under = self._collect_leaves()
        self.wnid_to_node = self.get_wnid_to_node()
        self.nodes = [self.wnid_to_node[wnid] for wnid in sorted(self.wnid_to_node)]
        self.inodes = [node for node in self.nodes if not node.is_leaf()]
        self.leaves = [self.wnid_to_node[wnid] for wnid in self.wnids_leaves]
        self.__dict__.pop("_b200_device_trees", None)  # device tables belong to the old hierarchy

    def __getstate__(self):
        """Copies / pickles (copy.deepcopy of a criterion, torch.save of a model that holds its tree) drop the device-side tables:
        they hold raw library handles and are rebuilt on first use."""
        state = self.__dict__.copy()
        state.pop("_b200_device_trees", None)
        return state

    def _collect_leaves(self):
        """wnid -> tuple of leaf wnids below it (depth-first, child order); one post-order pass."""
        G = self.G
        under = {}
        for start in G.nodes:
            if start in under:
                continue
            stack = [(start, iter(G.succ[start]))]
            while stack:
                node, it = stack[-1]
                advanced = False
                for child in it:
                    if child not in under:
                        stack.append((child, iter(G.succ[child])))
                        advanced = True
                        break
                if advanced:
                    continue
                stack.pop()
                children = list(G.succ[node])
                if not children:
                    under[node] = (node,)
                else:
                    seen, out = set(), []
                    for child in children:
                        for leaf in under[child]:
                            if leaf not in seen:  # DAG hierarchies may reach a leaf twice
                                seen.add(leaf)
                                out.append(leaf)
                    under[node] = tuple(out)
        return under

    def _leaf_wnids_under(self, wnid):
        return self._leaves_under[wnid]

    def update_from_model(self, model, arch, dataset, classes=None, path_wnids=None, path_graph=None):
        from nbdt.hierarchy import generate_hierarchy  # offline tooling; optional

        assert model is not None, "`model` cannot be NoneType"
        path_graph = generate_hierarchy(dataset=dataset, method="induced", arch=arch, model=model, path=path_graph)
        tree = Tree(dataset, path_graph=path_graph, path_wnids=path_wnids, classes=classes, hierarchy="induced")
        self.load_hierarchy(
            dataset=tree.dataset, path_graph=tree.path_graph, path_wnids=tree.path_wnids, classes=tree.classes
        )

    @classmethod
    def create_from_args(cls, args, classes=None):
        return cls(args.dataset, args.path_graph, args.path_wnids, classes=classes, hierarchy=args.hierarchy)

    @property
    def root(self):
        for node in self.inodes:
            if node.is_root():
                return node
        raise UserWarning("Should not be reachable. Tree should always have root")

    def get_wnid_to_node(self):
        return {wnid: Node(self, wnid) for wnid in self.G}

    def get_leaf_to_steps(self):
        """leaf wnid -> root-to-leaf list of {"node", "name", "next_index"} (next_index = child
        index taken at that node; -1 at the leaf)."""
        leaf_to_path = get_leaf_to_path(self.G)
        leaf_to_steps = {}
        for leaf in self.wnids_leaves:
            path = leaf_to_path[leaf]
            next_indices = [index for index, _ in path[1:]] + [-1]
            leaf_to_steps[leaf] = [
                {"node": self.wnid_to_node[wnid], "name": self.wnid_to_node[wnid].name, "next_index": nxt}
                for nxt, (_, wnid) in zip(next_indices, path)
            ]
        return leaf_to_steps

    def device_tree(self, device):
        """Device-resident tables for the CUDA kernels (cached per device)."""
        from nbdt._b200.tables import device_tree

        return device_tree(self, device)

    def visualize(self, path_html, dataset=None, **kwargs):
        from nbdt.hierarchy import generate_hierarchy_vis_from  # offline tooling; optional

        generate_hierarchy_vis_from(self.G, dataset=dataset, path_html=path_html, **kwargs)
