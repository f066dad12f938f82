"""Generate and check hierarchies (reference: nbdt/hierarchy.py:59-127 generate_hierarchy, :137-190 test_hierarchy).
HTML visualisation (generate_hierarchy_vis*) is not part of this package."""
from nbdt.graph import build_induced_graph, build_minimal_wordnet_graph, build_random_graph, get_graph_path_from_args, prune_single_successor_nodes
from nbdt.thirdparty.nx import get_depth, get_leaves, get_roots, read_graph, write_graph
from nbdt.thirdparty.wn import get_wnids_from_dataset
from nbdt.utils import Colors


def print_graph_stats(G, name):
    num_children = [len(G.succ[n]) for n in G.nodes]
    print("[{}] \t Nodes: {} \t Depth: {} \t Max Children: {}".format(name, len(G.nodes), get_depth(G), max(num_children)))


def assert_all_wnids_in_graph(G, wnids):
    missing = [w for w in wnids if w.strip() not in G.nodes]
    assert not missing, missing


def generate_hierarchy(dataset, method, seed=0, branching_factor=2, extra=0, no_prune=False, fname="", path="", single_path=False,
                       induced_linkage="ward", induced_affinity="euclidean", checkpoint=None, arch=None, model=None, **kwargs):
    """Builds the hierarchy, writes it as node-link JSON and returns the path.  method "induced": from `checkpoint`, the pretrained
    `arch`, or a live `model` (its state_dict; this is what Tree.update_from_model / SoftTreeLoss.set_epoch call)."""
    wnids = get_wnids_from_dataset(dataset)
    if method == "wordnet":
        G = build_minimal_wordnet_graph(wnids, single_path)
    elif method == "random":
        G = build_random_graph(wnids, seed=seed, branching_factor=branching_factor)
    elif method == "induced":
        G = build_induced_graph(wnids, dataset=dataset, checkpoint=checkpoint, model=arch, linkage=induced_linkage, affinity=induced_affinity,
                                branching_factor=branching_factor, state_dict=model.state_dict() if model is not None else None)
    else:
        raise NotImplementedError(f'Method "{method}" not yet handled.')
    print_graph_stats(G, "matched")
    assert_all_wnids_in_graph(G, wnids)
    if not no_prune:
        G = prune_single_successor_nodes(G)
        print_graph_stats(G, "pruned")
        assert_all_wnids_in_graph(G, wnids)
    if extra > 0:
        raise NotImplementedError("--extra needs the WordNet corpus")
    path = get_graph_path_from_args(dataset=dataset, method=method, seed=seed, branching_factor=branching_factor, extra=extra, no_prune=no_prune,
                                    fname=fname, path=path, multi_path=not single_path and False, induced_linkage=induced_linkage,
                                    induced_affinity=induced_affinity, checkpoint=checkpoint, arch=arch)
    write_graph(G, path)
    Colors.green("==> Wrote tree to {}".format(path))
    return path


def test_hierarchy(args=None, path=None, dataset=None):
    """Single root, every dataset class is a leaf, no stray leaves: the invariants the shipped hierarchies satisfy."""
    if path is None:
        path = get_graph_path_from_args(**vars(args))
        dataset = args.dataset
    wnids = get_wnids_from_dataset(dataset)
    G = read_graph(path)
    roots = list(get_roots(G))
    assert len(roots) == 1, f"{len(roots)} roots"
    leaves = set(get_leaves(G))
    assert leaves == set(wnids), (sorted(leaves - set(wnids))[:5], sorted(set(wnids) - leaves)[:5])
    multi = [n for n in G.nodes if len(G.pred[n]) > 1]
    print_graph_stats(G, "tested")
    if multi:
        Colors.red(f"==> {len(multi)} nodes have several parents")
    else:
        Colors.green("==> All checks pass!")
    return G


def generate_hierarchy_vis_from(G, dataset, path_html, **kwargs):
    raise NotImplementedError("HTML visualisation of hierarchies is not part of nbdt-b200")
