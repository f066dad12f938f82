"""Building hierarchies (reference: nbdt/graph.py): the INDUCED hierarchy from a classifier's weight rows, random hierarchies,
graph file naming.  The agglomerative clustering runs on the GPU (nbdt_ward_linkage, csrc/aux_ops.cu) when the weights live
there or a CUDA device exists -- SoftTreeLoss re-induces the hierarchy from the live model during training
(reference nbdt/loss.py:303-315) -- and through scipy only when there is no device at all (offline tooling on a CPU box).
The minimal-WordNet hierarchy needs the WordNet corpus and is built by `build_minimal_wordnet_graph` only when nltk has it.
"""
import argparse
import ctypes
import os
import random
from pathlib import Path

import networkx as nx
import torch

from nbdt.thirdparty.wn import FakeSynset, synset_to_name, synset_to_wnid, wnid_to_synset
from nbdt.utils import DATASETS, METHODS, get_directory

MODEL_FC_KEYS = (
    "fc.weight", "linear.weight", "module.linear.weight", "module.net.linear.weight", "output.weight", "module.output.weight",
    "output.fc.weight", "module.output.fc.weight", "classifier.weight", "classifier.1.weight", "module.classifier.1.weight",
    "model.last_layer.3.weight",
)


def get_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--dataset", choices=DATASETS, default="CIFAR10")
    parser.add_argument("--extra", type=int, default=0, help="Percent extra nodes to add to the tree")
    parser.add_argument("--multi-path", action="store_true")
    parser.add_argument("--no-prune", action="store_true", help="Do not prune.")
    parser.add_argument("--fname", type=str, default="")
    parser.add_argument("--path", type=str, default="")
    parser.add_argument("--method", choices=METHODS, default="induced")
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--branching-factor", type=int, default=2)
    parser.add_argument("--checkpoint", type=str)
    parser.add_argument("--arch", type=str, default="ResNet18")
    parser.add_argument("--induced-linkage", type=str, default="ward")
    parser.add_argument("--induced-affinity", type=str, default="euclidean")
    parser.add_argument("--vis-out-fname", type=str)
    parser.add_argument("--vis-zoom", type=float, default=1.0)
    parser.add_argument("--color", choices=("blue", "blue-green"), default="blue")
    return parser


def generate_graph_fname(method, seed=0, branching_factor=2, extra=0, no_prune=False, fname="", path="", multi_path=False,
                         induced_linkage="ward", induced_affinity="euclidean", checkpoint=None, arch=None, **kwargs):
    """graph-<method>[-seedS][-linkageL][-affinityA][-<arch or checkpoint stem>][-branchB][-extraE][-noprune][-multi]"""
    if path:
        return Path(path).stem
    if fname:
        return fname
    parts = [f"graph-{method}"]
    if method == "random" and seed != 0:
        parts.append(f"seed{seed}")
    if method == "induced":
        assert checkpoint or arch, "Induced hierarchy needs either `arch` or `checkpoint`"
        if induced_linkage not in ("ward", None):
            parts.append(f"linkage{induced_linkage}")
        if induced_affinity not in ("euclidean", None):
            parts.append(f"affinity{induced_affinity}")
        stem = arch
        if checkpoint:
            stem = Path(checkpoint).stem
            if stem.startswith("ckpt-") and stem.count("-") >= 2:
                stem = "-".join(stem.split("-")[2:]).replace("-induced", "")
        parts.append(str(stem))
    if method in ("random", "induced") and branching_factor != 2:
        parts.append(f"branch{branching_factor}")
    if extra > 0:
        parts.append(f"extra{extra}")
    if no_prune:
        parts.append("noprune")
    if multi_path:
        parts.append("multi")
    return "-".join(parts)


def get_graph_path_from_args(dataset, method, path="", **kwargs):
    if path:
        return path
    return os.path.join(get_directory(dataset), generate_graph_fname(method=method, **kwargs) + ".json")


def set_node_label(G, synset):
    nx.set_node_attributes(G, {synset_to_wnid(synset): synset_to_name(synset)}, "label")


def get_roots(G):
    return (n for n in G.nodes if len(G.pred[n]) == 0)


# ------------------------------------------------------------------------------------------------
# induced hierarchy
# ------------------------------------------------------------------------------------------------
def get_centers_from_state_dict(state_dict):
    for key in MODEL_FC_KEYS:
        if key in state_dict:
            return state_dict[key].squeeze().detach()
    return None


def get_centers_from_checkpoint(checkpoint):
    data = torch.load(checkpoint, map_location=torch.device("cpu"))
    state_dict = data
    for key in ("net", "state_dict"):
        if isinstance(data, dict) and key in data:
            state_dict = data[key]
            break
    fc = get_centers_from_state_dict(state_dict)
    assert fc is not None, f"Could not find FC weights in checkpoint {checkpoint} with keys: {list(state_dict)[-6:]}"
    return fc


def get_centers_from_model(model, num_classes, dataset):
    from nbdt import models

    factory = getattr(models, model)
    try:
        net = factory(pretrained=True, num_classes=num_classes, dataset=dataset)
    except TypeError:
        net = factory(pretrained=True, num_classes=num_classes)
    fc = get_centers_from_state_dict(net.state_dict())
    assert fc is not None, f"Could not find FC weights in model {model}"
    return fc


def ward_children(centers):
    """children_ of sklearn.cluster.AgglomerativeClustering(linkage="ward").fit(centers) -- (C-1, 2) cluster ids, merge s creating
    id C + s, smaller id first -- and the merge heights.  Device path: nbdt_ward_linkage."""
    centers = centers.detach().float()
    C, D = centers.shape
    if C == 2:
        return [(0, 1)], [float((centers[0] - centers[1]).norm())]
    if centers.is_cuda or torch.cuda.is_available():
        from nbdt._b200 import lib as L

        dev = centers.device if centers.is_cuda else torch.device("cuda", torch.cuda.current_device())
        x = centers.to(dev).contiguous()
        lib = L.load()
        children = torch.empty((C - 1, 2), dtype=torch.int32, device=dev)
        heights = torch.empty((C - 1,), dtype=torch.float64, device=dev)
        work = torch.empty((lib.nbdt_ward_workspace_bytes(C),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            L.check(lib.nbdt_ward_linkage(ctypes.c_void_p(x.data_ptr()), C, D, ctypes.c_void_p(children.data_ptr()),
                                          ctypes.c_void_p(heights.data_ptr()), ctypes.c_void_p(work.data_ptr()), L.stream_ptr(dev)))
        return [tuple(p) for p in children.cpu().tolist()], heights.cpu().tolist()
    from scipy.cluster.hierarchy import ward  # no CUDA device: offline tooling only

    Z = ward(centers.double().numpy())
    return [(int(a), int(b)) for a, b in Z[:, :2]], [float(h) for h in Z[:, 2]]


def get_wordnet_meaning(G, synsets):
    """Deepest common WordNet hypernym not yet in G; a generated id f<#nodes> when there is none (always, without the corpus)."""
    if not any(s.pos() == "f" for s in synsets):
        common = set(synsets[0].common_hypernyms(synsets[1]))
        for s in synsets[2:]:
            common &= set(synsets[0].common_hypernyms(s))
        while common:
            cand = max(common, key=lambda s: s.max_depth())
            if synset_to_wnid(cand) not in G.nodes:
                return cand
            common.discard(cand)
    return FakeSynset.create_from_offset(len(G.nodes))


def build_induced_graph(wnids, checkpoint, model=None, linkage="ward", affinity="euclidean", branching_factor=2, dataset="CIFAR10",
                        state_dict=None):
    """Binary hierarchy over the classes by Ward clustering of the classifier's weight rows (reference graph.py:402-457)."""
    num_classes = len(wnids)
    assert checkpoint or model or state_dict, "Need to specify either `checkpoint` or `method` or `state_dict`."
    assert linkage == "ward" and affinity == "euclidean" and branching_factor == 2, "only Ward / euclidean / binary induction is implemented"
    if state_dict:
        centers = get_centers_from_state_dict(state_dict)
    elif checkpoint:
        centers = get_centers_from_checkpoint(checkpoint)
    else:
        centers = get_centers_from_model(model, num_classes, dataset)
    assert centers is not None, "Could not find FC weights"
    assert num_classes == centers.size(0), (
        f"The model FC supports {centers.size(0)} classes. However, the dataset {dataset} features {num_classes} classes. "
        "Try passing the `--dataset` with the right number of classes.")
    G = nx.DiGraph()
    for wnid in wnids:
        G.add_node(wnid)
        set_node_label(G, wnid_to_synset(wnid))
    children, _ = ward_children(centers)
    index_to_wnid = {}
    for index, pair in enumerate(children):
        child_wnids = [wnids[c] if c < num_classes else index_to_wnid[c - num_classes] for c in pair]
        parent = get_wordnet_meaning(G, [wnid_to_synset(w) for w in child_wnids])
        parent_wnid = synset_to_wnid(parent)
        G.add_node(parent_wnid)
        set_node_label(G, parent)
        index_to_wnid[index] = parent_wnid
        for w in child_wnids:
            G.add_edge(parent_wnid, w)
    assert len(list(get_roots(G))) == 1, list(get_roots(G))
    return G


# ------------------------------------------------------------------------------------------------
# random hierarchy, pruning
# ------------------------------------------------------------------------------------------------
def build_random_graph(wnids, seed=0, branching_factor=2):
    """Balanced random hierarchy: shuffle the leaves (seed >= 0), group them bottom-up by `branching_factor`."""
    rng = random.Random(seed)
    wnids = list(wnids)
    if seed >= 0:
        rng.shuffle(wnids)
    level = wnids
    while len(level) > 1:
        level = [level[i:i + branching_factor] for i in range(0, len(level), branching_factor)]
    G = nx.DiGraph()
    G.add_node("0")
    nx.set_node_attributes(G, {"0": ""}, "label")
    counter = [1]

    def attach(group, parent):
        if len(group) == 1:
            child = group[0]
            if isinstance(child, str):
                G.add_node(child)
                set_node_label(G, wnid_to_synset(child))
                G.add_edge(parent, child)
            else:
                attach(child, parent)
            return
        for cand in group:
            if isinstance(cand, str):
                G.add_node(cand)
                set_node_label(G, wnid_to_synset(cand))
                G.add_edge(parent, cand)
            else:
                name = str(counter[0])
                counter[0] += 1
                G.add_node(name)
                nx.set_node_attributes(G, {name: ""}, "label")
                G.add_edge(parent, name)
                attach(cand, name)

    attach(level[0] if level and not isinstance(level[0], str) else level, "0")
    return G


def prune_single_successor_nodes(G):
    """Contract every inner node that has exactly one child into that child."""
    for node in list(G.nodes):
        if node in G and len(G.succ[node]) == 1:
            succ = next(iter(G.succ[node]))
            G = nx.contracted_nodes(G, succ, node, self_loops=False)
    return G


def build_minimal_wordnet_graph(wnids, multi_path=False):
    raise NotImplementedError("the minimal-WordNet hierarchy needs the WordNet corpus; the shipped graph-wordnet*.json files were built with it")


def augment_graph(G, extra, allow_imaginary=False, seed=0, max_retries=10000):
    raise NotImplementedError("graph augmentation (--extra) needs the WordNet corpus")
