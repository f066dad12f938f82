"""Golden vectors for the induced-hierarchy construction, produced by EXECUTING the unmodified reference
(nbdt.graph.build_induced_graph, /root/reference/nbdt/graph.py:402-457) on seeded classifier matrices.

    PYTHONHASHSEED=0 python tests/golden/make_golden_induced.py

One patch beyond oracle/ref_shims.py: scikit-learn >= 1.4 renamed AgglomerativeClustering's `affinity` to `metric` (the reference
passes `affinity=`, graph.py:437-439 -> TypeError with the installed 1.9); the wrapper below forwards it.  Output:
tests/golden/induced_graphs.json = {case: {"C", "D", "seed", "dataset", "children" (sklearn's children_), "heights",
"graph" (node-link data of the graph the reference built)}}.  The weights are regenerated from the seed by the tests.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

CASES = [("cifar10", "CIFAR10", 10, 64, 0), ("cifar100", "CIFAR100", 100, 640, 1), ("tiny200", "TinyImagenet200", 200, 512, 2),
         ("imagenet1000", "Imagenet1000", 1000, 1280, 3)]


def weights(C, D, seed):
    g = torch.Generator().manual_seed(seed)
    # clustered rows (a trained classifier's rows are far from i.i.d.): class centres around a few group centres
    groups = torch.randn(max(2, C // 7), D, generator=g)
    return (groups[torch.randint(len(groups), (C,), generator=g)] + 0.5 * torch.randn(C, D, generator=g)).float()


def main():
    ref_shims.install()
    import nbdt.graph as RG
    from networkx.readwrite.json_graph import node_link_data
    from scipy.cluster.hierarchy import ward
    from sklearn.cluster import AgglomerativeClustering as AC

    def patched(linkage="ward", n_clusters=2, affinity="euclidean"):
        return AC(linkage=linkage, n_clusters=n_clusters, metric=affinity)

    RG.AgglomerativeClustering = patched
    os.chdir(ref_shims.REFERENCE_ROOT)
    out = {}
    for name, dataset, C, D, seed in CASES:
        W = weights(C, D, seed)
        wnids = RG.get_wnids_from_dataset(dataset) if hasattr(RG, "get_wnids_from_dataset") else None
        if wnids is None:
            from nbdt.thirdparty.wn import get_wnids_from_dataset

            wnids = get_wnids_from_dataset(dataset)
        G = RG.build_induced_graph(wnids, checkpoint=None, state_dict={"fc.weight": W}, dataset=dataset)
        children = AC(linkage="ward", n_clusters=2).fit(W.numpy()).children_
        Z = ward(W.double().numpy())
        assert (Z[:, :2].astype(int) == children).all()
        out[name] = {"C": C, "D": D, "seed": seed, "dataset": dataset, "children": children.tolist(), "heights": Z[:, 2].tolist(),
                     "graph": node_link_data(G, edges="links")}
        print(name, len(G.nodes), "nodes")
    with open(os.path.join(ROOT, "tests", "golden", "induced_graphs.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
