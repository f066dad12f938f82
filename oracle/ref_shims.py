"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import preamble that lets the *unmodified* reference (`/root/reference`, Python)
execute in this container so golden vectors can be generated from it
(`tests/golden/make_golden.py`).  The reference does not import as shipped:

  * `nltk` is imported eagerly (reference nbdt/thirdparty/wn.py:13) and lazily
    outside a try block (wn.py:49); it is not installed.  A stub whose lookup
    raises makes every node a FakeSynset (wn.py:54-57) -- only `Node.name`
    strings change, numerics do not.
  * `pytorchcv` is imported unconditionally (nbdt/models/wideresnet.py:1-5,
    nbdt/models/__init__.py:3); it is not installed.
  * networkx >= 3.6 reads edges from key "edges"; the shipped hierarchy JSONs
    use "links" (read at nbdt/thirdparty/nx.py:75).

`/root/reference` exists only in the build container; nothing under `-m gpu`
tests, smoke() or bench.py may call `install()`.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NBDT_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nbdt"))


def vendored_root():
    """oracle/_ref (made by oracle/make_ref.py: an unmodified copy of the reference that travels to the GPU box), or None."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    return p if os.path.isdir(os.path.join(p, "nbdt")) else None


def install(root=None, strict_hash=True):
    """Make `import nbdt` resolve to the executing reference.  Idempotent.
    root: tree to import from (default /root/reference; bench.py passes the vendored copy).  strict_hash=False drops the
    PYTHONHASHSEED requirement (timing runs: the leaf order inside a child's set does not change the arithmetic volume)."""
    global REFERENCE_ROOT
    if root is not None:
        REFERENCE_ROOT = root
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if strict_hash and os.environ.get("PYTHONHASHSEED") != "0":
        raise RuntimeError(
            "run with PYTHONHASHSEED=0: the reference's leaf ordering per child comes "
            "from a Python set of strings (nbdt/thirdparty/nx.py:21-25)"
        )
    for name in list(sys.modules):
        if name == "nbdt" or name.startswith("nbdt."):
            raise RuntimeError("an `nbdt` package is already imported; run the oracle in its own process")

    nltk = types.ModuleType("nltk")
    nltk.__path__ = []
    nltk.data = types.SimpleNamespace(find=lambda *a, **k: None)
    nltk.download = lambda *a, **k: None
    corpus = types.ModuleType("nltk.corpus")

    class _NoWordNet:
        NOUN = "n"

        def synset_from_pos_and_offset(self, *a):
            raise LookupError("wordnet corpus unavailable")

        def synsets(self, *a, **k):
            return []

    corpus.wordnet = _NoWordNet()
    nltk.corpus = corpus
    sys.modules["nltk"] = nltk
    sys.modules["nltk.corpus"] = corpus

    def _missing(*a, **k):
        raise NotImplementedError("pytorchcv is not installed")

    wrn = types.ModuleType("pytorchcv.models.wrn_cifar")
    wrn.wrn28_10_cifar10 = wrn.wrn28_10_cifar100 = wrn.get_wrn_cifar = _missing
    eff = types.ModuleType("pytorchcv.models.efficientnet")
    eff.__all__ = []
    for name, mod in (
        ("pytorchcv", types.ModuleType("pytorchcv")),
        ("pytorchcv.models", types.ModuleType("pytorchcv.models")),
        ("pytorchcv.models.wrn_cifar", wrn),
        ("pytorchcv.models.efficientnet", eff),
    ):
        sys.modules[name] = mod

    import networkx.readwrite.json_graph as jg

    if not getattr(jg.node_link_graph, "_nbdt_links_patch", False):
        orig = jg.node_link_graph

        def node_link_graph(data, *a, **k):
            k.setdefault("edges", "links")
            return orig(data, *a, **k)

        node_link_graph._nbdt_links_patch = True
        jg.node_link_graph = node_link_graph

    sys.path.insert(0, REFERENCE_ROOT)
    import nbdt  # noqa: F401  (the reference)

    assert os.path.realpath(nbdt.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    return nbdt
