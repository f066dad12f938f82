"""WordNet-id helpers (reference: nbdt/thirdparty/wn.py:16-94).

nltk / the WordNet corpus are optional here: without them every id resolves to a
FakeSynset, exactly the reference's own fallback (wn.py:54-57)."""
import os
from pathlib import Path

from nbdt.utils import fwd, get_directory, makeparentdirs


def maybe_install_wordnet():
    try:
        import nltk
    except ImportError:
        return
    try:
        nltk.data.find("corpora/wordnet")
    except Exception as e:
        print(e)
        try:
            nltk.download("wordnet")
        except Exception as e2:  # offline
            print(e2)


def get_wnids(path_wnids):
    if not os.path.exists(path_wnids):
        parent = Path(fwd()).parent
        print(f"No such file or directory: {path_wnids}. Looking in {parent}")
        path_wnids = parent / path_wnids
    with open(path_wnids) as f:
        return [line.strip() for line in f.readlines()]


def get_wnids_from_dataset(dataset, root="./nbdt/wnids"):
    return get_wnids(f"{get_directory(dataset, root)}.txt")


def write_wnids(wnids, path):
    makeparentdirs(path)
    with open(str(path), "w") as f:
        f.write("\n".join(wnids))


class FakeSynset:
    """Stand-in synset for ids that are not in WordNet (induced inner nodes, custom datasets)."""

    def __init__(self, wnid):
        assert isinstance(wnid, str)
        self.wnid = wnid

    @staticmethod
    def create_from_offset(offset):
        return FakeSynset("f{:08d}".format(offset))

    def offset(self):
        return int(self.wnid[1:])

    def pos(self):
        return "f"

    def name(self):
        return "(generated)"

    def definition(self):
        return "(generated)"


_WORDNET = None


def _wordnet():
    global _WORDNET
    if _WORDNET is None:
        try:
            from nltk.corpus import wordnet

            wordnet.synset_from_pos_and_offset("n", 1740)  # forces the corpus to load
            _WORDNET = wordnet
        except Exception:
            _WORDNET = False
    return _WORDNET


def synset_to_wnid(synset):
    return f"{synset.pos()}{synset.offset():08d}"


def wnid_to_synset(wnid):
    wn = _wordnet()
    if wn:
        try:
            return wn.synset_from_pos_and_offset(wnid[0], int(wnid[1:]))
        except Exception:
            pass
    return FakeSynset(wnid)


def synset_to_name(synset):
    return synset.name().split(".")[0]


def wnid_to_name(wnid):
    return synset_to_name(wnid_to_synset(wnid))
