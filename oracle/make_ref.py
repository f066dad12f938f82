"""TEST / BENCH INFRASTRUCTURE: vendor the UNMODIFIED reference into oracle/_ref/ so that it can run where /root/reference does
not exist (the GPU box).

    python oracle/make_ref.py        (also run by __graft_entry__.build() when /root/reference is present)

oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels with the snapshot.
`bench.py --impl reference` and the `cpu_baseline` leg import it -- with the three import shims of oracle/ref_shims.py, the
package itself byte-for-byte as shipped -- and time the reference's own loss (nbdt/loss.py) and backbones (nbdt/models) on the
host cores.  Only the Python package, its hierarchies and wnids are copied (no checkpoints, no templates, no .git).
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("NBDT_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build():
    """Returns the path of the vendored copy, or None when there is no reference tree to copy from."""
    if not os.path.isdir(os.path.join(SRC, "nbdt")) or os.path.realpath(SRC) == os.path.realpath(DST):
        return DST if os.path.isdir(os.path.join(DST, "nbdt")) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, "nbdt"), os.path.join(DST, "nbdt"), symlinks=True,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "templates", "bin"))
    for f in ("main.py", "LICENSE"):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    with open(os.path.join(DST, "README.vendored"), "w") as f:
        f.write("Unmodified copy of %s/nbdt made by oracle/make_ref.py; not part of the repository history.\n" % SRC)
    return DST


if __name__ == "__main__":
    print(build())
