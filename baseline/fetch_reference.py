#!/usr/bin/env python
"""Recipe for ``baseline/_ref/`` (git-ignored, travels to the GPU box with the snapshot).

proroklab/magat_pathplanning has no setup.py / pyproject (``pip install /root/reference`` fails with "neither
'setup.py' nor 'pyproject.toml' found"), and its package ``__init__`` files import matplotlib / easydict /
tensorboardX / ..., which this image does not have.  What the hot path needs from it is five plain-Python files; this
script copies them UNMODIFIED, preserving their relative paths, from the read-only checkout into ``baseline/_ref/``:

    utils/graphUtils/graphML.py                        the reference layer (GraphFilterBatchAttentional, functionals)
    graphs/weights_initializer.py, graphs/models/resnet_pytorch.py
    graphs/models/decentralplanner_GAT*.py             the CNN -> GAT -> MLP planners that construct the layer
    graphs/models/decentralplanner.py                  the GNN baseline planner (GraphFilterBatch)

``oracle/ref_loader.py`` loads them by file path with the packages they import stubbed.  Users: the ``-m gpu`` planner
integration test (reference planner on CPU vs the same planner on our CUDA layer) and ``bench.py --impl reference``
(the reference's own CPU implementation timed on the box's host cores).  Nothing in the product imports it.
"""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MAGAT_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["utils/graphUtils/graphML.py", "graphs/weights_initializer.py", "graphs/models/resnet_pytorch.py",
         "graphs/models/decentralplanner.py"]


def main():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        print(f"fetch_reference: no reference checkout at {SRC}; leaving {DST} as it is", file=sys.stderr)
        return 1
    files = FILES + [os.path.relpath(p, SRC) for p in sorted(glob.glob(os.path.join(SRC, "graphs/models/decentralplanner_GAT*.py")))]
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("Unmodified files of proroklab/magat_pathplanning copied by baseline/fetch_reference.py from "
                f"{SRC}:\n" + "\n".join(files) + "\n")
    print(f"fetch_reference: {len(files)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
