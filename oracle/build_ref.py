"""Recipe for `oracle/_ref/`: the UNMODIFIED reference modules of the hot path, staged for the CPU arm of bench.py.

The reference is pure Python with no packaging metadata (no setup.py / pyproject: `pip install --target baseline/_ref
/root/reference` has nothing to build), so the "build" of its hot path is a file copy from where the sources lie:

    /root/reference/src/modules/{perception,nca,ncagraph,graph_augmentation}.py   (the step)
    /root/reference/src/utils/{nca_init,damage}.py, src/training/pool.py          (seed, damage, pool)

into `oracle/_ref/` (git-ignored -- reference sources never enter the history -- but NOT gpurun-ignored, so the copy
travels to the GPU box with the snapshot, like a built .so).  `__graft_entry__.build()` runs this when /root/reference
exists (the build container); on the GPU box the staged copy is used as is.  `bench.py --impl reference` and the
`cpu_baseline` / `gpu_eager` legs import from it (kind "reference") and fall back to the oracle port (kind "port") when
it is absent.  TEST / BASELINE INFRASTRUCTURE ONLY: nothing under graph_neural_cellular_automata_b200/ imports it.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["modules/perception.py", "modules/nca.py", "modules/ncagraph.py", "modules/graph_augmentation.py",
         "utils/nca_init.py", "utils/damage.py", "training/pool.py"]


def build(reference: str = "/root/reference", verbose: bool = True) -> bool:
    src = os.path.join(reference, "src")
    if not os.path.isdir(src):
        return os.path.isdir(DST)
    for rel in FILES:
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), d)
    for pkg in ("modules", "utils", "training"):
        open(os.path.join(DST, pkg, "__init__.py"), "a").close()
    if verbose:
        print(f"[oracle/_ref] staged {len(FILES)} reference files from {src}")
    return True


def import_reference():
    """(NeuralCA, NeuralCAGraph, make_seed) from oracle/_ref, or None when it was never staged."""
    if not os.path.isfile(os.path.join(DST, "modules", "ncagraph.py")):
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from modules.nca import NeuralCA                 # noqa: E402
    from modules.ncagraph import NeuralCAGraph       # noqa: E402
    from utils.nca_init import make_seed             # noqa: E402
    return NeuralCA, NeuralCAGraph, make_seed


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref", "ready" if ok else "not available (no /root/reference)")
