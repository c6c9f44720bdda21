"""ORACLE — TEST INFRASTRUCTURE ONLY.  Recipe that stages the reference's own hot-path modules under `oracle/_ref/`.

    python -m oracle.build_ref            (dev container; needs /root/reference; also run by `__graft_entry__.build()`)

`/root/reference` does not exist on the GPU box, and `oracle/_ref/` is git-ignored (never part of the repository's history)
but NOT gpurun-ignored, so the staged files travel to the box like the in-tree built `.so`.  The files are byte-identical
copies of `/root/reference/src/scldm/{layers,nnets,vae,stochastic_layers,distributions,constants,optimizers}.py` and
`transport/*.py` (the modules SURVEY.md section 8c found importable through a namespace stub); `MANIFEST.json` records the
sha256 of every source so a reader can check that nothing was edited.  They are used ONLY as
  * the checker in tests (`oracle/ref_loader.py` falls back to this directory when `/root/reference` is absent),
  * the thing timed by `bench.py --impl reference` / `cpu_baseline` (`kind: "reference"`),
never on the product path.
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC = "/root/reference/src/scldm"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "scldm")
FILES = ["layers.py", "nnets.py", "vae.py", "stochastic_layers.py", "distributions.py", "constants.py", "optimizers.py",
         "transport/__init__.py", "transport/transport.py", "transport/integrators.py", "transport/path.py", "transport/utils.py"]


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/_ref: {SRC} not present (GPU box?) - keeping whatever is staged")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: staged {len(FILES)} reference modules from {SRC}")
    return True


if __name__ == "__main__":
    build()
