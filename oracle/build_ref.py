"""Recipe for ``oracle/_ref``: the reference's OWN hot-path modules, vendored unmodified.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (sajadn/Exemplar-VAE) is pure Python with no
build system; "building" it means placing the modules the hot path imports where they can travel to
the GPU box (``/root/reference`` does not exist there).  The copies land in ``oracle/_ref/`` which is
git-ignored (never part of the repository history) but not gpurun-ignored, exactly like the built
``.so``.  Nothing in the product imports it; only ``bench.py --impl reference`` / the ``cpu_baseline``
leg (through ``oracle/ref_runner.py``) and ``oracle/make_golden.py`` do.

    python oracle/build_ref.py            # run in the build container (needs /root/reference)
"""
from __future__ import annotations

import os
import shutil

REFERENCE = os.environ.get("EXVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

# the modules on the path of SURVEY.md §8(a)/(f); everything else (CLI, data loaders, plots, PixelCNN) stays behind
FILES = [
    "models/__init__.py", "models/BaseModel.py", "models/AbsModel.py", "models/AbsHModel.py", "models/VAE.py",
    "models/HVAE_2level.py", "models/convHVAE_2level.py", "models/fully_conv.py",
    "utils/__init__.py", "utils/distributions.py", "utils/nn.py", "utils/training.py", "utils/optimizer.py",
    "utils/utils.py", "utils/knn_on_latent.py",
]


def build_ref(verbose: bool = False) -> str | None:
    """Copy the hot-path modules; returns the destination, or None when the reference tree is absent
    (GPU box: the prebuilt copy that travelled with the snapshot is used as is)."""
    if not os.path.isdir(REFERENCE):
        return DEST if os.path.isdir(DEST) else None
    for rel in FILES:
        src = os.path.join(REFERENCE, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("vendored", rel)
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as f:
        f.write("Unmodified copies of sajadn/Exemplar-VAE hot-path modules, made by oracle/build_ref.py from "
                f"{REFERENCE}.\nNot part of the repository history (git-ignored); baseline/oracle use only.\n")
    return DEST


if __name__ == "__main__":
    print(build_ref(verbose=True))
