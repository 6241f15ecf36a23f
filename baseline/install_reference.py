"""Puts the UNMODIFIED reference sources of the hot path under baseline/_ref/ (git-ignored, travels to the GPU box).

    python baseline/install_reference.py            # authoring container only: needs /root/reference

The reference has no packaging (no setup.py / pyproject), so `pip install --target baseline/_ref /root/reference` has nothing
to build; what an install would have produced is exactly these files: lib/diffusion/** (score network, SDE, sampler,
losses, trainer, evaler), configs/*.py, the grid mask the sampler multiplies with, main_diffusion.py and the licence.
Nothing is edited: baseline/reference_arm.py applies the two run-time patches SURVEY section 8(c) lists (an
`ml_collections` stand-in on sys.path, and `Tensor.cuda` as the identity when no GPU is present) from the outside.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MDB_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref")
ITEMS = ["lib/diffusion", "configs", "data/grid_mask_64.pt", "main_diffusion.py", "LICENSE"]


def install(force=False):
    if not os.path.isdir(REF):
        return None  # the GPU box: the prebuilt copy travelled with the snapshot
    if os.path.isdir(DST) and not force and all(os.path.exists(os.path.join(DST, i)) for i in ITEMS):
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    for item in ITEMS:
        src, dst = os.path.join(REF, item), os.path.join(DST, item)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isdir(src):
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copyfile(src, dst)
    for root, _, files in os.walk(DST):  # the reference tree is read-only; the copy must be removable
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
