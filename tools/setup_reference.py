#!/usr/bin/env python
"""Copy the unmodified reference (zichunhao/lgn-autoencoder, /root/reference, read-only) to the git-ignored baseline/_ref/ so
that it travels to the GPU box with `gpurun`: the reference arm of bench.py and tools/run_reference_cli.py
import / run it from there.  Nothing is modified; assets and notebooks are left out.  No-op when the source is absent (GPU box)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LGAE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def setup(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(SRC, "lgn")):
        return DST if os.path.isdir(DST) else ""
    if os.path.isdir(DST) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "assets", "*.ipynb", "__pycache__", "env"))
    return DST


if __name__ == "__main__":
    print(setup(force="--force" in sys.argv))
