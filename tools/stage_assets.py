#!/usr/bin/env python
"""Stage the reference's binary ASSETS (not sources) next to the repo so they travel to the GPU box.

Copies color.png / disk.png / sky.png (src/renderer/textures/, used at ray_pipeline.rs:63-70) and
lucy.obj (src/renderer/objects/, loaded at scene/mod.rs:23-26) from /root/reference into
assets/_ref/, which is git-ignored (kept out of history) but not gpurun-ignored.  When the
reference tree is absent (GPU box) this is a no-op; when the staged files are absent, the host
code falls back to the seeded synthetic assets in bhusie_b200/assets.py and says so.
"""
import os
import shutil
import sys

REF = "/root/reference/src/renderer"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "assets", "_ref")
FILES = ["textures/color.png", "textures/disk.png", "textures/sky.png", "objects/lucy.obj"]


def stage(verbose: bool = True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print("reference tree not present; nothing staged")
        return False
    os.makedirs(DST, exist_ok=True)
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DST, os.path.basename(rel))
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
            if verbose:
                print("staged", dst)
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() or True else 1)
