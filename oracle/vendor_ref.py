"""TEST INFRASTRUCTURE — recipe that places the UNMODIFIED reference modules of the hot path under ``oracle/_ref/``.

``/root/reference`` exists only in the build container.  The reference is plain Python without a ``setup.py`` /
``pyproject.toml`` (``pip install /root/reference`` has nothing to install), so this script is the install step of the CPU
arm: it copies the .py files of the packages SURVEY.md §8(a) cites — ``video_depth_anything/``, ``depth_anything_v2/``,
``models/``, ``utils/`` and ``sam2/modeling/`` (imported by the DepthAnythingV2 memory block) — byte for byte into
``oracle/_ref/``, which is git-ignored (never part of the history) but travels to the GPU box with the snapshot like a
built ``.so``.  ``bench.py --impl reference`` / ``cpu_baseline`` then time the reference's own ``nn.Module``s on the box's host
cores (``oracle/cpu_arm.py``).  ``__graft_entry__.build()`` runs it whenever ``/root/reference`` is present."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ["video_depth_anything", "depth_anything_v2", "models", "utils", os.path.join("sam2", "modeling")]
EXTRA_FILES = [os.path.join("sam2", "__init__.py"), "LICENSE"]


def vendor(src_root: str = "/root/reference", dest: str = DEST) -> dict:
    if not os.path.isdir(os.path.join(src_root, "video_depth_anything")):
        raise FileNotFoundError(f"{src_root} does not hold the reference")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    manifest = {}
    for pkg in PACKAGES:
        for dirpath, _, files in os.walk(os.path.join(src_root, pkg)):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                rel = os.path.relpath(os.path.join(dirpath, f), src_root)
                out = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                shutil.copyfile(os.path.join(src_root, rel), out)
                manifest[rel] = hashlib.sha256(open(out, "rb").read()).hexdigest()[:16]
    for rel in EXTRA_FILES:
        if os.path.exists(os.path.join(src_root, rel)):
            os.makedirs(os.path.dirname(os.path.join(dest, rel)) or dest, exist_ok=True)
            shutil.copyfile(os.path.join(src_root, rel), os.path.join(dest, rel))
    json.dump({"source": src_root, "files": manifest}, open(os.path.join(dest, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    m = vendor(*(sys.argv[1:2]))
    print(f"vendored {len(m)} reference files into {DEST}")
