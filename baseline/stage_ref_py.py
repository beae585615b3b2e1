#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE.  Stages the reference's *Python* tree (models/, lib/, cfgs/,
epipolar_utils.py, utils.py ...) under baseline/_ref/py/ so that configs[4] — the reference's
own SFMnet.forward and epipolar_utils callers, unmodified — can run on the GPU box, where
/root/reference does not exist.  baseline/_ref/ is git-ignored (nothing of the reference
enters the history) but not gpurun-ignored, so the staged copy travels with the snapshot.

    python baseline/stage_ref_py.py [--ref /root/reference] [--force]

Only plain copies are made; no file is edited.  RANSAC_FiveP (the native extension) is not
staged here: its compiled forms are built in place by oracle/build_ref.sh.
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "py")
SKIP_DIRS = {"RANSAC_FiveP", ".git", "__pycache__"}
KEEP_EXT = {".py", ".yml", ".yaml", ".txt", ".md"}


def stage(ref="/root/reference", force=False):
    if not os.path.isdir(ref):
        print(f"reference not present at {ref}; nothing to stage", file=sys.stderr)
        return None
    stamp = os.path.join(DEST, ".staged")
    if os.path.exists(stamp) and not force:
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    n = 0
    for root, dirs, files in os.walk(ref):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, ref)
        for f in files:
            if os.path.splitext(f)[1] not in KEEP_EXT:
                continue
            dst_dir = os.path.join(DEST, rel) if rel != "." else DEST
            os.makedirs(dst_dir, exist_ok=True)
            shutil.copyfile(os.path.join(root, f), os.path.join(dst_dir, f))
            n += 1
    with open(stamp, "w") as fh:
        fh.write(f"{n} files copied unmodified from {ref}\n")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    d = stage(a.ref, a.force)
    print(d if d else "not staged")
