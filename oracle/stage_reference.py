"""Stages the reference's five torch-only hot-path modules into the git-ignored ``oracle/_ref/`` so that the GPU box
(which has no ``/root/reference``; ``gpurun`` ships ``oracle/_ref/`` with the snapshot) can time and check against the REAL
reference code (``bench.py --impl reference`` reports ``kind: "reference"``; ``oracle/_refload.py`` finds it).

TEST INFRASTRUCTURE ONLY. The files are copied byte for byte at build time (``__graft_entry__.build()``), never committed:
``oracle/_ref/`` is listed in ``.gitignore``. The reference is pure Python (no build step of its own).

    python oracle/stage_reference.py [reference_root]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
# reference files on the north-star path that import nothing but torch
FILES = [
    "mimo/__init__.py",
    "mimo/losses.py",
    "mimo/models/__init__.py",
    "mimo/models/utils.py",
    "mimo/models/mimo_components/__init__.py",
    "mimo/models/mimo_components/model.py",
    "mimo/models/mimo_components/components.py",
    "mimo/models/mimo_components/loss_buffer.py",
]


def stage(ref_root: str = "/root/reference") -> bool:
    if not os.path.isfile(os.path.join(ref_root, "mimo", "models", "mimo_components", "model.py")):
        return False
    for rel in FILES:
        src, dst = os.path.join(ref_root, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isfile(src):
            shutil.copyfile(src, dst)
        else:  # package markers that the reference does not ship
            open(dst, "a").close()
    return True


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged into", DEST if ok else "(nothing: reference checkout not found)")
