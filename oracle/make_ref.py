"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the reference's own Python files for the CPU reference arm.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no C / C++ sources to compile, setup.py: ext_modules=[]).  Its hot-path modules are
copied UNMODIFIED from the read-only tree at /root/reference into the git-ignored directory ``oracle/_ref/`` -- never
into the repository's history -- so that they travel to the GPU box with the snapshot (like the built .so files) and
``bench.py --impl reference`` / the ``cpu_baseline`` leg can time the reference itself there ("kind": "reference")
through ``oracle/ref_harness.py`` (SPH_REFERENCE_ROOT=oracle/_ref).  Only what ``ref_harness.load_reference()`` imports is
staged; ``oracle/_ref/MANIFEST.json`` lists each file with its sha256 so a reader can check nothing was edited."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SPH_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

# directories (every *.py directly inside) and single files that ref_harness.load_reference() touches
DIRS = ["sphdet/iou", "sphdet/bbox/nms", "sphdet/bbox/coder", "sphdet/losses"]
FILES = ["sphdet/bbox/box_formator.py", "sphdet/bbox/kent_formator.py", "mmdet/models/losses/utils.py", "mmdet/models/losses/smooth_l1_loss.py",
         "tests/utils/generate_data.py"]


def stage() -> int:
    if not os.path.isfile(os.path.join(SRC, "sphdet", "iou", "sph_iou_api.py")):
        print("oracle/make_ref.py: no reference tree at %s (GPU box?): keeping whatever oracle/_ref holds" % SRC)
        return 0
    rel = list(FILES)
    for d in DIRS:
        rel += [os.path.join(d, f) for f in sorted(os.listdir(os.path.join(SRC, d))) if f.endswith(".py")]
    manifest = {}
    for r in rel:
        src, dst = os.path.join(SRC, r), os.path.join(DST, r)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[r] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return len(rel)


if __name__ == "__main__":
    n = stage()
    print("staged %d reference files under %s" % (n, DST))
    sys.exit(0)
