#!/usr/bin/env python
"""Install the UNMODIFIED reference modules of the planning hot path into baseline/_ref/ (git-ignored; it travels to
the GPU box with the gpurun snapshot, where /root/reference does not exist), so that `bench.py --impl reference` can
time the reference's own pure-Python planner on the GPU box's host cores.

`python -m pip install --no-index --no-build-isolation [--no-deps] --find-links /opt/wheelhouse --target
baseline/_ref /root/reference` was tried first and fails in metadata generation: the reference's setup.py names a
package `gym_auv` over a flat tree with several top-level packages (setuptools refuses the automatic discovery), and
the planner modules (path_planning/*.py) are plain scripts, not a package, so even a successful install would not
contain them.  This script therefore copies, byte for byte, the files the path imports
(path_planning/rrt_dubins.py:1-16) and the shark-grid CSV of config 3 and records their SHA-256.

Nothing under baseline/_ref is product code or test oracle; only bench.py's reference arm reads it.
Run in the build container:  python baseline/install_ref.py
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("AUVRRT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "path_planning/rrt_dubins.py", "path_planning/cost.py", "path_planning/catalina.py",
    "path_planning/motion_plan_state.py", "path_planning/sharkOccupancyGrid.py", "path_planning/sharkEstimate.py",
    "path_planning/shark_data/AUVGrid_prob_500_turn.csv",
]


def main():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        print("install_ref: no reference tree at %s (nothing installed)" % SRC)
        return 1
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    print("install_ref: %d files -> %s" % (len(FILES), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
