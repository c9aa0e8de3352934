#!/usr/bin/env python
"""TEST INFRASTRUCTURE — recipe that makes the UNMODIFIED reference available on the GPU box for timing.

The reference (timothydmorton/isochrones) is pure Python + numba: there is nothing to compile, but `/root/reference` does
not exist on the GPU box.  In the build container this script packs the reference's own Python sources, untouched, into
`oracle/_ref/isochrones_reference.tar.gz` — `oracle/_ref/` is git-ignored (no reference source enters the repository or
its history) and NOT gpurun-ignored, so the archive travels with the snapshot exactly like the built `.so` files.
`bench.py`'s `cpu_baseline` leg unpacks it into a temporary directory and runs `oracle/time_reference.py --root` on it,
so that the as-shipped reference (`BasicStarModel.lnpost(p)`, one Python call per row; and the same loop over a process
pool of all cores) is timed on the box's own host cores next to the C port.  Called by `__graft_entry__.build()`.
"""
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "isochrones_reference.tar.gz")


def build(reference_root="/root/reference"):
    src = os.path.join(reference_root, "isochrones")
    if not os.path.isdir(src):
        return None                                   # the GPU box: use the archive that travelled
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    with tarfile.open(ARCHIVE, "w:gz") as tar:
        for dirpath, dirnames, filenames in os.walk(src):
            dirnames[:] = [d for d in dirnames if d not in ("tests", "__pycache__", "data", "extras")]
            for f in sorted(filenames):
                if f.endswith(".py"):
                    full = os.path.join(dirpath, f)
                    tar.add(full, arcname=os.path.relpath(full, reference_root))
    return ARCHIVE


if __name__ == "__main__":
    print(build(*sys.argv[1:2]))
