"""Recipe for ``oracle/_ref``: the UNMODIFIED reference operators, placed where the GPU box can import them.

    python oracle/build_ref.py            (run in the build container; __graft_entry__.build() calls it)

The reference (quentinll/pertrenderer) is four pure-Python files, ``randomras/{__init__,smoothrast,smoothagg,
random_rasterizer}.py``.  ``/root/reference`` does not exist on the GPU box, and reference sources are never copied into
the repository's history: this script copies them, byte for byte, into ``oracle/_ref/randomras/`` -- a directory that is
git-ignored but travels to the GPU box with the snapshot, like the built ``.so`` -- and records their SHA-256 in
``oracle/_ref/MANIFEST.json``.  ``oracle/ref_loader.py`` imports the package from there with the pytorch3d modules it
names at import time stubbed (pytorch3d 0.4.0 is not installable in this image; the hot path touches pytorch3d objects by
attribute access only).

Test infrastructure only: ``bench.py --impl reference`` / ``also.reference_torch_cuda`` time it, nothing in
``pertrenderer_b200`` imports it.
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["__init__.py", "smoothrast.py", "smoothagg.py", "random_rasterizer.py"]


def install(verbose: bool = True) -> bool:
    """Copy the reference package into oracle/_ref.  Returns False (and leaves oracle/_ref alone) when the reference
    tree is absent, e.g. on the GPU box."""
    src = os.path.join(REF, "randomras")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/_ref: {src} not present, nothing to do")
        return False
    dst = os.path.join(DST, "randomras")
    os.makedirs(dst, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        with open(os.path.join(dst, f), "rb") as fh:
            manifest["files"][f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    if verbose:
        print(f"oracle/_ref: installed {len(FILES)} unmodified reference files into {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
