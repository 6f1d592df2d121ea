"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference hot path as files that can travel.

TEST / MEASUREMENT INFRASTRUCTURE — never imported by the product package.

``/root/reference`` exists only in the build container.  So that ``bench.py --impl reference`` can
time the reference's own code on the GPU box's host cores, this script copies the four modules of
the path verbatim (no edits; sha256 recorded in ``MANIFEST.json``) into the git-ignored directory
``oracle/_ref/ultrasphere/``.  The directory is listed in ``.gitignore`` (the reference's sources
never enter this repository's history) but not in ``.gpurunignore``, so it ships to the GPU box
like the built ``.so``.  ``oracle/refload.py`` loads them from there when ``/root/reference`` is
absent, behind the same stand-ins for the missing third-party packages (``oracle/stubs``).

    python oracle/make_ref.py        # also run by __graft_entry__.build() when the mount is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC = "/root/reference/src/ultrasphere"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "ultrasphere")
FILES = ("_coordinates.py", "_creation.py", "_integral.py", "_random.py")


def main() -> bool:
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        print(f"make_ref: {SRC} is not mounted; nothing copied")
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {"source": SRC, "files": {}}
    for name in FILES:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        with open(os.path.join(DST, name), "rb") as fh:
            manifest["files"][name] = hashlib.sha256(fh.read()).hexdigest()
    version = "unknown"
    toml = "/root/reference/pyproject.toml"
    if os.path.isfile(toml):
        for line in open(toml):
            if line.strip().startswith("version"):
                version = line.split("=", 1)[1].strip().strip('"')
                break
    manifest["version"] = version
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print(f"make_ref: copied {len(FILES)} files of ultrasphere {version} to {DST}")
    return True


if __name__ == "__main__":
    main()
