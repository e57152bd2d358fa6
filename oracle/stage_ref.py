"""Stage the UNMODIFIED reference sources of the hot path into ``oracle/_ref/`` (git-ignored, shipped to the GPU box).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is a byte
copy of the files that define the path: ``models.py`` (the caller), ``layers/attention.py``, ``layers/encoding.py`` and
``evaluate.py`` (only its ``greedy_search``, evaluate.py:185-202, is used -- by the fixture generator).  Nothing is edited, and
nothing under ``oracle/_ref/`` is ever committed: the copies exist so that
  * ``tests/test_reference_dropin_gpu.py`` can run the reference's own ``models.py`` over ``mmbidaf_b200.layers`` on the B200,
  * ``bench.py --impl reference`` / ``cpu_baseline`` / ``eager_cuda`` can time the real reference (``kind: "reference"``),
on a box where ``/root/reference`` does not exist.  Called from ``__graft_entry__.build()`` in the build container (where the
reference tree is mounted); on the GPU box the already staged files are used as they are.

    python oracle/stage_ref.py            # copies, prints a manifest with sha256 of every file
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("MMBIDAF_REFERENCE", "/root/reference")
FILES = ("models.py", "layers/attention.py", "layers/encoding.py", "evaluate.py")


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def staged() -> bool:
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


def stage(source: str = SOURCE) -> dict:
    """Copy FILES from ``source`` to oracle/_ref/ byte for byte.  Returns {file: sha256}.  No-op (returns the existing manifest)
    when the source tree is absent but a staged copy exists; raises when neither exists."""
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isdir(source):
        if staged() and os.path.exists(manifest_path):
            with open(manifest_path) as f:
                return json.load(f)
        raise FileNotFoundError(f"neither the reference tree ({source}) nor a staged copy ({DEST}) exists")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(source, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        assert _sha(src) == _sha(dst)
        manifest[rel] = _sha(dst)
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    for name, digest in stage().items():
        print(f"{digest[:16]}  oracle/_ref/{name}")
