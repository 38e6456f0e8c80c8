"""ORACLE support (test infrastructure): snapshot the reference's hot-path package into ``oracle/_ref/``.

    python -m oracle.snapshot_ref            # /root/reference/mmpfn/models/mmpfn -> oracle/_ref/mmpfn/...

``/root/reference`` does not exist on the GPU box.  ``oracle/_ref/`` is git-ignored (reference
sources never enter this repository's history) but NOT gpurun-ignored, so the UNMODIFIED reference
package travels with the snapshot and can be imported there through ``oracle/ref_compat.py``.  It is
used only as the checker / the baseline being timed:

* ``tests/test_gpu_plugin.py`` — the reference's own ``MMPFNClassifier`` with this repo's CUDA model
  plugged in, against the reference's CPU fp32 ``predict_proba``;
* ``bench.py --impl reference`` — the reference's own ``predict_proba`` on the host cores;
* ``bench.py`` ``gpu_reference`` — the reference on ``device="cuda"`` (torch eager + SDPA) on the same B200.

Only ``mmpfn/__init__.py``, ``mmpfn/models/__init__.py`` and ``mmpfn/models/mmpfn/**.py`` are taken
(460 kB of Python; no data, no vendored DINOv2 / vanilla TabPFN copy).  ``__graft_entry__.build()``
runs this when ``/root/reference`` is present.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PACKAGE = os.path.join("mmpfn", "models", "mmpfn")


def snapshot(src_root: str = SRC_ROOT, dst_root: str = DST_ROOT) -> str | None:
    """Copies the package (byte for byte); returns the destination or None when there is no source."""
    src_pkg = os.path.join(src_root, PACKAGE)
    if not os.path.isdir(src_pkg):
        return None
    if os.path.isdir(dst_root):
        shutil.rmtree(dst_root)
    digest = hashlib.sha256()
    n = 0
    for rel in (os.path.join("mmpfn", "__init__.py"), os.path.join("mmpfn", "models", "__init__.py")):
        s, d = os.path.join(src_root, rel), os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.exists(s):
            shutil.copyfile(s, d)
        else:
            open(d, "w").close()
    for base, dirs, files in os.walk(src_pkg):
        dirs[:] = sorted(x for x in dirs if x != "__pycache__")
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            s = os.path.join(base, f)
            rel = os.path.relpath(s, src_root)
            d = os.path.join(dst_root, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            with open(s, "rb") as fh:
                digest.update(rel.encode())
                digest.update(fh.read())
            n += 1
    with open(os.path.join(dst_root, "SNAPSHOT.txt"), "w") as fh:
        fh.write(f"unmodified copy of {src_pkg} ({n} files), sha256 {digest.hexdigest()}\n"
                 "test/baseline infrastructure only; git-ignored; made by oracle/snapshot_ref.py\n")
    return dst_root


if __name__ == "__main__":
    out = snapshot()
    print(out if out else f"no reference under {SRC_ROOT}", file=sys.stderr if out is None else sys.stdout)
