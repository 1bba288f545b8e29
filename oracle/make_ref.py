#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference, installed here (build container) so that it travels to the GPU
box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored -- like the built .so files).

TEST INFRASTRUCTURE.  Nothing under cosmotransitions_b200/ imports it.  Users: bench.py (`--impl reference` and the
`cpu_baseline` leg: the reference's own CPU path under multiprocessing.Pool), tests/ (live GPU-vs-reference checks,
the real pathDeformation.fullTunneling with the GPU class plugged in as tunneling_class=).

    python oracle/make_ref.py          # pip-installs /root/reference (from a scratch copy) into oracle/_ref

The install is the offline one the bench contract allows: pip --no-index --no-build-isolation --no-deps --target.
setup.py maps examples/ to the sub-package cosmoTransitions.examples, so testModel1.py and fullTunneling.py come along.
finiteT writes its two spline-table cache files next to itself on first import (finiteT.py:153-163,183-193): that
import is done once here so that the files ship and the box needs no write access at import time.
"""
import os
import shutil
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("CTB_REFERENCE", "/root/reference")
DEST = os.path.join(REPO, "oracle", "_ref")


def make(force=False):
    """Returns DEST if the reference is (now) installed there, None if it cannot be (no source tree here)."""
    marker = os.path.join(DEST, "cosmoTransitions", "tunneling1D.py")
    if os.path.exists(marker) and not force:
        return DEST
    if not os.path.isdir(REF_SRC):
        return None
    tmp = tempfile.mkdtemp(prefix="ctref_src_")
    src = os.path.join(tmp, "reference")
    shutil.copytree(REF_SRC, src)          # the build writes egg-info / build dirs into the source tree
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
           "--find-links", "/opt/wheelhouse", "--target", DEST, src]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    shutil.rmtree(tmp, ignore_errors=True)
    if res.returncode or not os.path.exists(marker):
        sys.stderr.write(res.stdout)
        raise RuntimeError("pip install of the reference into oracle/_ref failed")
    # first import: writes finiteT_b.dat.txt / finiteT_f.dat.txt beside finiteT.py
    subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import cosmoTransitions.finiteT" % DEST],
                   check=True)
    return DEST


if __name__ == "__main__":
    d = make(force="--force" in sys.argv)
    print(d if d else "reference source tree not found (%s): oracle/_ref not built" % REF_SRC)
