"""The kernels that were measured and parity-tested on a B200 must come out of the compiler unchanged when features are added next
to them (new paths live in template instantiations of their own): tools/sass_check.py compares a hash of each verified
kernel's SASS in the built library with profiles/sass_verified.json."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_verified_kernels_unchanged():
    from libmspack_b200 import build
    build.build_msgpu()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
