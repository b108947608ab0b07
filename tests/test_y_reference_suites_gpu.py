"""gpu: the reference's OWN regression programs (libmspack/test/cabd_test.c, chmd_test.c) compiled from the reference
sources where they lie together with the reference's container parsers (cabd.c / chmd.c / system.c), but with
lzxd_* / qtmd_* / mszipd_* coming from libmspack_dropin.so, i.e. from the GPU kernels (oracle/Makefile targets
cabd_gpu / chmd_gpu; the binaries are built in this container and travel to the GPU box).  This is the drop-in
boundary of SURVEY.md 8(b) exercised by the reference's own tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("binary,expect", [("cabd_test_gpu", "ALL 433 TESTS PASSED"), ("chmd_test_gpu", "ALL 195 TESTS PASSED")])
def test_reference_test_program_passes_on_the_dropin(binary, expect):
    exe = os.path.join(ROOT, "oracle", "_ref", binary)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, timeout=900, cwd=ROOT)   # fixture paths are relative to the repo root
    out = (r.stdout + r.stderr).decode("utf-8", "replace")                   # (the programs print raw member names)
    assert r.returncode == 0, out[-2000:]
    assert expect in out, out[-2000:]
