"""The glibc-exact libm ports (mbelib-neo_b200/csrc/mbe_libm.cuh: sincosf, sinf, cosf, exp2f, expf) against the host
libm, which is the reference's libm.  The header is host+device code; this test compiles the host
side (tests/helpers/libm_check.cpp) and sweeps ~1M float bit patterns (stride 4099 over all 2^32; the full sweep,
stride 1, has been run offline with 0 mismatches on an FMA-capable host)."""
import os
import shutil
import subprocess

import pytest

from __graft_entry__ import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_libm_ports_match_host_glibc(tmp_path):
    exe = str(tmp_path / "libm_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-I", os.path.join(ROOT, "mbelib-neo_b200", "csrc"),
                           os.path.join(ROOT, "tests", "helpers", "libm_check.cpp"), "-o", exe])
    out = subprocess.run([exe, "4099"], capture_output=True, text=True)
    lines = dict(l.split() for l in out.stdout.strip().splitlines())
    assert int(lines["tested"]) > 1000000
    # the x86-64 glibc variant selected at run time decides the last few ulps-of-ulps cases: with the FMA variant
    # (every current x86-64 server) the ports are exact; the non-FMA variant differs in <= 34 of 2^32 arguments
    for fn in ("sincosf", "sinf", "cosf", "exp2f", "expf"):
        assert int(lines[fn]) <= 1, (fn, lines)
