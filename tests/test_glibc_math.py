"""The restatement of glibc's log() and pow(10, .) that the device code compiles (gapprober_b200/csrc/fb_glibc_math.h)
against the host libm itself — the reference's libm — bit for bit, on the CPU: 10^7 generated inputs per function
(probabilities down to the subnormal range, the near-1 branch, log scores down to and beyond underflow, NaN/inf) plus
explicit edge values.  Also: the committed tables are the ones in this host's libm.so.6."""
import os
import subprocess
import sys

import pytest

import glibc_math_lib as GM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_fma = pytest.mark.skipif(not GM.host_has_fma(), reason="host CPU without FMA+AVX2: its glibc dispatches to another log/pow variant")


@needs_fma
@pytest.mark.parametrize("what,name", [(0, "log"), (1, "pow10")])
def test_restatement_equals_host_libm(what, name):
    assert GM.lib().gm_check(what, 10_000_000, 1234567 + what) == 0, name


@needs_fma
def test_edge_values():
    assert GM.lib().gm_check_edges() == 0


def test_committed_tables_are_this_libms(tmp_path):
    """tools/extract_libm_tables.py run again on the libm of this machine reproduces the committed header."""
    libm = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm):
        pytest.skip("no x86-64 glibc here")
    hdr = os.path.join(ROOT, "gapprober_b200", "csrc", "fb_glibc_tables.h")
    before = open(hdr).read()
    try:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "extract_libm_tables.py"), libm], stdout=subprocess.DEVNULL)
        assert open(hdr).read() == before
    finally:
        open(hdr, "w").write(before)
