"""The device build of fb_glibc_math.h against the host libm, bit for bit, through the C ABI (fb_devmath_eval)."""
import ctypes as C

import numpy as np
import pytest

import glibc_math_lib as GM

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("what,name", [(0, "log"), (1, "pow10")])
def test_device_math_equals_host_libm(built, what, name):
    if not GM.host_has_fma():
        pytest.skip("host CPU without FMA+AVX2: its glibc dispatches to another log/pow variant")
    from gapprober_b200.figbird import device_math
    n = 12_000_000
    x = np.empty(n)
    dp = C.POINTER(C.c_double)
    GM.lib().gm_inputs(what, x.ctypes.data_as(dp), n, 97531 + what)
    want = np.empty(n)
    GM.lib().gm_libm(what, x.ctypes.data_as(dp), want.ctypes.data_as(dp), n)
    got = device_math(name, x)
    same = (got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))
    bad = np.flatnonzero(~same)
    assert bad.size == 0, [(float(x[i]).hex(), float(want[i]).hex(), float(got[i]).hex()) for i in bad[:5]]
