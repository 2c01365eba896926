"""GPU: the engine's shared-reciprocal fp64 division (lb_d3q19.cuh, DivBy) is bit-identical to IEEE
division wherever its fast path is taken; operands outside the fast-path domain are flagged (and are
then divided with `/` by the kernels)."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 20161129, 0xDEADBEEF])
def test_divby_equals_ieee_division(seed):
    from hybird_b200 import abi
    lib = abi.load_library()
    res = (C.c_uint64 * 3)()
    abi.check(lib.lbGpuSelfTest(1 << 30, seed, C.byref(res)))
    mismatches, checked, flagged = int(res[0]), int(res[1]), int(res[2])
    assert checked > (1 << 29), (checked, flagged)
    assert flagged > 0  # the generator does produce out-of-domain operands
    assert mismatches == 0, "%d of %d quotients differ from IEEE division" % (mismatches, checked)
