"""
The fp64 division / reciprocal / square-root sequences the trace kernels use instead of the compiler's
(``csrc/common.cuh``: MUFU seed + one third-order step, no branches), measured through ``optk_debug_math``
against correctly rounded results: within 1 ulp over twelve decades, IEEE special values as the plain operators.
"""

import numpy as np
import pytest
import torch

from optika_b200 import _lib

pytestmark = pytest.mark.gpu

N = 1 << 21


def run(op, a, b=None):
    lib = _lib.lib()
    ta = torch.as_tensor(a, dtype=torch.float64, device="cuda")
    tb = torch.as_tensor(b, dtype=torch.float64, device="cuda") if b is not None else None
    out = torch.empty_like(ta)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.optk_debug_math(op, ta.numel(), ta.data_ptr(), tb.data_ptr() if tb is not None else None, out.data_ptr(), stream))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def ulps(got, want_long):
    """|got - want| in units of the last place of want (want in extended precision)."""
    want = want_long.astype(np.float64)
    return np.abs((got.astype(np.longdouble) - want_long) / np.spacing(np.abs(want)).astype(np.longdouble)).astype(float)


def operands(seed, low=-6, high=6, signed=True):
    rng = np.random.default_rng(seed)
    x = 10.0 ** rng.uniform(low, high, N) * rng.uniform(1.0, 2.0, N)
    if signed:
        x *= rng.choice([-1.0, 1.0], N)
    return x


def test_division_within_one_ulp(cuda_device):
    a, b = operands(1), operands(2)
    err = ulps(run(0, a, b), a.astype(np.longdouble) / b.astype(np.longdouble))
    assert err.max() <= 1.0, err.max()
    assert np.mean(err <= 0.5) > 0.95  # almost always the correctly rounded quotient


def test_division_by_a_finite_divisor(cuda_device):
    """fdiv_finite: no residual step (<= 1.5 ulp); b = 0, a = inf, NaN and 0 / 0 as IEEE without a repair."""
    a, b = operands(6), operands(7)
    err = ulps(run(6, a, b), a.astype(np.longdouble) / b.astype(np.longdouble))
    assert err.max() <= 1.5, err.max()
    inf, nan = np.inf, np.nan
    a = np.array([1.0, -1.0, 1.0, -1.0, 0.0, inf, -inf, nan, 1.0, 0.0, 5.0])
    b = np.array([0.0, 0.0, -0.0, -0.0, 0.0, 2.0, 3.0, 1.0, nan, 4.0, -2.5])
    with np.errstate(all="ignore"):
        want = a / b
    got = run(6, a, b)
    assert np.array_equal(got, want, equal_nan=True), (got, want)


def test_division_for_newton_steps(cuda_device):
    """fdiv_newton: one Newton step on the seed, relative error below 2^-36."""
    a, b = operands(8), operands(9)
    got = run(7, a, b)
    assert np.max(np.abs(got / (a / b) - 1)) < 2.0 ** -36


@pytest.mark.parametrize("op", [1, 4], ids=["repaired", "raw"])
def test_reciprocal_within_one_ulp(cuda_device, op):
    a = operands(3)
    err = ulps(run(op, a), 1 / a.astype(np.longdouble))
    assert err.max() <= 1.0, err.max()


def test_square_root_within_one_ulp(cuda_device):
    a = operands(4, signed=False)
    err = ulps(run(2, a), np.sqrt(a.astype(np.longdouble)))
    assert err.max() <= 1.0, err.max()


@pytest.mark.parametrize("op", [3, 5], ids=["repaired", "raw"])
def test_reciprocal_square_root_within_one_ulp(cuda_device, op):
    a = operands(5, signed=False)
    err = ulps(run(op, a), 1 / np.sqrt(a.astype(np.longdouble)))
    assert err.max() <= 1.0, err.max()
    # the worst seeds sit just below a power of four; sweep one binade densely as well
    dense = np.linspace(1.0, 4.0, N, endpoint=False)
    err = ulps(run(op, dense), 1 / np.sqrt(dense.astype(np.longdouble)))
    assert err.max() <= 1.0, err.max()


def test_special_values_follow_ieee(cuda_device):
    inf, nan = np.inf, np.nan
    a = np.array([1.0, -1.0, 0.0, -0.0, inf, -inf, nan, 1.0, 0.0, inf, 3.0, 0.0])
    b = np.array([0.0, 0.0, 1.0, 2.0, 2.0, -3.0, 1.0, nan, 0.0, inf, inf, -0.0])
    with np.errstate(all="ignore"):
        want = a / b
    got = run(0, a, b)
    assert np.array_equal(got, want, equal_nan=True)
    sign_matters = ~np.isnan(want) & (want != 0)  # (-0) / 2 comes back as +0: the residual step adds +0; nothing downstream looks
    assert np.array_equal(np.signbit(got[sign_matters]), np.signbit(want[sign_matters]))
    x = np.array([0.0, -0.0, inf, -inf, nan, 4.0, -4.0])
    with np.errstate(all="ignore"):
        assert np.array_equal(run(1, x), 1 / x, equal_nan=True)
        assert np.array_equal(run(2, x), np.sqrt(x), equal_nan=True)
        want = 1 / np.sqrt(x)
    got = run(3, x)
    assert np.array_equal(got[[0, 2, 4, 5, 6]], want[[0, 2, 4, 5, 6]], equal_nan=True)  # -0 -> -inf in IEEE; the kernels never ask
    # the unrepaired variants give NaN where the repaired ones give inf / 0: documented, and only used where that is a NaN downstream anyway
    raw = run(5, np.array([0.0, inf, 4.0, nan, -1.0]))
    assert np.isnan(raw[[0, 1, 3, 4]]).all() and raw[2] == 0.5
