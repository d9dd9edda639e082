"""Accuracy of the lean FP64 elementary functions (csrc/acro_math.cuh) that replace the CUDA
library's sincos / atan2 / sqrt / division inside the kernels: a few ulp against numpy over the
ranges the path uses, and the library's own answers for the special cases."""
import numpy as np
import pytest
import torch

from acrobotics_b200 import _native, device

pytestmark = pytest.mark.gpu


def probe(fn, a, b=None):
    lib = _native.load()
    a_d = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    b_d = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64)).cuda() if b is not None else None
    o0 = torch.empty_like(a_d)
    o1 = torch.empty_like(a_d)
    _native.check(lib.acro_math_probe_f64(fn, a_d.data_ptr(), b_d.data_ptr() if b_d is not None else None,
                                          o0.data_ptr(), o1.data_ptr(), a_d.numel(),
                                          device.current_stream_ptr()), "acro_math_probe_f64")
    torch.cuda.synchronize()
    return o0.cpu().numpy(), o1.cpu().numpy()


def ulp_err(got, ref):
    """|got - ref| in units of the spacing of ref (inf / nan must agree exactly)."""
    got, ref = np.asarray(got), np.asarray(ref)
    special = ~np.isfinite(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got[special & ~np.isnan(ref)], ref[special & ~np.isnan(ref)])
    fin = ~special
    return np.abs(got[fin] - ref[fin]) / np.spacing(np.maximum(np.abs(ref[fin]), 1e-300))


def _long(x):
    return np.asarray(x, dtype=np.longdouble)


def test_sincos_joint_range_and_beyond():
    rng = np.random.default_rng(1)
    x = np.concatenate([
        rng.uniform(-np.pi, np.pi, 400_000), rng.uniform(-1e3, 1e3, 200_000),
        rng.uniform(-1e8, 1e8, 200_000), rng.uniform(-1e-3, 1e-3, 50_000),
        np.arange(-64, 65) * (np.pi / 2), np.arange(-64, 65) * (np.pi / 4),
        np.nextafter(np.arange(1, 65) * (np.pi / 2), 0), [0.0, -0.0, 1e-300, -1e-310, 2.0 ** 30 - 1],
    ])
    s, c = probe(0, x)
    # the 80-bit libm is the reference (correct to well under a double ulp in this range)
    es = ulp_err(s, np.sin(_long(x)).astype(np.float64))
    ec = ulp_err(c, np.cos(_long(x)).astype(np.float64))
    assert es.max() <= 2.0 and ec.max() <= 2.0, (es.max(), ec.max())
    assert np.abs(s * s + c * c - 1).max() < 1e-15
    assert (s[x == 0] == 0).all()  # (the sign of sin(-0) is not preserved: +0)


def test_sincos_large_and_non_finite_go_to_the_library():
    x = np.array([2.0 ** 30, -2.0 ** 31, 1e15, -3e22, 1e300, np.inf, -np.inf, np.nan])
    s, c = probe(0, x)
    fin = np.isfinite(x)
    assert ulp_err(s[fin], np.sin(x[fin])).max() <= 2.0
    assert ulp_err(c[fin], np.cos(x[fin])).max() <= 2.0
    assert np.isnan(s[~fin]).all() and np.isnan(c[~fin]).all()


def test_atan2_everywhere():
    rng = np.random.default_rng(2)
    n = 500_000
    ang = rng.uniform(-np.pi, np.pi, n)
    rad = 10.0 ** rng.uniform(-12, 12, n)
    y = np.concatenate([rad * np.sin(ang), rng.standard_normal(n), rng.standard_normal(2000) * 1e-9])
    x = np.concatenate([rad * np.cos(ang), rng.standard_normal(n), np.ones(2000)])
    # octant boundaries, the folding threshold tan(pi/8) and the axes
    t = np.tan(np.pi / 8)
    by = np.array([1, 1, -1, -1, t, t, -t, 1, 1, 0.0, -0.0, 0.0, -0.0, 1, -1, 1, -1, 3e-320, 1.0])
    bx = np.array([1, -1, 1, -1, 1, -1, 1, t, -t, 1, 1, -1, -1, 0.0, 0.0, -0.0, -0.0, 1.0, 3e-320])
    y, x = np.concatenate([y, by, np.nextafter(by, 2)]), np.concatenate([x, bx, bx])
    got, _ = probe(1, y, x)
    ref = np.arctan2(_long(y), _long(x)).astype(np.float64)
    e = ulp_err(got, ref)
    assert e.max() <= 3.0, e.max()
    assert np.array_equal(np.signbit(got), np.signbit(ref))


def test_atan2_special_cases_match_ieee():
    inf, nan = np.inf, np.nan
    y = np.array([0.0, -0.0, 0.0, -0.0, inf, -inf, inf, inf, 1.0, 1.0, nan, 1.0, 1e308, 1e-308, 1e308, 0.0])
    x = np.array([0.0, 0.0, -0.0, -0.0, 1.0, 1.0, inf, -inf, inf, -inf, 1.0, nan, 1e308, 1e-308, 1e-308, 1e-320])
    got, _ = probe(1, y, x)
    ref = np.arctan2(y, x)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert ulp_err(got[ok], ref[ok]).max() <= 2.0
    assert np.array_equal(np.signbit(got[ok]), np.signbit(ref[ok]))


def test_sqrt_div_rcp():
    rng = np.random.default_rng(3)
    a = 10.0 ** rng.uniform(-100, 100, 300_000)
    b = 10.0 ** rng.uniform(-100, 100, 300_000) * rng.choice([-1.0, 1.0], 300_000)
    got, _ = probe(2, np.concatenate([a, [0.0, 1.0, 4.0, 1e-310, 1e300, np.inf]]))
    ref = np.sqrt(np.concatenate([a, [0.0, 1.0, 4.0, 1e-310, 1e300, np.inf]]))
    assert ulp_err(got, ref).max() <= 1.0
    assert (got[-6:-3] == [0.0, 1.0, 2.0]).all()
    num = a * rng.choice([-1.0, 1.0], a.size)
    got, _ = probe(3, num, b)
    assert ulp_err(got, (_long(num) / _long(b)).astype(np.float64)).max() <= 1.0
    got, _ = probe(4, b)
    assert ulp_err(got, (1 / _long(b)).astype(np.float64)).max() <= 1.0
    got, _ = probe(2, np.array([-1.0, np.nan]))
    assert np.isnan(got).all()
