"""The one constant of the FP32 filter's error bound that is not arithmetic: the accuracy of its
sin/cos (acrobotics_b200/csrc/acro_filter.cuh: sincos_filter = FP64 range reduction, FP32
rounding of the reduced angle, Cephes minimax polynomials on [-pi/4, pi/4]).
derive_filter_tolerance (csrc/abi.cu) uses e_t = 1.4e-7; this test re-measures the error of the
same operation sequence (FMA emulated exactly in double precision) over 8 M arguments."""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _mul(a, b):
    return (a.astype(np.float64) * b.astype(np.float64)).astype(f32)


def test_filter_sincos_error_is_below_the_bound_constant():
    rng = np.random.default_rng(0)
    worst_s = worst_c = 0.0
    for _ in range(4):
        r = rng.uniform(-np.pi / 4 * (1 + 1e-7), np.pi / 4 * (1 + 1e-7), 2_000_000)  # reduced angle (FP64)
        x = r.astype(f32)
        z = _mul(x, x)
        c = lambda v: np.full_like(x, v, dtype=f32)  # noqa: E731
        sp = _fma(_fma(c(-1.9515295891e-4), z, c(8.3321608736e-3)), z, c(-1.6666654611e-1))
        sp = _fma(_mul(sp, z), x, x)
        cp = _fma(_fma(c(2.443315711809948e-5), z, c(-1.388731625493765e-3)), z, c(4.166664568298827e-2))
        cp = _fma(_mul(cp, z), z, _fma(c(-0.5), z, c(1.0)))
        worst_s = max(worst_s, float(np.abs(sp.astype(np.float64) - np.sin(r)).max()))
        worst_c = max(worst_c, float(np.abs(cp.astype(np.float64) - np.cos(r)).max()))
    print(f"[filter-bound] sincos_filter: max |sin error| {worst_s:.3e}, max |cos error| {worst_c:.3e}; "
          "bound constant e_t = 1.4e-7")
    assert worst_s < 1.2e-7 and worst_c < 1.2e-7


def test_range_reduction_in_fp64_is_exact_enough():
    """q - k pi/2 as sincos_filter computes it (k by the magic-number trick, one FMA with the
    FP64 value of pi/2): the error is |k| * 6.1e-17 < 2.7e-9 up to |q| = 2^26, far below e_t."""
    from fractions import Fraction

    def fma(a, b, c):  # exact product and sum, one rounding (int / int division rounds correctly)
        v = Fraction(a) * Fraction(b) + Fraction(c)
        return v.numerator / v.denominator

    hi, magic = 1.57079632679489655800e+00, 6755399441055744.0
    rng = np.random.default_rng(1)
    qs = np.concatenate([rng.uniform(-2.0 ** 26, 2.0 ** 26, 3000), rng.uniform(-10, 10, 3000)])
    half_pi = Fraction("1.57079632679489661923132169163975144209858469968755291048747")
    worst = worst_r = 0.0
    for q in qs:
        tm = fma(float(q), 0.63661977236758134308, magic)
        kd = tm - magic
        k = int(kd)
        assert kd == k and abs(k - float(q) * 0.63661977236758134308) <= 0.5 + 1e-6
        r = fma(-kd, hi, float(q))
        worst = max(worst, abs(float(Fraction(r) - (Fraction(float(q)) - k * half_pi))))
        worst_r = max(worst_r, abs(r))
    assert worst < 2.7e-9 and worst_r < np.pi / 4 * (1 + 1e-7)
