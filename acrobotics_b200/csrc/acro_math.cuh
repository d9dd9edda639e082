// acro_math.cuh - lean FP64 sincos / atan2 / reciprocal / sqrt for the streaming kernels.
//
// The FP64 kernels of the path (K1 FK, K3/K4 Jacobians, K5 IK, the exact pass of K2) are bound
// by instruction issue, not by the FP64 pipe: the CUDA library's sincos / atan2 materialise
// every 64-bit polynomial coefficient with two UMOV / IMAD.MOV instructions at each use (171
// UMOV in the six-joint FK kernel), round with an F2I / I2F pair and carry a slow-path call.
// Here the coefficients sit in constant memory (two per LDCU.128 / one per LDC.64, then a
// uniform-register or register operand of DFMA), the quadrant comes from a magic-number add,
// atan2 does one division with the octant folded into it, and the rarely taken cases
// (|x| >= 2^30, zero / non-finite / extreme atan2 arguments) branch to the library.
// One deviation from IEEE conventions: sin(-0) returns +0.
//
// Accuracy (measured against the library and numpy on the GPU, tests/test_gpu_math.py):
// a few ulp (< 1e-15 relative), far inside the 1e-9 parity tolerance of SURVEY.md 8c.
// Coefficients: Chebyshev-node interpolants derived with mpmath (tools/lean_math_coefficients.py),
// |error| < 1e-17 on the reduced ranges.
#pragma once
#include <cuda_runtime.h>

#ifndef ACRO_LEAN_MATH
#define ACRO_LEAN_MATH 1
#endif

namespace acro {

// pi/2 in three pieces (53 + 42 + 53 bits), 2/pi, then the polynomials in z = r^2
static __constant__ double kLeanTrig[] = {
    1.5707963267948966,       // 0 pio2 hi
    6.123233995735495e-17,    // 1 pio2 mid
    1.2706558760139879e-29,   // 2 pio2 lo
    0.6366197723675814,       // 3 2/pi
    // sin(r) = r + r z (S1 + z (S2 + ...)), |r| <= pi/4
    -0.16666666666666616,     // 4
    0.008333333333320297,     // 5
    -0.0001984126982859984,   // 6
    2.755731335888172e-06,    // 7
    -2.5050714198795763e-08,  // 8
    1.5894573136298841e-10,   // 9
    // cos(r) = 1 + z (-1/2 + z (C2 + z (C3 + ...)))
    0.04166666666666645,      // 10
    -0.0013888888888860956,   // 11
    2.4801587283810357e-05,   // 12
    -2.7557313079436857e-07,  // 13
    2.087558003775504e-09,    // 14
    -1.1353262834133352e-11,  // 15
};

// atan(t) = t + t z (A1 + z (A2 + ...)), z = t^2, |t| <= tan(pi/8)
static __constant__ double kLeanAtan[] = {
    -0.33333333333333326, 0.19999999999997012,  -0.1428571428532936,  0.11111111085301238,
    -0.0909090805829841,  0.07692281083678901,  -0.06666205023178091, 0.058768317098427496,
    -0.05217297919655059, 0.044994415067496046, -0.03338645225903317, 0.01511237643149517,
    0.7853981633974483,      // 12 pi/4 hi
    3.061616997868383e-17,   // 13 pi/4 lo
    1.5707963267948966,      // 14 pi/2 hi
    6.123233995736766e-17,   // 15 pi/2 lo
    3.141592653589793,       // 16 pi hi
    1.2246467991473532e-16,  // 17 pi lo
};

__device__ __forceinline__ void sincos_lean(double x, double* s, double* c) {
  // |x| < 2^30: the quadrant index fits an int with room to spare and the three-step
  // reduction is exact in its first step (NaN / Inf have a larger exponent field)
  if (((unsigned)__double2hiint(x) & 0x7fffffffu) >= 0x41d00000u) {
    sincos(x, s, c);
    return;
  }
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: the integer lands in the low word
  const double t = fma(x, kLeanTrig[3], magic);
  const int k = __double2loint(t);
  const double kd = t - magic;
  double r = fma(-kd, kLeanTrig[0], x);
  r = fma(-kd, kLeanTrig[1], r);
  r = fma(-kd, kLeanTrig[2], r);
  const double z = r * r;
  double ps = fma(z, kLeanTrig[9], kLeanTrig[8]);
  double pc = fma(z, kLeanTrig[15], kLeanTrig[14]);
  ps = fma(z, ps, kLeanTrig[7]);
  pc = fma(z, pc, kLeanTrig[13]);
  ps = fma(z, ps, kLeanTrig[6]);
  pc = fma(z, pc, kLeanTrig[12]);
  ps = fma(z, ps, kLeanTrig[5]);
  pc = fma(z, pc, kLeanTrig[11]);
  ps = fma(z, ps, kLeanTrig[4]);
  pc = fma(z, pc, kLeanTrig[10]);
  const double rz = r * z;
  pc = fma(z, pc, -0.5);
  const double sv = fma(rz, ps, r);
  const double cv = fma(z, pc, 1.0);
  // quadrant: sin = {s, c, -s, -c}[k & 3], cos = {c, -s, -c, s}[k & 3]
  const bool odd = k & 1;
  const double a = odd ? cv : sv;
  const double b = odd ? sv : cv;
  const int sa = (k & 2) << 30, sb = ((k + 1) & 2) << 30;
  *s = __hiloint2double(__double2hiint(a) ^ sa, __double2loint(a));
  *c = __hiloint2double(__double2hiint(b) ^ sb, __double2loint(b));
}

// 1 / d for a normal d well inside the exponent range (callers guarantee it): hardware
// estimate (about 20 bits) refined to full precision, no special-case path
__device__ __forceinline__ double rcp_lean(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  e = fma(e, e, e);
  r = fma(r, e, r);  // ~60 bits before rounding
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
}

// n / d, correctly rounded up to an ulp, same preconditions on d.  The quotient estimate is
// refined alongside the reciprocal (five dependent operations behind the hardware estimate).
__device__ __forceinline__ double div_lean(double n, double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  const double q0 = n * r;
  e = fma(e, e, e);
  const double q = fma(q0, e, q0);
  r = fma(r, e, r);
  const double rem = fma(-d, q, n);
  return fma(rem, r, q);
}

// atan2 without data-dependent branches on the common path: the octant is folded into one
// division and one affine map of the reduced angle, selections act on 0 / 1 / 2 factors.
__device__ __forceinline__ double atan2_lean(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool steep = ay > ax;  // false when either is NaN: a NaN |x| then stays in mx
  const double mx = steep ? ay : ax, mn = steep ? ax : ay;
  // the larger magnitude must be a normal number in [2^-500, 2^500] (zero, subnormal, huge,
  // Inf and NaN go to the library, which also settles the signed-zero conventions); a NaN in
  // mn propagates through the arithmetic below
  const unsigned eh = (unsigned)__double2hiint(mx);
  if (eh - 0x20b00000u >= 0x3e800000u) return atan2(y, x);
  // one division: t = mn / mx, or with the eighth turn folded in, (mn - mx) / (mn + mx)
  const bool fold = mn > 0.41421356237309503 * mx;
  const double f = __hiloint2double(fold ? 0x3ff00000 : 0, 0);  // 1.0 or 0.0
  const double num = fma(-f, mx, mn);
  const double den = fma(f, mn, mx);
  const double t = div_lean(num, den);
  const double z = t * t;
  // atan(t) = t + t z p(z), |t| <= 0.4143; p split into its even and odd halves in w = z^2
  // (two chains of five instead of one of eleven)
  const double w = z * z;
  double pe = fma(w, kLeanAtan[10], kLeanAtan[8]);
  double po = fma(w, kLeanAtan[11], kLeanAtan[9]);
  pe = fma(w, pe, kLeanAtan[6]);
  po = fma(w, po, kLeanAtan[7]);
  pe = fma(w, pe, kLeanAtan[4]);
  po = fma(w, po, kLeanAtan[5]);
  pe = fma(w, pe, kLeanAtan[2]);
  po = fma(w, po, kLeanAtan[3]);
  pe = fma(w, pe, kLeanAtan[0]);
  po = fma(w, po, kLeanAtan[1]);
  const double p = fma(z, po, pe);
  double a = fma(t * z, p, t);
  a = fma(f, kLeanAtan[12], fma(f, kLeanAtan[13], a));  // + pi/4 when folded: a in [0, pi/4]
  // octant: result = m pi/2 + s a with (m, s) = (0, +) | (1, -) steep | (2, -) x < 0 |
  // (1, +) both
  const bool negx = __double2hiint(x) < 0;
  const int flip = (steep != negx) ? (int)0x80000000 : 0;
  const double sa = __hiloint2double(__double2hiint(a) ^ flip, __double2loint(a));
  const double m = __hiloint2double(steep ? 0x3ff00000 : (negx ? 0x40000000 : 0), 0);
  a = fma(m, kLeanAtan[15], fma(m, kLeanAtan[14], sa));
  return __hiloint2double((__double2hiint(a) & 0x7fffffff) | (__double2hiint(y) & 0x80000000),
                          __double2loint(a));
}

// sqrt of a normal, non-negative d in [2^-500, 2^500] (else the library): hardware
// reciprocal-square-root estimate, two coupled refinements, one final correction
__device__ __forceinline__ double sqrt_lean(double d) {
  const unsigned eh = (unsigned)__double2hiint(d);
  if (eh - 0x20b00000u >= 0x3e800000u) return sqrt(d);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double g = d * y;         // ~ sqrt(d)
  double h = 0.5 * y;       // ~ 1 / (2 sqrt(d))
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  const double rem = fma(-g, g, d);
  return fma(rem, h, g);
}

#if ACRO_LEAN_MATH
__device__ __forceinline__ void sincos_f64(double x, double* s, double* c) { sincos_lean(x, s, c); }
__device__ __forceinline__ double atan2_f64(double y, double x) { return atan2_lean(y, x); }
__device__ __forceinline__ double sqrt_f64(double d) { return sqrt_lean(d); }
#else
__device__ __forceinline__ void sincos_f64(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ double atan2_f64(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ double sqrt_f64(double d) { return sqrt(d); }
#endif

}  // namespace acro
