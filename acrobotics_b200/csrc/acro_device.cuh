// acro_device.cuh - device-side data layout and math shared by every kernel of
// libacro_b200 (sm_100a).  Internal; the public C ABI is include/acro_b200.h.
//
// Design notes:
//  * one thread = one joint configuration; the 3x4 running transform lives in
//    registers, the DH chain is unrolled at compile time (template NDOF);
//  * the robot descriptor travels as a __grid_constant__ kernel parameter, so it
//    sits in the constant bank, is read with warp-uniform LDC, and needs no
//    global mutable state (handles are immutable => thread safe);
//  * scene boxes are staged once per CTA into shared memory (uniform broadcast
//    reads in the cull loop, per-lane reads in the SAT loop).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/acro_b200.h"
#include "acro_math.cuh"

namespace acro {

constexpr int kMaxDof = ACRO_MAX_DOF;
constexpr int kMaxBoxes = ACRO_MAX_ROBOT_BOXES;
constexpr int kMaxPairs = ACRO_MAX_SELF_PAIRS;
constexpr int kMaxScene = ACRO_MAX_SCENE_BOXES;
constexpr int kMaxSaved = 16;    // robot boxes that are the earlier member of a self pair
constexpr int kSceneStride = 20; // reals per scene box in the packed layout

// ---------------------------------------------------------------------------
// packed descriptors
// ---------------------------------------------------------------------------
template <typename real>
struct BoxDev {
  real half[3];
  real radius;   // |half|: bounding-sphere radius
  real off[12];  // 3x4 offset in the carrier frame
  int32_t save_slot;     // >= 0: world pose is kept for later self pairs
  int32_t rot_identity;  // offset is a pure translation
  int32_t cull_class;    // index of this box's bounding radius among the robot's distinct radii
  int32_t reserved;
};

// What the FP32 filter reads per robot box, packed into 32 bytes (two 16-byte constant loads
// per box instead of a dozen scalar ones; always floats).  meta: bits 0-3 slot, 4-8 cull_class,
// 9-13 save_slot + 1, 14 rot_identity, 16-23 first self pair, 24-31 number of self pairs.
struct alignas(16) BoxHot {
  float off[3];  // translation of the box in its carrier frame
  float radius;
  float half[3];
  uint32_t meta;
};

template <typename real>
struct RobotDev {
  int32_t ndof;
  int32_t prismatic_mask;
  int32_t has_tool_tip;
  int32_t check_self;
  int32_t n_boxes;
  int32_t base_identity;
  // boxes are sorted by processing order: slot 0 = base boxes, slots 1..ndof =
  // link boxes, slot ndof+1 = tool boxes (carried by the last link frame)
  int32_t slot_begin[kMaxDof + 3];
  real a[kMaxDof], d[kMaxDof], theta[kMaxDof], ca[kMaxDof], sa[kMaxDof];
  real base[12];
  real tip[12];
  BoxDev<real> boxes[kMaxBoxes];
  // self pairs, grouped by their LATER box: partners of box b are
  // partner[pair_begin[b] .. pair_begin[b+1]) and hold the earlier box index
  uint8_t pair_begin[kMaxBoxes + 1];
  uint8_t partner[kMaxPairs];
  // the filter's view of the same data (see BoxHot); partner_slot[k] = save_slot of partner[k]
  BoxHot hot[kMaxBoxes];
  uint8_t partner_slot[kMaxPairs];
  // (cos alpha, sin alpha, a, d) of every joint as one 16-byte constant load (floats)
  float4 dh4[kMaxDof];
};

// Scene box, packed as kSceneStride reals:
//  [0..8] R2 row-major, [9..11] T2, [12..14] half extents B, [15] radius,
//  [16] max half extent, [17..19] unused
template <typename real>
struct SceneView {
  const real* boxes;  // device pointer, n * kSceneStride
  int32_t n;
};

// FP32 cull table of a scene (kernel parameter of the wavefront K2 kernel, read
// with warp-uniform constant loads).  Entry j describes scene box j in its own frame:
//  e[4k..4k+3], k = 0..2 : column k of R2 and t'_k = -(R2^T T2)_k, so that a point c has
//                          box-frame coordinate y_k = col_k . c + t'_k
//  e[12..14]             : B_k (1 + 1e-6) + 1e-6 |t'|_1, rounded up (padded half extents)
//  e[15]                 : unused
struct CullTable {
  float e[kMaxScene][16];
};

// Candidate grid of a (robot, scene) pair (kernel parameter of the filtered K2 kernel).
// A cube of dim^3 cells around the robot base; cells[(cls * dim^3) + (iz*dim + iy)*dim + ix]
// is a bit set over the scene boxes: bit j is CLEAR only if, for every point c of the
// (slightly enlarged) cell, a sphere of the class radius around c is separated from scene
// box j along one of that box's face axes - i.e. FCL's own face-axis test separates every
// robot box of that radius class centred anywhere in the cell.
struct GridView {
  const void* cells;  // uint32 words when the scene has <= 32 boxes, else uint64
  float ox, oy, oz;   // lower corner
  float inv_cell;
  int32_t dim;
  float delta;        // FP32 filter tolerance (bound on |s_fp32 - s_fp64| of any SAT quantity)
  float q_lin_max;    // prismatic joint travel covered by the bound
  int32_t skip_nan;   // rows holding NaN are absent IK branches: not evaluated, verdict bit 0
};

// bounding radii of the robot's radius classes (kernel parameter of the grid builder)
struct GridClassRadii {
  double r[kMaxBoxes];
};

// ---------------------------------------------------------------------------
// small math helpers
// ---------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ void sincos_t(real x, real* s, real* c);
template <>
__device__ __forceinline__ void sincos_t<double>(double x, double* s, double* c) {
  sincos_f64(x, s, c);  // acro_math.cuh
}
// FP32 chain (poses good to 1e-5): Cody-Waite reduction by pi/2 in three exact steps and the
// Cephes minimax polynomials on [-pi/4, pi/4] (abs error ~1.5e-7) for |x| <= 1024; the library
// routine beyond that.  Half the instructions of sincosf and no slow-path stack frame.
template <>
__device__ __forceinline__ void sincos_t<float>(float x, float* s, float* c) {
  if (!(fabsf(x) <= 1024.0f)) {
    sincosf(x, s, c);
    return;
  }
  const float t = fmaf(x, 0.636619772f, 12582912.0f);  // 1.5 * 2^23: the integer lands in the mantissa
  const int k = __float_as_int(t);
  const float kf = t - 12582912.0f;
  float r = fmaf(kf, -1.5703125f, x);
  r = fmaf(kf, -4.837512969970703125e-4f, r);
  r = fmaf(kf, -7.54978995489188216e-8f, r);
  const float z = r * r;
  float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
  sp = fmaf(sp * z, r, r);
  float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
  cp = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
  const float a = (k & 1) ? cp : sp;
  const float b = (k & 1) ? sp : cp;
  *s = (k & 2) ? -a : a;
  *c = ((k + 1) & 2) ? -b : b;
}

template <typename real>
__device__ __forceinline__ real fma_t(real a, real b, real c);
template <>
__device__ __forceinline__ double fma_t<double>(double a, double b, double c) {
  return fma(a, b, c);
}
template <>
__device__ __forceinline__ float fma_t<float>(float a, float b, float c) {
  return fmaf(a, b, c);
}

template <typename real>
__device__ __forceinline__ real abs_t(real x) {
  return x < real(0) ? -x : x;
}
template <>
__device__ __forceinline__ double abs_t<double>(double x) {
  return fabs(x);
}
template <>
__device__ __forceinline__ float abs_t<float>(float x) {
  return fabsf(x);
}

template <typename real>
struct Eps;
template <>
struct Eps<double> {
  static constexpr double v = 2.220446049250313e-16;
};
template <>
struct Eps<float> {
  static constexpr float v = 1.1920929e-07f;
};

// 3x4 rigid transform in registers: R row-major + p
template <typename real>
struct Tf {
  real r[9];
  real p[3];
};

template <typename real>
__device__ __forceinline__ void tf_load(Tf<real>& T, const real* m) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    T.r[3 * r + 0] = m[4 * r + 0];
    T.r[3 * r + 1] = m[4 * r + 1];
    T.r[3 * r + 2] = m[4 * r + 2];
    T.p[r] = m[4 * r + 3];
  }
}

// T <- T * DH(ct, st, ca, sa, a, d)      (reference link.py:35-58, robot.py:64-66)
// Factored form: u = R[:,0]ct + R[:,1]st ; w = R[:,1]ct - R[:,0]st ;
//   col0 = u ; col1 = w ca + R[:,2] sa ; col2 = R[:,2] ca - w sa ;
//   p += a u + d R[:,2]          (30 FMA-class instructions per joint)
template <typename real>
__device__ __forceinline__ void tf_mul_dh(Tf<real>& T, real ct, real st, real ca, real sa, real a,
                                          real d) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const real r0 = T.r[3 * r + 0], r1 = T.r[3 * r + 1], r2 = T.r[3 * r + 2];
    const real u = fma_t(r1, st, r0 * ct);
    const real w = fma_t(r1, ct, -(r0 * st));
    T.r[3 * r + 0] = u;
    T.r[3 * r + 1] = fma_t(r2, sa, w * ca);
    T.r[3 * r + 2] = fma_t(r2, ca, -(w * sa));
    T.p[r] = fma_t(d, r2, fma_t(a, u, T.p[r]));
  }
}

// Packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2: two FP32 lanes per instruction, one operand may
// be a scalar broadcast).  The helpers keep the operation order of the scalar code, so results
// are bit-identical to it.
#ifndef ACRO_F32X2
#define ACRO_F32X2 1
#endif
__device__ __forceinline__ float2 mul2(float s, float2 v) { return __fmul2_rn(make_float2(s, s), v); }
__device__ __forceinline__ float2 fma2(float s, float2 v, float2 acc) {
  return __ffma2_rn(make_float2(s, s), v, acc);
}

#if ACRO_F32X2
// rows 0 and 1 of every column as one packed lane pair, row 2 scalar: 20 instead of 30 issue slots
template <>
__device__ __forceinline__ void tf_mul_dh<float>(Tf<float>& T, float ct, float st, float ca,
                                                 float sa, float a, float d) {
  const float2 c0 = make_float2(T.r[0], T.r[3]), c1 = make_float2(T.r[1], T.r[4]),
               c2 = make_float2(T.r[2], T.r[5]);
  const float2 u = fma2(st, c1, mul2(ct, c0));
  const float2 w = fma2(ct, c1, mul2(-st, c0));
  const float2 n1 = fma2(sa, c2, mul2(ca, w));
  const float2 n2 = fma2(ca, c2, mul2(-sa, w));
  const float2 pn = fma2(d, c2, fma2(a, u, make_float2(T.p[0], T.p[1])));
  T.r[0] = u.x; T.r[3] = u.y;
  T.r[1] = n1.x; T.r[4] = n1.y;
  T.r[2] = n2.x; T.r[5] = n2.y;
  T.p[0] = pn.x; T.p[1] = pn.y;
  {
    const float r0 = T.r[6], r1 = T.r[7], r2 = T.r[8];
    const float us = fmaf(r1, st, r0 * ct);
    const float ws = fmaf(r1, ct, -(r0 * st));
    T.r[6] = us;
    T.r[7] = fmaf(r2, sa, ws * ca);
    T.r[8] = fmaf(r2, ca, -(ws * sa));
    T.p[2] = fmaf(d, r2, fmaf(a, us, T.p[2]));
  }
}
#endif

// out = T * M  (M is a 3x4 constant from the descriptor)
template <typename real>
__device__ __forceinline__ void tf_mul(Tf<real>& out, const Tf<real>& T, const real* m) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const real r0 = T.r[3 * r + 0], r1 = T.r[3 * r + 1], r2 = T.r[3 * r + 2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out.r[3 * r + c] = fma_t(r2, m[8 + c], fma_t(r1, m[4 + c], r0 * m[c]));
    out.p[r] = fma_t(r2, m[11], fma_t(r1, m[7], fma_t(r0, m[3], T.p[r])));
  }
}

// joint i of the chain: revolute theta = q, prismatic d = q   (link.py:40-44)
template <typename real>
__device__ __forceinline__ void chain_step(Tf<real>& T, const RobotDev<real>& rb, int i, real qi) {
  real th, dd;
  if ((rb.prismatic_mask >> i) & 1) {
    th = rb.theta[i];
    dd = qi;
  } else {
    th = qi;
    dd = rb.d[i];
  }
  real s, c;
  sincos_t<real>(th, &s, &c);
  tf_mul_dh<real>(T, c, s, rb.ca[i], rb.sa[i], rb.a[i], dd);
}

// ---------------------------------------------------------------------------
// coalesced row access: a CTA moves `rows` consecutive rows of `W` reals between
// global memory and shared memory as one flat, fully coalesced block.
// ---------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ void cta_load_rows(real* smem, const real* __restrict__ g,
                                              int64_t row0, int64_t n_rows_total, int rows,
                                              int W) {
  const int64_t base = row0 * W;
  int64_t avail = (n_rows_total - row0) * (int64_t)W;
  const int count = (int)(avail < (int64_t)rows * W ? (avail < 0 ? 0 : avail) : (int64_t)rows * W);
  for (int k = threadIdx.x; k < count; k += blockDim.x) smem[k] = g[base + k];
}

template <typename real>
__device__ __forceinline__ void cta_store_rows(real* __restrict__ g, const real* smem,
                                               int64_t row0, int64_t n_rows_total, int rows,
                                               int W) {
  const int64_t base = row0 * W;
  int64_t avail = (n_rows_total - row0) * (int64_t)W;
  const int count = (int)(avail < (int64_t)rows * W ? (avail < 0 ? 0 : avail) : (int64_t)rows * W);
  for (int k = threadIdx.x; k < count; k += blockDim.x) g[base + k] = smem[k];
}

}  // namespace acro
