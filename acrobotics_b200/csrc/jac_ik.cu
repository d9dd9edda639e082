// jac_ik.cu - K3/K4 geometric Jacobians and K5 analytical inverse kinematics.
//
// Jacobians replace the casadi AD functions of reference
// src/acrobotics/robot.py:157-171 (fk_casadi :129-136, fk_rpy_casadi :138-144).
// IK replaces Kuka.ik (src/acrobotics/robot_examples.py:188-218), arm_2.ik
// (inverse_kinematics/arm_2.py:5-104), spherical_wrist.ik
// (inverse_kinematics/spherical_wrist.py:5-30) and anthropomorphic_arm.ik
// (inverse_kinematics/anthropomorphic_arm.py:5-55).
//
// All HBM bound: rows are staged through shared memory so that global traffic
// is flat and coalesced.
#include <math_constants.h>

#include "acro_collide.cuh"
#include "acro_launch.h"

#ifndef ACRO_JAC_POS_CTAS
#define ACRO_JAC_POS_CTAS 5
#endif
#ifndef ACRO_IK_CTAS
#define ACRO_IK_CTAS 6
#endif
#ifndef ACRO_JAC_RPY_CTAS
#define ACRO_JAC_RPY_CTAS 4
#endif

namespace acro {

// flat coalesced copy of `W` reals per configuration from padded smem rows
template <int W>
__device__ __forceinline__ void flush_rows(double* __restrict__ out, const double* s_rows,
                                           int64_t row0, int64_t n) {
  constexpr int WP = W | 1;
  const int64_t left = n - row0;
  const int rows = left < kTile ? (int)left : kTile;
  for (int k = threadIdx.x; k < rows * W; k += blockDim.x) {
    const int r = k / W, c = k - r * W;
    out[(row0 + r) * W + c] = s_rows[r * WP + c];
  }
}

template <int W>
__device__ __forceinline__ void fetch_rows(double* s_rows, const double* __restrict__ in,
                                           int64_t row0, int64_t n) {
  constexpr int WP = W | 1;
  const int64_t left = n - row0;
  const int rows = left < kTile ? (int)left : kTile;
  for (int k = threadIdx.x; k < rows * W; k += blockDim.x) {
    const int r = k / W, c = k - r * W;
    s_rows[r * WP + c] = in[(row0 + r) * W + c];
  }
}

// ---------------------------------------------------------------------------
// Jacobians
// ---------------------------------------------------------------------------
template <int NDOF, bool RPY>
__global__ void __launch_bounds__(kTile, RPY ? ACRO_JAC_RPY_CTAS : ACRO_JAC_POS_CTAS)
    jac_kernel(const __grid_constant__ RobotDev<double> rb, const double* __restrict__ q,
               int64_t n, int layout, double* __restrict__ J, bool vec_in, bool vec_out) {
  constexpr int ROWS = RPY ? 6 : 3;
  constexpr int W = ROWS * NDOF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t idx = row0 + threadIdx.x;
  double qv[NDOF];
  load_q_direct<double, NDOF>(qv, q, idx, n, layout, vec_in);

  // joint axis z_{i-1} and origin o_{i-1} are read off the frame *before* link i
  double z[NDOF][3], o[NDOF][3];
  Tf<double> T;
  tf_load<double>(T, rb.base);
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    z[i][0] = T.r[2]; z[i][1] = T.r[5]; z[i][2] = T.r[8];
    o[i][0] = T.p[0]; o[i][1] = T.p[1]; o[i][2] = T.p[2];
    chain_step<double>(T, rb, i, qv[i]);
  }
  if (rb.has_tool_tip) {
    Tf<double> E;
    tf_mul<double>(E, T, rb.tip);
    T = E;
  }
  alignas(16) double row[W + (W & 1)];
  // angle terms of robot.py:138-144, differentiated by the chain rule
  const double T00 = T.r[0], T01 = T.r[1], T02 = T.r[2], T12 = T.r[5], T22 = T.r[8];
  const double s = sqrt_f64(T00 * T00 + T01 * T01);
  const double inv_x = 1.0 / (T22 * T22 + T12 * T12);  // r_x = atan2_f64(-T12, T22)
  const double inv_y = 1.0 / (s * s + T02 * T02);      // r_y = atan2_f64( T02, s)
  const double inv_z = 1.0 / (T22 * T22 + T01 * T01);  // r_z = atan2_f64(-T01, T22)
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    const bool prismatic = (rb.prismatic_mask >> i) & 1;
    const double z0 = z[i][0], z1 = z[i][1], z2 = z[i][2];
    if (prismatic) {
      row[0 * NDOF + i] = z0;
      row[1 * NDOF + i] = z1;
      row[2 * NDOF + i] = z2;
      if (RPY) {
        row[3 * NDOF + i] = 0.0;
        row[4 * NDOF + i] = 0.0;
        row[5 * NDOF + i] = 0.0;
      }
    } else {
      const double r0 = T.p[0] - o[i][0], r1 = T.p[1] - o[i][1], r2 = T.p[2] - o[i][2];
      row[0 * NDOF + i] = z1 * r2 - z2 * r1;
      row[1 * NDOF + i] = z2 * r0 - z0 * r2;
      row[2 * NDOF + i] = z0 * r1 - z1 * r0;
      if (RPY) {
        // dR = [z]x R : only the five entries the angle formulas touch
        const double d00 = z1 * T.r[6] - z2 * T.r[3];
        const double d01 = z1 * T.r[7] - z2 * T.r[4];
        const double d02 = z1 * T.r[8] - z2 * T.r[5];
        const double d12 = z2 * T.r[2] - z0 * T.r[8];
        const double d22 = z0 * T.r[5] - z1 * T.r[2];
        const double ds = (T00 * d00 + T01 * d01) / s;
        // d atan2_f64(y, x) = (x dy - y dx) / (x^2 + y^2)
        row[3 * NDOF + i] = (T22 * (-d12) - (-T12) * d22) * inv_x;
        row[4 * NDOF + i] = (s * d02 - T02 * ds) * inv_y;
        row[5 * NDOF + i] = (T22 * (-d01) - (-T01) * d22) * inv_z;
      }
    }
  }
  stage_store_rows<double, W>(smem_raw, row, J, row0, n, W, 0, vec_out);
}

cudaError_t launch_jac(const RobotDev<double>& rb, const double* q, int64_t n, int layout,
                       double* J, bool rpy, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  const bool vi = aligned16(q), vo = aligned16(J);
  ACRO_DISPATCH_NDOF(rb.ndof, {
    if (rpy) {
      const size_t smem = row_stage_bytes<double, 6 * NDOF>();
      auto k = jac_kernel<NDOF, true>;
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k<<<(unsigned)tiles, kTile, smem, st>>>(rb, q, n, layout, J, vi, vo);
    } else {
      jac_kernel<NDOF, false><<<(unsigned)tiles, kTile, row_stage_bytes<double, 3 * NDOF>(), st>>>(
          rb, q, n, layout, J, vi, vo);
    }
  });
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// analytical IK building blocks
// ---------------------------------------------------------------------------
struct Arm2Sol {
  double q[4][3];  // [pos_a, pos_b, neg_a, neg_b] x (q1, q2, q3)
  bool pos, neg;
  // by-products reused by kuka_ik: sin/cos(q1_pos), cos/sin(q3) per q1 branch, and the
  // (y, x) arguments of the q2 = atan2_f64(y, x) of every solution
  double s1, c1, c3[2], s3[2], y2[4], x2[4];
};

// reach test + snap of arm_2.py:37-45; returns false when unreachable
__device__ __forceinline__ bool c3_branch(double& c3, double& s3) {
  const double tol = 1e-6;
  if (c3 > (1 + tol) || c3 < -(1 + tol)) return false;
  if (c3 > (1 - tol) || c3 < -(1 - tol)) c3 = c3 > 0 ? 1.0 : -1.0;  // np.sign
  s3 = sqrt_f64(1 - c3 * c3);
  return true;
}

// atan2(-y, -x) from r = atan2(y, x): the reflected point lies half a turn away, on the side
// IEEE atan2 picks from the sign of y (signed zeros included).  Agrees with a direct
// evaluation to an ulp of pi; the closed forms below use it wherever the reference evaluates
// both atan2(y, x) and atan2(-y, -x).
__device__ __forceinline__ double atan2_reflected(double r, double y) {
  return signbit(y) ? r + CUDART_PI : r - CUDART_PI;
}

// inverse_kinematics/arm_2.py:5-104.  Transcendentals are evaluated once and reused through
// exact identities (the reference recomputes them): sin/cos(q1_neg) = -sin/cos(q1_pos),
// atan2_f64(-s3, c3) = -atan2_f64(s3, c3), cos(q3b) = c3 and sin(q3a) = s3 (arm_2.py:65, :76-77).
__device__ __forceinline__ void arm2_ik(double x, double y, double z, double a1, double a2,
                                        double a3, Arm2Sol& S) {
  const double nan = CUDART_NAN;
  const double q1_pos = atan2_f64(y, x);
  const double q1_neg = atan2_reflected(q1_pos, y);
  const double den = 2 * a2 * a3;
  const double num = a1 * a1 - a2 * a2 - a3 * a3 + x * x + y * y + z * z;
  const double L = sqrt_f64(x * x + y * y);
#pragma unroll
  for (int k = 0; k < 4; ++k) S.q[k][0] = S.q[k][1] = S.q[k][2] = nan;

  double s1, c1, c3, s3 = 0;
  sincos_f64(q1_pos, &s1, &c1);
  S.s1 = s1;
  S.c1 = c1;
  c3 = (num - 2 * a1 * s1 * y - 2 * a1 * x * c1) / den;
  S.pos = c3_branch(c3, s3);
  S.c3[0] = c3;
  S.s3[0] = s3;
  if (S.pos) {
    const double q3a = atan2_f64(s3, c3);
    const double cc = c3, ss = s3;
    S.q[0][0] = q1_pos;
    S.y2[0] = -a3 * ss * (L - a1) + z * (a2 + a3 * cc);
    S.x2[0] = a3 * ss * z + (L - a1) * (a2 + a3 * cc);
    S.q[0][2] = q3a;
    S.q[1][0] = q1_pos;
    S.y2[1] = a3 * ss * (L - a1) + z * (a2 + a3 * cc);
    S.x2[1] = -a3 * ss * z + (L - a1) * (a2 + a3 * cc);
    S.q[1][2] = -q3a;
    S.q[0][1] = atan2_f64(S.y2[0], S.x2[0]);
    S.q[1][1] = atan2_f64(S.y2[1], S.x2[1]);
  }
  c3 = (num + 2 * a1 * s1 * y + 2 * a1 * x * c1) / den;
  S.neg = c3_branch(c3, s3);
  S.c3[1] = c3;
  S.s3[1] = s3;
  if (S.neg) {
    const double q3a = atan2_f64(s3, c3);
    const double ss = s3, cc = c3;
    S.q[2][0] = q1_neg;
    S.y2[2] = a3 * ss * (L + a1) + z * (a2 + a3 * cc);
    S.x2[2] = a3 * ss * z - (L + a1) * (a2 + a3 * cc);
    S.q[2][2] = q3a;
    S.q[3][0] = q1_neg;
    S.y2[3] = -a3 * ss * (L + a1) + z * (a2 + a3 * cc);
    S.x2[3] = -a3 * ss * z - (L + a1) * (a2 + a3 * cc);
    S.q[3][2] = -q3a;
    S.q[2][1] = atan2_f64(S.y2[2], S.x2[2]);
    S.q[3][1] = atan2_f64(S.y2[3], S.x2[3]);
  }
}

// inverse_kinematics/spherical_wrist.py:5-30; Rb, R row-major 3x3.  The mirrored solution
// follows from the first one: atan2_f64(-a1, -a0) and atan2_f64(-sz, nz) are reflections,
// atan2_f64(-A, a2) = -atan2_f64(A, a2).
__device__ __forceinline__ void wrist_ik(const double* Rb, const double* R, double* sol1,
                                         double* sol2) {
  // Rrel = Rb^T R ; a = Rrel[:,2], s_z = Rrel[2,1], n_z = Rrel[2,0]
  const double a0 = Rb[0] * R[2] + Rb[3] * R[5] + Rb[6] * R[8];
  const double a1 = Rb[1] * R[2] + Rb[4] * R[5] + Rb[7] * R[8];
  const double a2 = Rb[2] * R[2] + Rb[5] * R[5] + Rb[8] * R[8];
  const double sz = Rb[2] * R[1] + Rb[5] * R[4] + Rb[8] * R[7];
  const double nz = Rb[2] * R[0] + Rb[5] * R[3] + Rb[8] * R[6];
  const double A = sqrt_f64(a0 * a0 + a1 * a1);
  sol1[0] = atan2_f64(a1, a0);
  sol1[1] = atan2_f64(A, a2);
  sol1[2] = atan2_f64(sz, -nz);
  sol2[0] = atan2_reflected(sol1[0], a1);
  sol2[1] = -sol1[1];
  sol2[2] = atan2_reflected(sol1[2], sz);
}

// Pose in the arm's frame: Tw = tf_inverse(tf_base) @ T @ tf_inverse(tf_tool_tip)
// (robot_examples.py:190-194, :14-23).  Tin: 16 doubles row-major.
__device__ __forceinline__ void kuka_pose_in_arm_frame(const RobotDev<double>& rb, const double* Tin,
                                                       Tf<double>& Tw) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    Tw.r[3 * r] = Tin[4 * r];
    Tw.r[3 * r + 1] = Tin[4 * r + 1];
    Tw.r[3 * r + 2] = Tin[4 * r + 2];
    Tw.p[r] = Tin[4 * r + 3];
  }
  if (!rb.base_identity) {
    Tf<double> X;
    const double* b = rb.base;
    const double dx = Tw.p[0] - b[3], dy = Tw.p[1] - b[7], dz = Tw.p[2] - b[11];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
        X.r[3 * i + j] = b[i] * Tw.r[j] + b[4 + i] * Tw.r[3 + j] + b[8 + i] * Tw.r[6 + j];
      X.p[i] = b[i] * dx + b[4 + i] * dy + b[8 + i] * dz;
    }
    Tw = X;
  }
  if (rb.has_tool_tip) {
    Tf<double> X;
    const double* t = rb.tip;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j)  // R' = R * Rt^T
        X.r[3 * i + j] = Tw.r[3 * i] * t[4 * j] + Tw.r[3 * i + 1] * t[4 * j + 1] +
                         Tw.r[3 * i + 2] * t[4 * j + 2];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
      X.p[i] = Tw.p[i] - (X.r[3 * i] * t[3] + X.r[3 * i + 1] * t[7] + X.r[3 * i + 2] * t[11]);
    Tw = X;
  }
}

// One q1 branch of Kuka.ik (robot_examples.py:188-218 = arm_2.py:5-104 for the arm +
// spherical_wrist.py:5-30 for the wrist): br = 0 is q1 = atan2_f64(y, x), br = 1 the reflected
// one.  The two elbow solutions of the branch (arm slots k = 2 br and 2 br + 1) with their two
// wrist solutions each go to emit(kk, ...) (kk = 0, 1: 12 doubles, NaN when the branch is out
// of reach); returns the 4 validity bits of the branch.  The arithmetic is that of the
// whole-pose form, operation for operation (identities for the repeated transcendentals: see
// arm2_ik / wrist_ik), so the two-threads-per-pose kernel returns the same bits.
template <typename Emit>
__device__ __forceinline__ unsigned kuka_ik_branch(const RobotDev<double>& rb, const Tf<double>& Tw,
                                                   const int br, Emit emit) {
  const double a1 = rb.a[0], a2 = rb.a[1], a3 = rb.d[3], d6 = rb.d[5];
  // wrist centre (:196-199)
  const double x = Tw.p[0] - Tw.r[2] * d6;
  const double y = Tw.p[1] - Tw.r[5] * d6;
  const double z = Tw.p[2] - Tw.r[8] * d6;
  const double nan = CUDART_NAN;
  const double half_pi_cos = 6.123233995736766e-17;  // np.cos(np.pi / 2)
  const double q1_pos = atan2_f64(y, x);
  const double q1 = br ? atan2_reflected(q1_pos, y) : q1_pos;
  const double den = 2 * a2 * a3;
  const double num = a1 * a1 - a2 * a2 - a3 * a3 + x * x + y * y + z * z;
  const double L = sqrt_f64(x * x + y * y);
  double s1, c1;
  sincos_f64(q1_pos, &s1, &c1);
  double c3 = br ? (num + 2 * a1 * s1 * y + 2 * a1 * x * c1) / den
                 : (num - 2 * a1 * s1 * y - 2 * a1 * x * c1) / den;
  double s3 = 0;
  if (!c3_branch(c3, s3)) {
    const double2 nn = make_double2(nan, nan);
    emit(0, nn, nn, nn, nn, nn, nn);
    emit(1, nn, nn, nn, nn, nn, nn);
    return 0u;
  }
  const double q3a = atan2_f64(s3, c3);
  const double e = a2 + a3 * c3;
  // (y2, x2) of q2 = atan2_f64(y2, x2) for the two elbow solutions (arm_2.py:47-104)
  double y2[2], x2[2];
  if (!br) {
    y2[0] = -a3 * s3 * (L - a1) + z * e;
    x2[0] = a3 * s3 * z + (L - a1) * e;
    y2[1] = a3 * s3 * (L - a1) + z * e;
    x2[1] = -a3 * s3 * z + (L - a1) * e;
  } else {
    y2[0] = a3 * s3 * (L + a1) + z * e;
    x2[0] = a3 * s3 * z - (L + a1) * e;
    y2[1] = -a3 * s3 * (L + a1) + z * e;
    x2[1] = -a3 * s3 * z - (L + a1) * e;
  }
  const double sg = br ? -1.0 : 1.0;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    const double q2 = atan2_f64(y2[kk], x2[kk]);
    const double q3 = (kk ? -q3a : q3a) + CUDART_PI / 2;  // :204
    // orientation of Arm2(a1, a2, a3=d4).fk(q_arm)   (:207, robot_examples.py:114-131), from
    // the sines and cosines already at hand
    Tf<double> F;
#pragma unroll
    for (int i = 0; i < 9; ++i) F.r[i] = (i % 4 == 0) ? 1.0 : 0.0;
    F.p[0] = F.p[1] = F.p[2] = 0.0;
    tf_mul_dh<double>(F, sg * c1, sg * s1, half_pi_cos, 1.0, a1, 0.0);
    const double h = sqrt_f64(x2[kk] * x2[kk] + y2[kk] * y2[kk]);
    const double inv = h > 0.0 ? 1.0 / h : 0.0;
    tf_mul_dh<double>(F, h > 0.0 ? x2[kk] * inv : 1.0, y2[kk] * inv, 1.0, 0.0, a2, 0.0);
    tf_mul_dh<double>(F, kk ? s3 : -s3, c3, half_pi_cos, 1.0, a3, 0.0);
    double w1[3], w2[3];
    wrist_ik(F.r, Tw.r, w1, w2);
    emit(kk, make_double2(q1, q2), make_double2(q3, w1[0]), make_double2(w1[1], w1[2]),
         make_double2(q1, q2), make_double2(q3, w2[0]), make_double2(w2[1], w2[2]));
  }
  return 0xFu;
}

// Two threads per pose (one per q1 branch): half the dependent FP64 chain and half the
// registers per thread, 24 instead of 16 resident warps per SM - the kernel is bound by the
// latency of its atan2 / sqrt chains, not by bandwidth or FP64 throughput.  Thread 2p + br
// writes units [12 br, 12 br + 12) of row p of the staging tile (odd unit stride: conflict
// free), the CTA then writes the tile as whole 128-byte lines.
constexpr int kIkPoses = kTile / 2;  // poses per CTA

template <bool WITH_CC>
__global__ void __launch_bounds__(kTile, WITH_CC ? 1 : ACRO_IK_CTAS)
    ik_kuka_kernel(const __grid_constant__ RobotDev<double> rb, const SceneView<double> sc,
                   const double* __restrict__ T, int64_t n, double* __restrict__ q_out,
                   uint8_t* __restrict__ valid, uint8_t* __restrict__ free_mask, bool vec_in,
                   bool vec_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using RS = RowStage<double, 48>;
  constexpr size_t kTileBytes = (size_t)kIkPoses * RS::STRIDE * 16;
  double* s_scene = reinterpret_cast<double*>(smem_raw + kTileBytes);
  const int64_t row0 = (int64_t)blockIdx.x * kIkPoses;
  const int t = threadIdx.x;
  const int pl = t >> 1, br = t & 1;  // local pose, q1 branch
  const int64_t idx = row0 + pl;
  if (WITH_CC) {
    for (int k = t; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];
    __syncthreads();
  }
  alignas(16) double Tin[16];
  if (idx < n) {
    load_row<double, 16>(Tin, T + idx * 16, vec_in);
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) Tin[k] = (k % 5 == 0) ? 1.0 : 0.0;
  }
  Tf<double> Tw;
  kuka_pose_in_arm_frame(rb, Tin, Tw);
  double2* s_tile2 = reinterpret_cast<double2*>(smem_raw);
  double* s_plain = reinterpret_cast<double*>(smem_raw);
  const unsigned bits4 = kuka_ik_branch(rb, Tw, br, [&](int kk, double2 u0, double2 u1, double2 u2,
                                                        double2 u3, double2 u4, double2 u5) {
    const double2 u[6] = {u0, u1, u2, u3, u4, u5};
    if (vec_out) {
#pragma unroll
      for (int e = 0; e < 6; ++e) s_tile2[pl * RS::STRIDE + 12 * br + 6 * kk + e] = u[e];
    } else {
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        s_plain[pl * 49 + 24 * br + 12 * kk + 2 * e] = u[e].x;
        s_plain[pl * 49 + 24 * br + 12 * kk + 2 * e + 1] = u[e].y;
      }
    }
  });
  // the pose's 8 validity bits: branch 0 -> bits 0..3, branch 1 -> bits 4..7
  const unsigned other = __shfl_xor_sync(0xffffffffu, bits4, 1);
  const unsigned vbits = br ? (other | (bits4 << 4)) : (bits4 | (other << 4));
  if (idx < n && br == 0) valid[idx] = (uint8_t)vbits;
  if (WITH_CC) {
    unsigned freebits = 0;
    for (int k4 = 0; k4 < 4; ++k4) {  // this thread's four slots (path/path_pt_base.py:107-109)
      const int sl = 4 * br + k4;
      const bool act = idx < n && ((vbits >> sl) & 1);
      double qv[6];
#pragma unroll
      for (int e = 0; e < 6; ++e)
        qv[e] = !act ? 0.0
                     : (vec_out ? reinterpret_cast<const double*>(s_tile2 + pl * RS::STRIDE)[6 * sl + e]
                                : s_plain[pl * 49 + 6 * sl + e]);
      Verdict<double, false> v;
      config_collision<double, 6, false>(rb, qv, s_scene, sc.n, 0.0, act, v);
      if (act && !v.hit) freebits |= 1u << sl;
    }
    freebits |= __shfl_xor_sync(0xffffffffu, freebits, 1);
    if (idx < n && br == 0) free_mask[idx] = (uint8_t)freebits;
  }
  // flat copy-out of the tile: whole 128-byte lines
  __syncthreads();
  const int64_t left = n - row0;
  const int rows = left < kIkPoses ? (int)left : kIkPoses;
  if (vec_out) {
    const int4* s_tile = reinterpret_cast<const int4*>(smem_raw);
    const int total = rows * RS::UNITS;
    for (int e = t; e < total; e += kTile) {
      const int r = e / RS::UNITS, c = e - r * RS::UNITS;
      __stcs(reinterpret_cast<int4*>(q_out + (row0 + r) * 48) + c, s_tile[r * RS::STRIDE + c]);
    }
  } else {
    const int total = rows * 48;
    for (int e = t; e < total; e += kTile) {
      const int r = e / 48, c = e - r * 48;
      __stcs(q_out + (row0 + r) * 48 + c, s_plain[r * 49 + c]);
    }
  }
}

constexpr size_t kIkTileBytes = (size_t)kIkPoses * (RowStage<double, 48>::STRIDE * 16 > 49 * 8
                                                        ? RowStage<double, 48>::STRIDE * 16
                                                        : 49 * 8);

cudaError_t launch_ik_kuka(const RobotDev<double>& rb, const double* T, int64_t n, double* q_out,
                           uint8_t* valid, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kIkPoses - 1) / kIkPoses;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  const size_t smem = kIkTileBytes;
  auto k = ik_kuka_kernel<false>;
  k<<<(unsigned)tiles, kTile, smem, st>>>(rb, SceneView<double>{nullptr, 0}, T, n, q_out, valid,
                                          nullptr, aligned16(T), aligned16(q_out));
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ik_kuka_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* T,
                              int64_t n, double* q_out, uint8_t* valid, uint8_t* free_mask,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kIkPoses - 1) / kIkPoses;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  const size_t smem = kIkTileBytes + sizeof(double) * (sc.n * kSceneStride);
  auto k = ik_kuka_kernel<true>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<(unsigned)tiles, kTile, smem, st>>>(rb, sc, T, n, q_out, valid, free_mask, aligned16(T),
                                          aligned16(q_out));
  count_launch();
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
    free_mask_kernel(const uint8_t* __restrict__ valid, const uint8_t* __restrict__ hit, int64_t n,
                     uint8_t* __restrict__ free_mask) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) free_mask[i] = valid[i] & (uint8_t)~hit[i];
}

cudaError_t launch_free_mask(const uint8_t* valid, const uint8_t* hit, int64_t n, uint8_t* free_mask,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  free_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(valid, hit, n, free_mask);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// stand-alone arm / wrist IK
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile)
    ik_arm2_kernel(double a1, double a2, double a3, const double* __restrict__ p, int64_t n,
                   double* __restrict__ q_out, uint8_t* __restrict__ valid) {
  __shared__ double s_rows[kTile * 13];
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t idx = row0 + threadIdx.x;
  fetch_rows<3>(s_rows, p, row0, n);
  __syncthreads();
  const double x = idx < n ? s_rows[threadIdx.x * 3] : 1.0;
  const double y = idx < n ? s_rows[threadIdx.x * 3 + 1] : 0.0;
  const double z = idx < n ? s_rows[threadIdx.x * 3 + 2] : 0.0;
  __syncthreads();
  Arm2Sol S;
  arm2_ik(x, y, z, a1, a2, a3, S);
  double* row = s_rows + threadIdx.x * 13;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int e = 0; e < 3; ++e) row[3 * k + e] = S.q[k][e];
  __syncthreads();
  flush_rows<12>(q_out, s_rows, row0, n);
  if (idx < n) valid[idx] = (uint8_t)((S.pos ? 3u : 0u) | (S.neg ? 12u : 0u));
}

// inverse_kinematics/anthropomorphic_arm.py:5-55
__global__ void __launch_bounds__(kTile)
    ik_anthro_kernel(double l2, double l3, const double* __restrict__ p, int64_t n,
                     double* __restrict__ q_out, uint8_t* __restrict__ valid) {
  __shared__ double s_rows[kTile * 13];
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t idx = row0 + threadIdx.x;
  fetch_rows<3>(s_rows, p, row0, n);
  __syncthreads();
  const double px = idx < n ? s_rows[threadIdx.x * 3] : 1.0;
  const double py = idx < n ? s_rows[threadIdx.x * 3 + 1] : 0.0;
  const double pz = idx < n ? s_rows[threadIdx.x * 3 + 2] : 0.0;
  __syncthreads();
  const double nan = CUDART_NAN;
  double sol[4][3];
  const double Lps = px * px + py * py + pz * pz;
  double c3 = (Lps - l2 * l2 - l3 * l3) / (2 * l2 * l3), s3 = 0;
  const bool ok = c3_branch(c3, s3);
  if (ok) {
    const double t3_1 = atan2_f64(s3, c3), t3_2 = atan2_f64(-s3, c3);
    const double Lxy = sqrt_f64(px * px + py * py);
    const double A = l2 + l3 * c3, B = l3 * s3;
    const double t2_1 = atan2_f64(pz * A - Lxy * B, Lxy * A + pz * B);
    const double t2_2 = atan2_f64(pz * A + Lxy * B, -Lxy * A + pz * B);
    const double t2_3 = atan2_f64(pz * A + Lxy * B, Lxy * A - pz * B);
    const double t2_4 = atan2_f64(pz * A - Lxy * B, -Lxy * A - pz * B);
    const double t1_1 = atan2_f64(py, px), t1_2 = atan2_f64(-py, -px);
    sol[0][0] = t1_1; sol[0][1] = t2_1; sol[0][2] = t3_1;
    sol[1][0] = t1_1; sol[1][1] = t2_3; sol[1][2] = t3_2;
    sol[2][0] = t1_2; sol[2][1] = t2_2; sol[2][2] = t3_1;
    sol[3][0] = t1_2; sol[3][1] = t2_4; sol[3][2] = t3_2;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) sol[k][0] = sol[k][1] = sol[k][2] = nan;
  }
  double* row = s_rows + threadIdx.x * 13;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int e = 0; e < 3; ++e) row[3 * k + e] = sol[k][e];
  __syncthreads();
  flush_rows<12>(q_out, s_rows, row0, n);
  if (idx < n) valid[idx] = ok ? 0xF : 0;
}

struct Mat3 {
  double m[9];
};

__global__ void __launch_bounds__(kTile)
    ik_wrist_kernel(const double* __restrict__ R, Mat3 Rb, int64_t n, double* __restrict__ q_out) {
  __shared__ double s_rows[kTile * 9];
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t idx = row0 + threadIdx.x;
  fetch_rows<9>(s_rows, R, row0, n);
  __syncthreads();
  double Rm[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Rm[k] = idx < n ? s_rows[threadIdx.x * 9 + k] : (k % 4 == 0);
  __syncthreads();
  double w1[3], w2[3];
  wrist_ik(Rb.m, Rm, w1, w2);
  double* row = s_rows + threadIdx.x * 7;
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    row[e] = w1[e];
    row[3 + e] = w2[e];
  }
  __syncthreads();
  flush_rows<6>(q_out, s_rows, row0, n);
}

cudaError_t launch_ik_arm2(double a1, double a2, double a3, const double* p, int64_t n,
                           double* q_out, uint8_t* valid, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  ik_arm2_kernel<<<(unsigned)tiles, kTile, 0, st>>>(a1, a2, a3, p, n, q_out, valid);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ik_anthro(double l2, double l3, const double* p, int64_t n, double* q_out,
                             uint8_t* valid, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  ik_anthro_kernel<<<(unsigned)tiles, kTile, 0, st>>>(l2, l3, p, n, q_out, valid);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ik_wrist(const double* R, const double* Rb9, int64_t n, double* q_out,
                            cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  Mat3 Rb;
  for (int k = 0; k < 9; ++k) Rb.m[k] = Rb9[k];
  ik_wrist_kernel<<<(unsigned)tiles, kTile, 0, st>>>(R, Rb, n, q_out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
