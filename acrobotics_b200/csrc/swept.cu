// swept.cu - K7: swept (continuous) collision check of joint-space segments.
//
// Replaces reference Robot.is_path_in_collision (src/acrobotics/robot.py:285-317),
// ShapeSoup.is_path_in_collision (src/acrobotics/geometry.py:68-81) and the python-fcl call
// fcl.continuousCollide(..., CCDM_LINEAR) of Shape.is_path_in_collision
// (src/acrobotics/shapes.py:60-75) for a batch of (q_start, q_goal) pairs.
//
// What that FCL call computes (default request: naive solver, num_max_iterations = 10,
// toc_err = 1e-4): every tool / link box is its own moving object
// between its world pose at q_start and at q_goal under FCL's InterpMotion - the centre on the
// CHORD between the two positions, the orientation about one fixed axis (angle-axis of
// R_goal R_start^T, angle in [0, pi]) at constant rate - sampled at t = i / (n - 1),
// i = 0..n-1 (n = 10), with the discrete Box-Box test (boxBox2, acro_collide.cuh) at every
// sample.  Base boxes are tested once at tf_base; no self collision (robot.py:288-289).
// The verdict is the stateless one: python-fcl accumulates is_collide in the result object
// the reference keeps per Shape, so the reference itself answers True for every later query of
// a shape that has collided once.
//
// One thread per pair: both FK chains advance in lock step, scene boxes in shared memory, the
// verdicts of a warp packed by ballot.  Per moving box a chord cull first drops every scene
// box that one of its own face axes separates from the sphere swept along the chord (that
// axis then separates every sample in FCL's test too).
#include "acro_collide.cuh"
#include "acro_launch.h"

namespace acro {

// Eigen AngleAxis(Rb Ra^T) through Eigen's Matrix3 -> Quaternion (interp_motion-inl.h,
// computeVelocity): angle in [0, pi], axis (1,0,0) for the identity.
__device__ __forceinline__ void relative_angle_axis(const double* Ra, const double* Rb,
                                                    double& angle, double* ax) {
  double m[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      m[3 * r + c] = fma(Rb[3 * r + 2], Ra[3 * c + 2],
                         fma(Rb[3 * r + 1], Ra[3 * c + 1], Rb[3 * r] * Ra[3 * c]));
  const double t = m[0] + m[4] + m[8];
  double w, v[3];
  if (t > 0.0) {
    double s = sqrt(t + 1.0);
    w = 0.5 * s;
    s = 0.5 / s;
    v[0] = (m[7] - m[5]) * s;
    v[1] = (m[2] - m[6]) * s;
    v[2] = (m[3] - m[1]) * s;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    v[i] = 0.5 * s;
    s = 0.5 / s;
    w = (m[3 * k + j] - m[3 * j + k]) * s;
    v[j] = (m[3 * j + i] + m[3 * i + j]) * s;
    v[k] = (m[3 * k + i] + m[3 * i + k]) * s;
  }
  double n = sqrt(fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0])));
  if (n != 0.0) {
    angle = 2.0 * atan2(n, fabs(w));
    if (w < 0.0) n = -n;
    const double inv = 1.0 / n;
    ax[0] = v[0] * inv;
    ax[1] = v[1] * inv;
    ax[2] = v[2] * inv;
  } else {
    angle = 0.0;
    ax[0] = 1.0;
    ax[1] = 0.0;
    ax[2] = 0.0;
  }
}

// R = Rot(axis, theta) * Ra
__device__ __forceinline__ void rotate_about(const double* ax, double theta, const double* Ra,
                                             double* R) {
  double s, c;
  sincos(theta, &s, &c);
  const double k = 1.0 - c;
  double D[9];
  D[0] = fma(k * ax[0], ax[0], c);
  D[4] = fma(k * ax[1], ax[1], c);
  D[8] = fma(k * ax[2], ax[2], c);
  const double xy = k * ax[0] * ax[1], xz = k * ax[0] * ax[2], yz = k * ax[1] * ax[2];
  D[1] = xy - s * ax[2];
  D[3] = xy + s * ax[2];
  D[2] = xz + s * ax[1];
  D[6] = xz - s * ax[1];
  D[5] = yz - s * ax[0];
  D[7] = yz + s * ax[0];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int col = 0; col < 3; ++col)
      R[3 * r + col] = fma(D[3 * r + 2], Ra[6 + col], fma(D[3 * r + 1], Ra[3 + col], D[3 * r] * Ra[col]));
}

#ifndef ACRO_SWEPT_MIN_CTAS
#define ACRO_SWEPT_MIN_CTAS 2
#endif

template <int NDOF, bool MARGIN>
__global__ void __launch_bounds__(kTile, ACRO_SWEPT_MIN_CTAS)
    swept_cc_kernel(const __grid_constant__ RobotDev<double> rb, const SceneView<double> sc,
                    const double* __restrict__ q0, const double* __restrict__ q1, int64_t n,
                    int n_samples, bool vec, uint32_t* __restrict__ mask,
                    uint32_t* __restrict__ near_mask, double margin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_scene = reinterpret_cast<double*>(smem_raw);
  for (int k = threadIdx.x; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];
  __syncthreads();

  const int64_t idx = (int64_t)blockIdx.x * kTile + threadIdx.x;
  const bool active = idx < n;
  // joint values: read as rows into registers, then parked in shared memory (column tid) so
  // that the ROLLED joint loop below can index them - one copy of the sweep code instead of
  // NDOF + 1 (the unrolled kernel stalled on instruction fetch: ncu no_instruction 4.7 of 8.4
  // cycles per issue)
  double* s_qa = s_scene + sc.n * kSceneStride;
  double* s_qb = s_qa + NDOF * kTile;
  {
    double qa[NDOF], qb[NDOF];
    load_q_direct<double, NDOF>(qa, q0, idx, n, ACRO_Q_AOS, vec);
    load_q_direct<double, NDOF>(qb, q1, idx, n, ACRO_Q_AOS, vec);
#pragma unroll
    for (int i = 0; i < NDOF; ++i) {
      s_qa[i * kTile + threadIdx.x] = qa[i];
      s_qb[i * kTile + threadIdx.x] = qb[i];
    }
  }

  Verdict<double, MARGIN> v;
  const double pad = MARGIN ? margin : 0.0;
  const unsigned full = 0xffffffffu;
  const int n_scene = sc.n;
  Tf<double> Ta, Tb;
  tf_load<double>(Ta, rb.base);
  Tb = Ta;

  // geometry_base: fixed at tf_base, one discrete test (robot.py:303-306)
  for (int b = rb.slot_begin[0]; b < rb.slot_begin[1]; ++b) {
    const BoxDev<double>& bx = rb.boxes[b];
    Tf<double> W;
    tf_mul<double>(W, Ta, bx.off);
    if (active && !v.done(margin) && n_scene > 0)
      box_vs_scene<double, MARGIN>(bx, W.r, W.p, s_scene, n_scene, margin, v);
  }

  auto sweep_slot = [&](int slot) {
    for (int b = rb.slot_begin[slot]; b < rb.slot_begin[slot + 1]; ++b) {
      const BoxDev<double>& bx = rb.boxes[b];
      if (!active || v.done(margin) || n_scene == 0) continue;
      Tf<double> Wa, Wb;
      tf_mul<double>(Wa, Ta, bx.off);
      tf_mul<double>(Wb, Tb, bx.off);
      // chord cull: sphere of radius |half| swept along the chord vs the scene box's own faces
      const double l0 = Wb.p[0] - Wa.p[0], l1 = Wb.p[1] - Wa.p[1], l2 = Wb.p[2] - Wa.p[2];
      const double reach =
          bx.radius + pad + 0.5 * sqrt(fma(l2, l2, fma(l1, l1, l0 * l0))) * (1.0 + 1e-12) + 1e-12;
      const double mid[3] = {fma(0.5, l0, Wa.p[0]), fma(0.5, l1, Wa.p[1]), fma(0.5, l2, Wa.p[2])};
      uint64_t cand = 0;
      for (int j = 0; j < n_scene; ++j) {
        const double* b2 = s_scene + j * kSceneStride;
        const double d0 = mid[0] - b2[kOB_T + 0], d1 = mid[1] - b2[kOB_T + 1],
                     d2 = mid[2] - b2[kOB_T + 2];
        const double y0 = fma(b2[6], d2, fma(b2[3], d1, b2[0] * d0));
        const double y1 = fma(b2[7], d2, fma(b2[4], d1, b2[1] * d0));
        const double y2 = fma(b2[8], d2, fma(b2[5], d1, b2[2] * d0));
        const bool sep = (fabs(y0) > b2[kOB_H + 0] + reach) | (fabs(y1) > b2[kOB_H + 1] + reach) |
                         (fabs(y2) > b2[kOB_H + 2] + reach);
        if (!sep) cand |= (1ull << j);
      }
      if (cand == 0) continue;
      double angle, ax[3];
      relative_angle_axis(Wa.r, Wb.r, angle, ax);
      for (int s = 0; s < n_samples && !v.done(margin); ++s) {
        const double t = (double)s / (double)(n_samples - 1);
        double R[9], c[3];
        rotate_about(ax, t * angle, Wa.r, R);
        c[0] = fma(t, l0, Wa.p[0]);
        c[1] = fma(t, l1, Wa.p[1]);
        c[2] = fma(t, l2, Wa.p[2]);
        uint64_t todo = cand;
        while (todo != 0 && !v.done(margin)) {
          const int j = __ffsll((long long)todo) - 1;
          todo &= todo - 1;
          const double* b2 = s_scene + j * kSceneStride;
          if (cull_separated<double>(bx.half[0], bx.half[1], bx.half[2], bx.radius, R, c, b2, pad))
            continue;
          double g;
          if (!sat_separated<double>(bx.half[0], bx.half[1], bx.half[2], R, c, b2, pad, g)) {
            if (MARGIN) {
              v.gmin = min(v.gmin, g);
              if (g <= 0.0) v.hit = true;
            } else {
              v.hit = true;
            }
          }
        }
      }
    }
  };

  // slots 1..NDOF: DH step of both chains, then that link's boxes; slot NDOF + 1: the tool on
  // the last link frame (robot.py:296-300)
#pragma unroll 1
  for (int slot = 1; slot <= NDOF + 1; ++slot) {
    if (__all_sync(full, !active || v.done(margin))) break;
    if (slot <= NDOF) {
      chain_step<double>(Ta, rb, slot - 1, s_qa[(slot - 1) * kTile + threadIdx.x]);
      chain_step<double>(Tb, rb, slot - 1, s_qb[(slot - 1) * kTile + threadIdx.x]);
    }
    sweep_slot(slot);
  }

  const unsigned word = __ballot_sync(full, active && v.hit);
  unsigned near_word = 0;
  if (MARGIN) near_word = __ballot_sync(full, active && v.gmin <= margin && v.gmin >= -margin);
  if ((threadIdx.x & 31) == 0 && idx < n) {
    mask[idx >> 5] = word;
    if (MARGIN) near_mask[idx >> 5] = near_word;
  }
}

cudaError_t launch_swept_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* q0,
                            const double* q1, int64_t n, int n_samples, uint32_t* mask,
                            uint32_t* near_mask, double margin, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL || n_samples < 2) return cudaErrorInvalidValue;
  ACRO_DISPATCH_NDOF(rb.ndof, {
    const size_t smem = sizeof(double) * (sc.n * kSceneStride + 2 * NDOF * kTile);
    const bool vec = NDOF % 2 == 0 && aligned16(q0) && aligned16(q1);
    if (near_mask)
      swept_cc_kernel<NDOF, true><<<(unsigned)tiles, kTile, smem, st>>>(
          rb, sc, q0, q1, n, n_samples, vec, mask, near_mask, margin);
    else
      swept_cc_kernel<NDOF, false><<<(unsigned)tiles, kTile, smem, st>>>(
          rb, sc, q0, q1, n, n_samples, vec, mask, nullptr, 0.0);
  });
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
