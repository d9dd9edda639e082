// acro_collide.cuh - the per-configuration FK -> collision core shared by the
// fused FK+CC kernel (K2), the path sweep (K6) and the IK -> CC filter.
//
// Box-box predicate = FCL boxBox2 (ODE dBoxBox2) as reached from reference
// src/acrobotics/shapes.py:50-58; orchestration = robot.py:244-273 (scene) and
// robot.py:206-221 (self pairs); per-box world pose = geometry.py:56-57.
#pragma once
#include "acro_device.cuh"

namespace acro {

// packed "other box" layout shared by scene boxes (shared memory) and saved
// robot boxes (local memory): [0..8] R row-major, [9..11] centre, [12..14] half
// extents, [15] bounding radius
constexpr int kOB_R = 0, kOB_T = 9, kOB_H = 12, kOB_RAD = 15;

// Separating-axis test.  Returns true as soon as one of the 15 FCL axes
// separates the boxes by more than `thr` (thr = 0 reproduces FCL's boolean:
// face axes `s2 > 0`, edge axes `s2 > eps`).  Otherwise returns false with
// g = the largest decision quantity over all 15 axes (<= thr).
template <typename real>
__device__ __forceinline__ bool sat_separated(const real A0, const real A1, const real A2,
                                              const real* __restrict__ R1, const real* c1,
                                              const real* b2, const real thr, real& g_out) {
  const real p0 = b2[kOB_T + 0] - c1[0];
  const real p1 = b2[kOB_T + 1] - c1[1];
  const real p2 = b2[kOB_T + 2] - c1[2];
  const real B0 = b2[kOB_H + 0], B1 = b2[kOB_H + 1], B2 = b2[kOB_H + 2];
  real R2[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) R2[k] = b2[kOB_R + k];

  // R = R1^T R2 ; Q = |R| ; pp = R1^T p.  The order in which the 15 axes are
  // visited is free (the verdict is "no axis separates"); box 2's face axes go
  // first and R is built column by column for them: they are the axes the
  // bounding-sphere cull pre-tested, and for plate-like obstacles (a table) they
  // are the ones that separate.
  real R[9], Q[9];
  const real A[3] = {A0, A1, A2};
  const real B[3] = {B0, B1, B2};
  real g = real(-1e30);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      R[3 * i + j] = fma_t(R1[6 + i], R2[6 + j], fma_t(R1[3 + i], R2[3 + j], R1[i] * R2[j]));
      Q[3 * i + j] = abs_t(R[3 * i + j]);
    }
    // face axis v_j of box 2
    const real t = fma_t(R2[6 + j], p2, fma_t(R2[3 + j], p1, R2[j] * p0));
    const real rad = fma_t(Q[6 + j], A2, fma_t(Q[3 + j], A1, Q[j] * A0)) + B[j];
    g = max(g, abs_t(t) - rad);
  }
  if (g > thr) return true;

  // face axes of box 1 (u1,u2,u3)
  real pp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pp[i] = fma_t(R1[6 + i], p2, fma_t(R1[3 + i], p1, R1[i] * p0));
  g = max(g, abs_t(pp[0]) - (fma_t(Q[2], B2, fma_t(Q[1], B1, Q[0] * B0)) + A0));
  g = max(g, abs_t(pp[1]) - (fma_t(Q[5], B2, fma_t(Q[4], B1, Q[3] * B0)) + A1));
  g = max(g, abs_t(pp[2]) - (fma_t(Q[8], B2, fma_t(Q[7], B1, Q[6] * B0)) + A2));
  if (g > thr) return true;

  // nine edge x edge axes with ODE's fudge on |R|; FCL compares with eps here
  real Qf[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Qf[k] = Q[k] + real(1.0e-6);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const real t = fma_t(pp[i2], R[3 * i1 + j], -(pp[i1] * R[3 * i2 + j]));
      const real rad = fma_t(A[i1], Qf[3 * i2 + j],
                             fma_t(A[i2], Qf[3 * i1 + j],
                                   fma_t(B[j1], Qf[3 * i + j2], B[j2] * Qf[3 * i + j1])));
      g = max(g, (abs_t(t) - rad) - Eps<real>::v);
    }
    if (g > thr) return true;
  }
  g_out = g;
  return false;
}

// Conservative cull that implies an FCL face axis separates:
//  * box 2's face axes with box 1 replaced by its bounding sphere (radius rA):
//      |v_k . d| > B_k + rA   =>   FCL's v_k test has s2 > 0
//  * box 1's face axes with box 2 replaced by its bounding sphere (radius rB).
// `pad` (>= 0) widens the thresholds (margin mode).
template <typename real>
__device__ __forceinline__ bool cull_separated(const real A0, const real A1, const real A2,
                                               const real rA, const real* __restrict__ R1,
                                               const real* c1, const real* b2, const real pad) {
  const real d0 = c1[0] - b2[kOB_T + 0];
  const real d1 = c1[1] - b2[kOB_T + 1];
  const real d2 = c1[2] - b2[kOB_T + 2];
  const real ra = rA + pad;
  const real y0 = fma_t(b2[6], d2, fma_t(b2[3], d1, b2[0] * d0));
  const real y1 = fma_t(b2[7], d2, fma_t(b2[4], d1, b2[1] * d0));
  const real y2 = fma_t(b2[8], d2, fma_t(b2[5], d1, b2[2] * d0));
  bool sep = (abs_t(y0) > b2[kOB_H + 0] + ra) | (abs_t(y1) > b2[kOB_H + 1] + ra) |
             (abs_t(y2) > b2[kOB_H + 2] + ra);
  if (!sep) {
    const real rb = b2[kOB_RAD] + pad;
    const real z0 = fma_t(R1[6], d2, fma_t(R1[3], d1, R1[0] * d0));
    const real z1 = fma_t(R1[7], d2, fma_t(R1[4], d1, R1[1] * d0));
    const real z2 = fma_t(R1[8], d2, fma_t(R1[5], d1, R1[2] * d0));
    sep = (abs_t(z0) > A0 + rb) | (abs_t(z1) > A1 + rb) | (abs_t(z2) > A2 + rb);
  }
  return sep;
}

// Verdict accumulator for one configuration.
template <typename real, bool MARGIN>
struct Verdict {
  bool hit = false;
  real gmin = real(1e30);  // MARGIN: smallest pair-wise g over non-robustly-separated pairs
  __device__ __forceinline__ bool done(real margin) const {
    return MARGIN ? (gmin < -margin) : hit;
  }
  __device__ __forceinline__ void add_pair(real g) {
    hit = true;  // sat_separated returned false with thr=0  <=> collision
    if (MARGIN) gmin = min(gmin, g);
  }
};

// one robot box (world pose Rw, cw) against a packed list of boxes
template <typename real, bool MARGIN>
__device__ __forceinline__ void box_vs_scene(const BoxDev<real>& bx, const real* Rw,
                                             const real* cw, const real* s_scene, int n_scene,
                                             real margin, Verdict<real, MARGIN>& v) {
  const real A0 = bx.half[0], A1 = bx.half[1], A2 = bx.half[2];
  uint64_t cand = 0;
  const real pad = MARGIN ? margin : real(0);
  for (int j = 0; j < n_scene; ++j) {
    if (!cull_separated<real>(A0, A1, A2, bx.radius, Rw, cw, s_scene + j * kSceneStride, pad))
      cand |= (1ull << j);
  }
  while (cand != 0 && !v.done(margin)) {
    const int j = __ffsll((long long)cand) - 1;
    cand &= cand - 1;
    real g;
    if (!sat_separated<real>(A0, A1, A2, Rw, cw, s_scene + j * kSceneStride, pad, g)) {
      if (MARGIN) {
        v.gmin = min(v.gmin, g);
        if (g <= real(0)) v.hit = true;
      } else {
        v.hit = true;
      }
    }
  }
}

// The whole FK -> collision evaluation of one configuration.
//   qv       joint values of this thread's configuration
//   s_scene  scene boxes in shared memory (n_scene * kSceneStride reals)
// Processing order: base boxes (carrier tf_base), then for every link i the DH
// step followed by that link's boxes, then the tool boxes on the last link
// frame.  A box that is the earlier member of a self pair is parked in local
// memory; the pair is evaluated when its later box is reached.
template <typename real, int NDOF, bool MARGIN>
__device__ __forceinline__ void config_collision(const RobotDev<real>& rb, const real* qv,
                                                 const real* s_scene, int n_scene, real margin,
                                                 bool active, Verdict<real, MARGIN>& v) {
  real saved[kMaxSaved][16];
  Tf<real> T;
  tf_load<real>(T, rb.base);
  const unsigned full = 0xffffffffu;
  const bool do_self = rb.check_self != 0;

  auto process_slot = [&](int slot, const Tf<real>& F) {
    for (int b = rb.slot_begin[slot]; b < rb.slot_begin[slot + 1]; ++b) {
      const BoxDev<real>& bx = rb.boxes[b];
      real Rw[9], cw[3];
      if (bx.rot_identity) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Rw[k] = F.r[k];
#pragma unroll
        for (int r = 0; r < 3; ++r)
          cw[r] = fma_t(F.r[3 * r + 2], bx.off[11],
                        fma_t(F.r[3 * r + 1], bx.off[7], fma_t(F.r[3 * r], bx.off[3], F.p[r])));
      } else {
        Tf<real> W;
        tf_mul<real>(W, F, bx.off);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rw[k] = W.r[k];
#pragma unroll
        for (int r = 0; r < 3; ++r) cw[r] = W.p[r];
      }
      if (active && !v.done(margin)) {
        if (n_scene > 0) box_vs_scene<real, MARGIN>(bx, Rw, cw, s_scene, n_scene, margin, v);
        if (do_self) {
          const real pad = MARGIN ? margin : real(0);
          for (int k = rb.pair_begin[b]; k < rb.pair_begin[b + 1]; ++k) {
            if (v.done(margin)) break;
            const int slot_e = rb.boxes[rb.partner[k]].save_slot;
            const real* ob = saved[slot_e];
            if (cull_separated<real>(bx.half[0], bx.half[1], bx.half[2], bx.radius, Rw, cw, ob,
                                     pad))
              continue;
            real g;
            if (!sat_separated<real>(bx.half[0], bx.half[1], bx.half[2], Rw, cw, ob, pad, g)) {
              if (MARGIN) {
                v.gmin = min(v.gmin, g);
                if (g <= real(0)) v.hit = true;
              } else {
                v.hit = true;
              }
            }
          }
        }
      }
      if (do_self && bx.save_slot >= 0) {
        real* ob = saved[bx.save_slot];
#pragma unroll
        for (int k = 0; k < 9; ++k) ob[kOB_R + k] = Rw[k];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          ob[kOB_T + r] = cw[r];
          ob[kOB_H + r] = bx.half[r];
        }
        ob[kOB_RAD] = bx.radius;
      }
    }
  };

  process_slot(0, T);  // geometry_base at tf_base (robot.py:257-260)
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    // whole warp decided => nothing left to do for it
    if (__all_sync(full, !active || v.done(margin))) return;
    chain_step<real>(T, rb, i, qv[i]);
    process_slot(i + 1, T);
  }
  process_slot(NDOF + 1, T);  // geometry_tool at the last link frame (robot.py:251-254)
}

}  // namespace acro
