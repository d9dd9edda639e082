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

// One moving box, set up: its world poses on the two chains (Wa kept, the chord l = pb - pa)
// and the chord cull - a scene box whose own face axis separates it from the sphere of radius
// |half| swept along the chord is separated at every sample in FCL's face-axis test.
// Returns the bit set of the scene boxes that remain.
__device__ __forceinline__ uint64_t setup_moving_box(const BoxDev<double>& bx, const Tf<double>& Ta,
                                                     const Tf<double>& Tb, const double* s_scene,
                                                     int n_scene, double pad, Tf<double>& Wa,
                                                     double* l, double& angle, double* ax) {
  Tf<double> Wb;
  tf_mul<double>(Wa, Ta, bx.off);
  tf_mul<double>(Wb, Tb, bx.off);
  l[0] = Wb.p[0] - Wa.p[0];
  l[1] = Wb.p[1] - Wa.p[1];
  l[2] = Wb.p[2] - Wa.p[2];
  const double reach = bx.radius + pad +
                       0.5 * sqrt(fma(l[2], l[2], fma(l[1], l[1], l[0] * l[0]))) * (1.0 + 1e-12) +
                       1e-12;
  const double mid[3] = {fma(0.5, l[0], Wa.p[0]), fma(0.5, l[1], Wa.p[1]), fma(0.5, l[2], Wa.p[2])};
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
  if (cand != 0) relative_angle_axis(Wa.r, Wb.r, angle, ax);
  return cand;
}

// One sample of FCL's loop for one moving box: pose at t = s / (n - 1), Box-Box vs the candidates
template <bool MARGIN>
__device__ __forceinline__ void sample_moving_box(const BoxDev<double>& bx, const Tf<double>& Wa,
                                                  const double* l, double angle, const double* ax,
                                                  int s, int n_samples, uint64_t cand,
                                                  const double* s_scene, double margin,
                                                  Verdict<double, MARGIN>& v) {
  const double pad = MARGIN ? margin : 0.0;
  const double t = (double)s / (double)(n_samples - 1);
  double R[9], c[3];
  rotate_about(ax, t * angle, Wa.r, R);
  c[0] = fma(t, l[0], Wa.p[0]);
  c[1] = fma(t, l[1], Wa.p[1]);
  c[2] = fma(t, l[2], Wa.p[2]);
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

// The boxes of one slot (a link's, or the tool's) as moving objects between their poses on the
// two chains.
template <bool MARGIN>
__device__ __forceinline__ void sweep_slot_boxes(const RobotDev<double>& rb, int slot,
                                                 const Tf<double>& Ta, const Tf<double>& Tb,
                                                 const double* s_scene, int n_scene,
                                                 int n_samples, double margin,
                                                 Verdict<double, MARGIN>& v) {
  const double pad = MARGIN ? margin : 0.0;
  for (int b = rb.slot_begin[slot]; b < rb.slot_begin[slot + 1]; ++b) {
    const BoxDev<double>& bx = rb.boxes[b];
    if (v.done(margin) || n_scene == 0) continue;
    Tf<double> Wa;
    double l[3], angle, ax[3];
    const uint64_t cand = setup_moving_box(bx, Ta, Tb, s_scene, n_scene, pad, Wa, l, angle, ax);
    if (cand == 0) continue;
    for (int s = 0; s < n_samples && !v.done(margin); ++s)
      sample_moving_box<MARGIN>(bx, Wa, l, angle, ax, s, n_samples, cand, s_scene, margin, v);
  }
}

// geometry_base: fixed at tf_base, one discrete test (robot.py:303-306)
template <bool MARGIN>
__device__ __forceinline__ void base_slot_boxes(const RobotDev<double>& rb, const Tf<double>& T0,
                                                const double* s_scene, int n_scene,
                                                double margin, Verdict<double, MARGIN>& v) {
  for (int b = rb.slot_begin[0]; b < rb.slot_begin[1]; ++b) {
    const BoxDev<double>& bx = rb.boxes[b];
    Tf<double> W;
    tf_mul<double>(W, T0, bx.off);
    if (!v.done(margin) && n_scene > 0)
      box_vs_scene<double, MARGIN>(bx, W.r, W.p, s_scene, n_scene, margin, v);
  }
}

// 4 CTAs of 128 threads per SM (128 registers, some spills to the stack): measured best on the
// rolled kernel - 2 / 3 / 4 / 5 / 6 / 8 CTAs: 7.03 / 6.13 / 5.69 / 5.67 / 5.94 / 7.36 ms per 4 M pairs
// (profiles/r2_swept_ab.jsonl); the FP64 chains need the extra warps to hide their latency
#ifndef ACRO_SWEPT_MIN_CTAS
#define ACRO_SWEPT_MIN_CTAS 4
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
  const unsigned full = 0xffffffffu;
  const int n_scene = sc.n;
  Tf<double> Ta, Tb;
  tf_load<double>(Ta, rb.base);
  Tb = Ta;

  if (active) base_slot_boxes<MARGIN>(rb, Ta, s_scene, n_scene, margin, v);

  // slots 1..NDOF: DH step of both chains, then that link's boxes; slot NDOF + 1: the tool on
  // the last link frame (robot.py:296-300)
#pragma unroll 1
  for (int slot = 1; slot <= NDOF + 1; ++slot) {
    if (__all_sync(full, !active || v.done(margin))) break;
    if (slot <= NDOF) {
      chain_step<double>(Ta, rb, slot - 1, s_qa[(slot - 1) * kTile + threadIdx.x]);
      chain_step<double>(Tb, rb, slot - 1, s_qb[(slot - 1) * kTile + threadIdx.x]);
    }
    if (active) sweep_slot_boxes<MARGIN>(rb, slot, Ta, Tb, s_scene, n_scene, n_samples, margin, v);
  }

  const unsigned word = __ballot_sync(full, active && v.hit);
  unsigned near_word = 0;
  if (MARGIN) near_word = __ballot_sync(full, active && v.gmin <= margin && v.gmin >= -margin);
  if ((threadIdx.x & 31) == 0 && idx < n) {
    mask[idx >> 5] = word;
    if (MARGIN) near_mask[idx >> 5] = near_word;
  }
}

// Persistent variant for large batches: every LANE owns one pair at a time and advances it by
// one step per round - either the set-up of its next box (DH steps when a slot closes, world
// poses, chord cull, angle-axis) or ONE sample of the box in progress; a lane whose pair is
// decided (first colliding sample, or all slots clear) sets its verdict bit and takes the next
// pair from a grid-wide counter, so pairs that end at their start pose, boxes without
// candidates and boxes that run all ten samples keep the warps equally full (the
// one-pair-per-thread kernel above ran at 12.5 of 32 lanes).  mask /
// near_mask are zeroed by the launcher; bits are set with atomicOr.
template <int NDOF, bool MARGIN>
__global__ void __launch_bounds__(kTile, ACRO_SWEPT_MIN_CTAS)
    swept_cc_steal_kernel(const __grid_constant__ RobotDev<double> rb, const SceneView<double> sc,
                          const double* __restrict__ q0, const double* __restrict__ q1, int64_t n,
                          int n_samples, bool vec, uint32_t* __restrict__ mask,
                          uint32_t* __restrict__ near_mask, double margin,
                          unsigned long long* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_scene = reinterpret_cast<double*>(smem_raw);
  for (int k = threadIdx.x; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];
  __syncthreads();
  double* s_qa = s_scene + sc.n * kSceneStride;
  double* s_qb = s_qa + NDOF * kTile;

  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int n_scene = sc.n;
  int64_t idx = -1;  // -1: wants a pair, -2: the counter is exhausted
  int slot = 0, b = 0, cur = 0, left = 0;  // slot in progress, next box, box being sampled, samples left
  uint64_t cand = 0;
  Tf<double> Wa;
  double l[3] = {0.0, 0.0, 0.0}, angle = 0.0, ax[3] = {1.0, 0.0, 0.0};
  Verdict<double, MARGIN> v;
  Tf<double> Ta, Tb;
  tf_load<double>(Ta, rb.base);
  Tb = Ta;

  while (true) {
    const unsigned want = __ballot_sync(full, idx == -1);
    if (want != 0) {
      unsigned long long first = 0;
      const int leader = __ffs((int)want) - 1;
      if (lane == leader) first = atomicAdd(counter, (unsigned long long)__popc(want));
      first = __shfl_sync(full, first, leader);
      if (idx == -1) {
        const int64_t mine = (int64_t)first + __popc(want & ((1u << lane) - 1u));
        if (mine < n) {
          idx = mine;
          double qa[NDOF], qb[NDOF];
          load_q_direct<double, NDOF>(qa, q0, idx, n, ACRO_Q_AOS, vec);
          load_q_direct<double, NDOF>(qb, q1, idx, n, ACRO_Q_AOS, vec);
#pragma unroll
          for (int i = 0; i < NDOF; ++i) {
            s_qa[i * kTile + threadIdx.x] = qa[i];
            s_qb[i * kTile + threadIdx.x] = qb[i];
          }
          v = Verdict<double, MARGIN>();
          tf_load<double>(Ta, rb.base);
          Tb = Ta;
          base_slot_boxes<MARGIN>(rb, Ta, s_scene, n_scene, margin, v);  // robot.py:303-306
          slot = 0;
          b = rb.slot_begin[1];  // slot 0 is done: the first round opens slot 1
          left = 0;
        } else {
          idx = -2;
        }
      }
    }
    if (__all_sync(full, idx == -2)) break;
    if (idx >= 0) {
      bool finished = false;
      if (left == 0) {
        // next box: close the slot when its boxes are used up (DH step of both chains), then
        // set up one box; a box with candidates starts its n_samples sample rounds
        if (b == rb.slot_begin[slot + 1]) {
          ++slot;
          if (slot > NDOF + 1 || n_scene == 0) {
            finished = true;
          } else {
            if (slot <= NDOF) {
              chain_step<double>(Ta, rb, slot - 1, s_qa[(slot - 1) * kTile + threadIdx.x]);
              chain_step<double>(Tb, rb, slot - 1, s_qb[(slot - 1) * kTile + threadIdx.x]);
            }
            b = rb.slot_begin[slot];
          }
        }
        if (!finished && b < rb.slot_begin[slot + 1]) {
          cand = setup_moving_box(rb.boxes[b], Ta, Tb, s_scene, n_scene, MARGIN ? margin : 0.0, Wa,
                                  l, angle, ax);
          cur = b++;
          if (cand != 0) left = n_samples;
        }
      } else {
        sample_moving_box<MARGIN>(rb.boxes[cur], Wa, l, angle, ax, n_samples - left, n_samples,
                                  cand, s_scene, margin, v);
        --left;
      }
      if (v.done(margin) || finished) {
        const uint32_t bit = 1u << (idx & 31);
        if (v.hit) atomicOr(mask + (idx >> 5), bit);
        if (MARGIN && v.gmin <= margin && v.gmin >= -margin) atomicOr(near_mask + (idx >> 5), bit);
        idx = -1;
      }
    }
  }
}

// n independent (moving box, fixed box) pairs: the fcl.continuousCollide seam itself
// (shapes.py:60-75).  a0 / a1: the moving box at its start / target pose (same half extents).
__global__ void __launch_bounds__(kTile)
    swept_boxes_kernel(const AcroBox* __restrict__ a0, const AcroBox* __restrict__ a1,
                       const AcroBox* __restrict__ b, int64_t n, int n_samples,
                       uint32_t* __restrict__ mask) {
  const int64_t idx = (int64_t)blockIdx.x * kTile + threadIdx.x;
  bool hit = false;
  if (idx < n) {
    const AcroBox& A = a0[idx];
    const AcroBox& A1 = a1[idx];
    const AcroBox& B = b[idx];
    double Ra[9], Rb[9], pa[3], l[3], ob[16];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Ra[3 * r + c] = A.tf[4 * r + c];
        Rb[3 * r + c] = A1.tf[4 * r + c];
        ob[kOB_R + 3 * r + c] = B.tf[4 * r + c];
      }
      pa[r] = A.tf[4 * r + 3];
      l[r] = A1.tf[4 * r + 3] - pa[r];
      ob[kOB_T + r] = B.tf[4 * r + 3];
      ob[kOB_H + r] = B.half[r];
    }
    double angle, ax[3];
    relative_angle_axis(Ra, Rb, angle, ax);
    for (int s = 0; s < n_samples && !hit; ++s) {
      const double t = (double)s / (double)(n_samples - 1);
      double R[9], c[3], g;
      rotate_about(ax, t * angle, Ra, R);
      c[0] = fma(t, l[0], pa[0]);
      c[1] = fma(t, l[1], pa[1]);
      c[2] = fma(t, l[2], pa[2]);
      hit = !sat_separated<double>(A.half[0], A.half[1], A.half[2], R, c, ob, 0.0, g);
    }
  }
  const unsigned word = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && idx < n) mask[idx >> 5] = word;
}

cudaError_t launch_swept_boxes(const AcroBox* a0, const AcroBox* a1, const AcroBox* b, int64_t n,
                               int n_samples, uint32_t* mask, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL || n_samples < 2) return cudaErrorInvalidValue;
  swept_boxes_kernel<<<(unsigned)tiles, kTile, 0, st>>>(a0, a1, b, n, n_samples, mask);
  count_launch();
  return cudaGetLastError();
}

// counter: one zero-initialised 8-byte device word -> the persistent kernel; nullptr -> one
// thread per pair
cudaError_t launch_swept_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* q0,
                            const double* q1, int64_t n, int n_samples, uint32_t* mask,
                            uint32_t* near_mask, double margin, unsigned long long* counter,
                            cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL || n_samples < 2) return cudaErrorInvalidValue;
  if (counter) {
    const size_t words = (size_t)((n + 31) / 32);
    cudaError_t e = cudaMemsetAsync(mask, 0, sizeof(uint32_t) * words, st);
    if (e == cudaSuccess && near_mask) e = cudaMemsetAsync(near_mask, 0, sizeof(uint32_t) * words, st);
    if (e != cudaSuccess) return e;
  }
  ACRO_DISPATCH_NDOF(rb.ndof, {
    const size_t smem = sizeof(double) * (sc.n * kSceneStride + 2 * NDOF * kTile);
    const bool vec = NDOF % 2 == 0 && aligned16(q0) && aligned16(q1);
    if (counter) {
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int64_t resident = (int64_t)sms * ACRO_SWEPT_MIN_CTAS;
      const unsigned grid = (unsigned)(tiles < resident ? tiles : resident);
      if (near_mask)
        swept_cc_steal_kernel<NDOF, true><<<grid, kTile, smem, st>>>(
            rb, sc, q0, q1, n, n_samples, vec, mask, near_mask, margin, counter);
      else
        swept_cc_steal_kernel<NDOF, false><<<grid, kTile, smem, st>>>(
            rb, sc, q0, q1, n, n_samples, vec, mask, nullptr, 0.0, counter);
    } else if (near_mask) {
      swept_cc_kernel<NDOF, true><<<(unsigned)tiles, kTile, smem, st>>>(
          rb, sc, q0, q1, n, n_samples, vec, mask, near_mask, margin);
    } else {
      swept_cc_kernel<NDOF, false><<<(unsigned)tiles, kTile, smem, st>>>(
          rb, sc, q0, q1, n, n_samples, vec, mask, nullptr, 0.0);
    }
  });
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
