// fk_cc_filt.cu - K2, production variant: fused FK -> collision as an FP32 *filter*
// with rigorous error pads, a candidate grid, and an exact FP64 fix-up pass.
//
// Contract and verdicts are those of fk_cc.cu (reference Robot.is_in_collision,
// src/acrobotics/robot.py:244-273 + :206-221, geometry.py:51-66, FCL boxBox2 via
// shapes.py:50-58), bit-identical to the FP64 evaluation:
//
//  * Every quantity FCL's separating-axis test compares with zero is evaluated in
//    FP32 together with a bound `delta` on its distance from the FP64 value
//    (GridView::delta, derived on the host from the size of the robot and the
//    scene).  s32 > delta on some axis  => the FP64 test separates on
//    that axis;  s32 < -delta on all 15 axes => the FP64 test finds no separating
//    axis.  Anything else is "ambiguous".  A configuration is decided when one of
//    its pairs surely collides or all of them are surely separated; otherwise its
//    bit is set in `amb_mask` and fk_cc_fixup_kernel recomputes exactly that
//    configuration in FP64 (config_collision<double>, the K2' code) and patches
//    the verdict.  For random inputs ~0.5 % of the configurations take that path.
//  * Scene candidates of a robot box come from one lookup in a precomputed grid
//    (GridView) instead of a loop over the scene; a pair the grid drops is one
//    that FCL's own face-axis test separates.
//  * One thread = one configuration (FP32 state fits in registers); the exact
//    tests of a warp's candidate pairs are redistributed over its lanes (items in
//    a small shared-memory list, poses fetched with warp shuffles), so the
//    15-axis test runs with dense warps although candidates are unevenly spread.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "acro_collide.cuh"
#include "acro_launch.h"

namespace acro {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kFThreads = 128;
constexpr int kItemCap = 64;      // candidate pairs per warp and round
#ifndef ACRO_FILTER_MIN_CTAS
#define ACRO_FILTER_MIN_CTAS 5
#endif
#ifndef ACRO_FILTER_DIRECT_MIN
#define ACRO_FILTER_DIRECT_MIN 20
#endif
constexpr int kSceneStrideF = 16; // floats per scene box: R2 row-major, T2, B, radius

enum : int { kSep = 0, kHit = 1, kAmb = 2 };

// sin/cos of a joint angle for the FP32 chain.  The argument is reduced in FP64
// (q - k pi/2, exact to ~1e-16 for |q| < 1e8), the polynomials run in FP32 on
// [-pi/4, pi/4] (Cephes sinf/cosf minimax coefficients, abs error < 1.2e-7).
__device__ __forceinline__ void sincos_filter(double q, float& s, float& c) {
  const double t = q * 0.63661977236758134308;  // 2/pi
  const int k = __double2int_rn(t);
  const double kd = (double)k;
  double r = fma(-kd, 1.57079632679489655800e+00, q);
  r = fma(-kd, 6.12323399573676603587e-17, r);
  const float x = (float)r;
  const float z = x * x;
  float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
  sp = fmaf(sp * z, x, x);
  float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
  cp = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
  const float a = (k & 1) ? cp : sp;
  const float b = (k & 1) ? sp : cp;
  s = (k & 2) ? -a : a;
  c = ((k + 1) & 2) ? -b : b;
}

// FCL boxBox2 in FP32 with tolerance: kSep when some axis separates by more than
// delta, kHit when every axis overlaps by more than delta, kAmb otherwise.
__device__ __forceinline__ int sat_filter(const float A0, const float A1, const float A2,
                                          const float* R1, const float* c1, const float* R2,
                                          const float* T2, const float B0, const float B1,
                                          const float B2, const float delta) {
  const float p0 = T2[0] - c1[0], p1 = T2[1] - c1[1], p2 = T2[2] - c1[2];
  const float A[3] = {A0, A1, A2};
  const float B[3] = {B0, B1, B2};
  float R[9], Q[9];
  float g = -1e30f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      R[3 * i + j] = fmaf(R1[6 + i], R2[6 + j], fmaf(R1[3 + i], R2[3 + j], R1[i] * R2[j]));
      Q[3 * i + j] = fabsf(R[3 * i + j]);
    }
    const float t = fmaf(R2[6 + j], p2, fmaf(R2[3 + j], p1, R2[j] * p0));
    const float rad = fmaf(Q[6 + j], A2, fmaf(Q[3 + j], A1, Q[j] * A0)) + B[j];
    g = fmaxf(g, fabsf(t) - rad);
  }
  if (g > delta) return kSep;
  float pp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pp[i] = fmaf(R1[6 + i], p2, fmaf(R1[3 + i], p1, R1[i] * p0));
#pragma unroll
  for (int i = 0; i < 3; ++i)
    g = fmaxf(g, fabsf(pp[i]) -
                     (fmaf(Q[3 * i + 2], B2, fmaf(Q[3 * i + 1], B1, Q[3 * i] * B0)) + A[i]));
  if (g > delta) return kSep;
#pragma unroll
  for (int k = 0; k < 9; ++k) Q[k] += 1.0e-6f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const float t = fmaf(pp[i2], R[3 * i1 + j], -(pp[i1] * R[3 * i2 + j]));
      const float rad = fmaf(A[i1], Q[3 * i2 + j],
                             fmaf(A[i2], Q[3 * i1 + j],
                                  fmaf(B[j1], Q[3 * i + j2], B[j2] * Q[3 * i + j1])));
      g = fmaxf(g, fabsf(t) - rad);
    }
  }
  if (g > delta) return kSep;
  return g < -delta ? kHit : kAmb;
}

template <typename MaskT>
__device__ __forceinline__ int popc_t(MaskT m);
template <>
__device__ __forceinline__ int popc_t<uint32_t>(uint32_t m) {
  return __popc(m);
}
template <>
__device__ __forceinline__ int popc_t<uint64_t>(uint64_t m) {
  return __popcll(m);
}
template <typename MaskT>
__device__ __forceinline__ int ffs_t(MaskT m);
template <>
__device__ __forceinline__ int ffs_t<uint32_t>(uint32_t m) {
  return __ffs((int)m) - 1;
}
template <>
__device__ __forceinline__ int ffs_t<uint64_t>(uint64_t m) {
  return __ffsll((long long)m) - 1;
}

}  // namespace

// ---------------------------------------------------------------------------
// the filter kernel
// ---------------------------------------------------------------------------
// Persistent warps: every warp walks over chunks of 32 configurations (chunk c = mask word c)
// with a grid-wide stride.  The joint values of the next chunk are prefetched into the warp's
// second shared-memory buffer with cp.async while the current chunk is processed; apart from
// staging the scene once per CTA nothing synchronises across warps.
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int NDOF, typename MaskT>
__global__ void __launch_bounds__(kFThreads, ACRO_FILTER_MIN_CTAS)
    fk_cc_filter_kernel(const __grid_constant__ RobotDev<float> rb, const GridView grid,
                        const SceneView<float> sc, const int n_saved,
                        const double* __restrict__ q, const int64_t n, const int layout,
                        uint32_t* __restrict__ mask, uint32_t* __restrict__ amb_mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [q buffers: 4 warps x 2 x 32*NDOF doubles][scene: n*16 floats]
  // [saved poses: 4 warps x n_saved*12*32 floats][slot half extents: 16*4 floats]
  // [item lists: 4 warps x kItemCap uint16]
  const int t = threadIdx.x;
  const unsigned lane = t & 31;
  const int warp = t >> 5;
  double* s_q = reinterpret_cast<double*>(smem_raw) + warp * 2 * 32 * NDOF;
  float* s_scene = reinterpret_cast<float*>(reinterpret_cast<double*>(smem_raw) +
                                            (kFThreads / 32) * 2 * 32 * NDOF);
  float* s_saved = s_scene + sc.n * kSceneStrideF + warp * n_saved * 12 * 32;
  float* s_half = s_scene + sc.n * kSceneStrideF + (kFThreads / 32) * n_saved * 12 * 32;
  uint16_t* list = reinterpret_cast<uint16_t*>(s_half + kMaxSaved * 4) + warp * kItemCap;
  const int n_scene = sc.n;

  // stage the scene (20-real packed layout -> 16 floats per box)
  for (int k = t; k < n_scene * kSceneStrideF; k += kFThreads) {
    const int j = k >> 4, e = k & 15;
    s_scene[k] = sc.boxes[j * kSceneStride + e];
  }
  for (int b = t; b < rb.n_boxes; b += kFThreads) {
    const int sl = rb.boxes[b].save_slot;
    if (sl >= 0) {
      s_half[4 * sl + 0] = rb.boxes[b].half[0];
      s_half[4 * sl + 1] = rb.boxes[b].half[1];
      s_half[4 * sl + 2] = rb.boxes[b].half[2];
      s_half[4 * sl + 3] = rb.boxes[b].radius;
    }
  }
  __syncthreads();

  const float delta = grid.delta;
  const bool do_self = rb.check_self != 0;
  const int64_t n_chunks = (n + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * (kFThreads / 32);
  const int ndof = NDOF;
  const int last_slot = rb.slot_begin[ndof + 2] > rb.slot_begin[ndof + 1] ? ndof + 1 : ndof;

  // asynchronous copy of one chunk's joint values into buffer `buf` (flat for AoS,
  // joint-major for SoA)
  auto prefetch = [&](int64_t chunk, int buf) {
    if (chunk < n_chunks) {
      double* dst = s_q + buf * 32 * NDOF;
      const int64_t row0 = chunk << 5;
      const int rows = (int)((n - row0) < 32 ? (n - row0) : 32);
      if (layout == ACRO_Q_AOS) {
        const double* src = q + row0 * NDOF;
#pragma unroll
        for (int j = 0; j < NDOF; ++j) {
          const int k = j * 32 + (int)lane;
          if (k < rows * NDOF) cp_async_8(dst + k, src + k);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NDOF; ++j)
          if ((int)lane < rows) cp_async_8(dst + j * 32 + lane, q + (int64_t)j * n + row0 + lane);
      }
    }
    cp_async_commit();
  };

  int64_t chunk = (int64_t)blockIdx.x * (kFThreads / 32) + warp;
  prefetch(chunk, 0);
  for (int it = 0; chunk < n_chunks; chunk += stride, ++it) {
    const int buf = it & 1;
    prefetch(chunk + stride, buf ^ 1);
    cp_async_wait<1>();
    __syncwarp();
    const double* my_q = s_q + buf * 32 * NDOF;
    const int64_t idx = (chunk << 5) + lane;
    bool done = !(idx < n);
    bool hit = false;
    bool amb = false;

    Tf<float> T;
    tf_load<float>(T, rb.base);

#pragma unroll 1
    for (int slot = 0; slot <= last_slot; ++slot) {
      if (slot >= 1 && slot <= ndof) {
        if (__all_sync(kFull, done)) break;
        const int i = slot - 1;
        double qi = 0.0;
        if (idx < n) qi = layout == ACRO_Q_AOS ? my_q[lane * NDOF + i] : my_q[i * 32 + lane];
        const bool prismatic = (rb.prismatic_mask >> i) & 1;
        // joint values the FP32 chain cannot follow within its error bound
        const double lim = prismatic ? (double)grid.q_lin_max : 1.0e8;
        if (!done && !(fabs(qi) <= lim)) {
          amb = !(grid.skip_nan && qi != qi);
          done = true;  // decided by the FP64 pass (or an absent IK branch)
          qi = 0.0;
        }
        float s, c;
        sincos_filter(prismatic ? (double)rb.theta[i] : qi, s, c);
        tf_mul_dh<float>(T, c, s, rb.ca[i], rb.sa[i], rb.a[i], prismatic ? (float)qi : rb.d[i]);
      }
      for (int b = rb.slot_begin[slot]; b < rb.slot_begin[slot + 1]; ++b) {
        const BoxDev<float>& bx = rb.boxes[b];
        float Rw[9], cw[3];
        if (bx.rot_identity) {
#pragma unroll
          for (int k = 0; k < 9; ++k) Rw[k] = T.r[k];
#pragma unroll
          for (int r = 0; r < 3; ++r)
            cw[r] = fmaf(T.r[3 * r + 2], bx.off[11],
                         fmaf(T.r[3 * r + 1], bx.off[7], fmaf(T.r[3 * r], bx.off[3], T.p[r])));
        } else {
          Tf<float> W;
          tf_mul<float>(W, T, bx.off);
#pragma unroll
          for (int k = 0; k < 9; ++k) Rw[k] = W.r[k];
#pragma unroll
          for (int r = 0; r < 3; ++r) cw[r] = W.p[r];
        }
        const float A0 = bx.half[0], A1 = bx.half[1], A2 = bx.half[2];

        // ---- candidates: scene boxes from the grid, self partners from a padded cull
        MaskT ms = 0;
        unsigned mself = 0;
        if (!done) {
          if (n_scene > 0) {
            const float fx = (cw[0] - grid.ox) * grid.inv_cell;
            const float fy = (cw[1] - grid.oy) * grid.inv_cell;
            const float fz = (cw[2] - grid.oz) * grid.inv_cell;
            const int ix = __float2int_rd(fx), iy = __float2int_rd(fy), iz = __float2int_rd(fz);
            const unsigned dim = (unsigned)grid.dim;
            if ((unsigned)ix < dim && (unsigned)iy < dim && (unsigned)iz < dim) {
              // < 2^32 cells by construction (24 classes x 160^3)
              const unsigned cell = (((unsigned)bx.cull_class * dim + iz) * dim + iy) * dim + ix;
              ms = __ldg(reinterpret_cast<const MaskT*>(grid.cells) + cell);
            } else {
              ms = n_scene >= (int)(8 * sizeof(MaskT)) ? ~MaskT(0) : ((MaskT(1) << n_scene) - 1);
            }
          }
          if (do_self) {
            for (int k = rb.pair_begin[b]; k < rb.pair_begin[b + 1]; ++k) {
              const int sl = rb.boxes[rb.partner[k]].save_slot;
              const float* sv = s_saved + sl * 12 * 32 + lane;
              // partner's face axes against this box's bounding sphere, padded
              const float d0 = cw[0] - sv[9 * 32], d1 = cw[1] - sv[10 * 32],
                          d2 = cw[2] - sv[11 * 32];
              const float ra = bx.radius + delta;
              const float y0 = fmaf(sv[6 * 32], d2, fmaf(sv[3 * 32], d1, sv[0] * d0));
              const float y1 = fmaf(sv[7 * 32], d2, fmaf(sv[4 * 32], d1, sv[1 * 32] * d0));
              const float y2 = fmaf(sv[8 * 32], d2, fmaf(sv[5 * 32], d1, sv[2 * 32] * d0));
              const bool sep = (fabsf(y0) > s_half[4 * sl + 0] + ra) |
                               (fabsf(y1) > s_half[4 * sl + 1] + ra) |
                               (fabsf(y2) > s_half[4 * sl + 2] + ra);
              if (!sep) mself |= 1u << sl;
            }
          }
        }

        // ---- rounds over the warp's candidate pairs.  A round is "direct" (every lane
        // tests its own next candidate with its own pose: no data movement) when most lanes
        // have one, or no lane has more than one; otherwise the remaining pairs of the whole
        // warp are redistributed over its lanes (items in a shared-memory list, the owner's
        // pose fetched with shuffles) so that the 15-axis test still runs on dense warps.
        auto load_other = [&](int other, int owner, float* R2, float* T2, float& B0, float& B1,
                              float& B2) {
          if (other < 64) {
            const float* ob = s_scene + other * kSceneStrideF;
#pragma unroll
            for (int e = 0; e < 9; ++e) R2[e] = ob[e];
            T2[0] = ob[9]; T2[1] = ob[10]; T2[2] = ob[11];
            B0 = ob[12]; B1 = ob[13]; B2 = ob[14];
          } else {
            const int sl = other - 64;
            const float* sv = s_saved + sl * 12 * 32 + owner;
#pragma unroll
            for (int e = 0; e < 9; ++e) R2[e] = sv[e * 32];
            T2[0] = sv[9 * 32]; T2[1] = sv[10 * 32]; T2[2] = sv[11 * 32];
            B0 = s_half[4 * sl]; B1 = s_half[4 * sl + 1]; B2 = s_half[4 * sl + 2];
          }
        };
        for (;;) {
          const int c = done ? 0 : popc_t<MaskT>(ms) + __popc(mself);
          const unsigned b1 = __ballot_sync(kFull, c > 0);
          if (b1 == 0u) break;
          const unsigned b2 = __ballot_sync(kFull, c > 1);
          if (b2 == 0u || __popc(b1) >= ACRO_FILTER_DIRECT_MIN) {
            if (c > 0) {
              int other;
              if (ms != 0) {
                other = ffs_t<MaskT>(ms);
                ms &= ms - 1;
              } else {
                other = 64 + __ffs((int)mself) - 1;
                mself &= mself - 1;
              }
              float R2[9], T2[3], B0, B1, B2;
              load_other(other, (int)lane, R2, T2, B0, B1, B2);
              const int res = sat_filter(A0, A1, A2, Rw, cw, R2, T2, B0, B1, B2, delta);
              if (res == kHit) {
                hit = true;
                done = true;
              } else if (res == kAmb) {
                amb = true;
              }
            }
            __syncwarp();
            continue;
          }
          int incl = c;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, d);
            if ((int)lane >= d) incl += v;
          }
          const int total = __shfl_sync(kFull, incl, 31);
          int pos = incl - c;
          if (!done) {
            while (pos < kItemCap && ms != 0) {
              const int j = ffs_t<MaskT>(ms);
              ms &= ms - 1;
              list[pos++] = (uint16_t)((lane << 8) | j);
            }
            while (pos < kItemCap && mself != 0) {
              const int j = __ffs((int)mself) - 1;
              mself &= mself - 1;
              list[pos++] = (uint16_t)((lane << 8) | (64 + j));
            }
          }
          __syncwarp();
          const int cnt = total < kItemCap ? total : kItemCap;
          for (int base = 0; base < cnt; base += 32) {
            const int k = base + (int)lane;
            const bool valid = k < cnt;
            const unsigned item = valid ? list[k] : (lane << 8);
            const int owner = item >> 8, other = item & 255;
            float R1[9], c1[3];
#pragma unroll
            for (int e = 0; e < 9; ++e) R1[e] = __shfl_sync(kFull, Rw[e], owner);
#pragma unroll
            for (int e = 0; e < 3; ++e) c1[e] = __shfl_sync(kFull, cw[e], owner);
            const bool owner_done = __shfl_sync(kFull, (int)done, owner) != 0;
            int res = kSep;
            if (valid && !owner_done) {
              float R2[9], T2[3], B0, B1, B2;
              load_other(other, owner, R2, T2, B0, B1, B2);
              res = sat_filter(A0, A1, A2, R1, c1, R2, T2, B0, B1, B2, delta);
            }
            const unsigned hb = __reduce_or_sync(kFull, res == kHit ? (1u << owner) : 0u);
            const unsigned ab = __reduce_or_sync(kFull, res == kAmb ? (1u << owner) : 0u);
            if ((hb >> lane) & 1u) {
              hit = true;
              done = true;
            }
            if ((ab >> lane) & 1u) amb = true;
          }
          __syncwarp();
        }

        if (do_self && bx.save_slot >= 0) {
          float* sv = s_saved + bx.save_slot * 12 * 32 + lane;
#pragma unroll
          for (int e = 0; e < 9; ++e) sv[e * 32] = Rw[e];
#pragma unroll
          for (int e = 0; e < 3; ++e) sv[(9 + e) * 32] = cw[e];
        }
      }
    }

    const unsigned word = __ballot_sync(kFull, hit);
    const unsigned aword = __ballot_sync(kFull, amb && !hit);
    if (lane == 0) {
      mask[chunk] = word;
      amb_mask[chunk] = aword;
    }
    __syncwarp();  // every lane is done with buffer `buf` before it is refilled
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// exact pass: recompute the ambiguous configurations in FP64
// ---------------------------------------------------------------------------
constexpr int kFixThreads = 256;  // one mask word (32 configurations) per thread

template <int NDOF>
__global__ void __launch_bounds__(kFixThreads)
    fk_cc_fixup_kernel(const __grid_constant__ RobotDev<double> rb, const SceneView<double> sc,
                       const double* __restrict__ q, const int64_t n, const int layout,
                       uint32_t* __restrict__ mask, const uint32_t* __restrict__ amb_mask,
                       const int64_t n_words) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_scene = reinterpret_cast<double*>(smem_raw);
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_scene + sc.n * kSceneStride);  // [256*32]
  __shared__ int s_warp_tot[kFixThreads / 32];
  __shared__ int s_total;
  const int t = threadIdx.x;
  const unsigned lane = t & 31;
  const int64_t w = (int64_t)blockIdx.x * kFixThreads + t;
  unsigned bits = w < n_words ? amb_mask[w] : 0u;
  // CTA-wide exclusive scan of the bit counts
  const int c = __popc(bits);
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(kFull, incl, d);
    if ((int)lane >= d) incl += v;
  }
  if (lane == 31) s_warp_tot[t >> 5] = incl;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int k = 0; k < kFixThreads / 32; ++k) {
      const int v = s_warp_tot[k];
      s_warp_tot[k] = run;
      run += v;
    }
    s_total = run;
  }
  __syncthreads();
  const int total = s_total;
  if (total == 0) return;  // CTA-uniform
  int pos = s_warp_tot[t >> 5] + incl - c;
  while (bits) {
    const int b = __ffs((int)bits) - 1;
    bits &= bits - 1;
    s_list[pos++] = (uint16_t)(t * 32 + b);
  }
  for (int k = t; k < sc.n * kSceneStride; k += kFixThreads) s_scene[k] = sc.boxes[k];
  __syncthreads();
  const int64_t cfg0 = (int64_t)blockIdx.x * kFixThreads * 32;
  for (int base = 0; base < total; base += kFixThreads) {
    const int k = base + t;
    const bool act = k < total;
    const int64_t idx = cfg0 + (act ? s_list[k] : 0);
    double qv[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i)
      qv[i] = act ? (layout == ACRO_Q_AOS ? q[idx * NDOF + i] : q[(int64_t)i * n + idx]) : 0.0;
    Verdict<double, false> v;
    if (__any_sync(kFull, act))
      config_collision<double, NDOF, false>(rb, qv, s_scene, sc.n, 0.0, act, v);
    if (act && v.hit) atomicOr(&mask[idx >> 5], 1u << (idx & 31));
  }
}

// ---------------------------------------------------------------------------
// candidate grid construction
// ---------------------------------------------------------------------------
template <typename MaskT>
__global__ void __launch_bounds__(256)
    build_grid_kernel(const SceneView<double> sc, MaskT* __restrict__ cells, int dim,
                      int n_classes, double ox, double oy, double oz, double cell,
                      const double* __restrict__ class_radius, double pad) {
  const int64_t per = (int64_t)dim * dim * dim;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= per * n_classes) return;
  const int cls = (int)(gid / per);
  const int64_t r = gid - cls * per;
  const int ix = (int)(r % dim), iy = (int)((r / dim) % dim), iz = (int)(r / ((int64_t)dim * dim));
  const double px = ox + (ix + 0.5) * cell, py = oy + (iy + 0.5) * cell, pz = oz + (iz + 0.5) * cell;
  const double reach = class_radius[cls] + pad;
  MaskT m = 0;
  for (int j = 0; j < sc.n; ++j) {
    const double* b = sc.boxes + (size_t)j * kSceneStride;
    const double d0 = px - b[9], d1 = py - b[10], d2 = pz - b[11];
    bool sep = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double y = b[k] * d0 + b[3 + k] * d1 + b[6 + k] * d2;
      sep |= fabs(y) > b[12 + k] + reach;
    }
    if (!sep) m |= MaskT(1) << j;
  }
  cells[gid] = m;
}

cudaError_t launch_build_grid(SceneView<double> sc, void* cells, int mask_bits, int dim,
                              int n_classes, const double* origin, double cell,
                              const double* class_radius_dev, double pad, cudaStream_t st) {
  const int64_t total = (int64_t)dim * dim * dim * n_classes;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (mask_bits == 32)
    build_grid_kernel<uint32_t><<<blocks, 256, 0, st>>>(sc, (uint32_t*)cells, dim, n_classes,
                                                       origin[0], origin[1], origin[2], cell,
                                                       class_radius_dev, pad);
  else
    build_grid_kernel<uint64_t><<<blocks, 256, 0, st>>>(sc, (uint64_t*)cells, dim, n_classes,
                                                       origin[0], origin[1], origin[2], cell,
                                                       class_radius_dev, pad);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------
size_t filter_smem_bytes(int ndof, int n_scene, int n_saved) {
  const int warps = kFThreads / 32;
  return sizeof(double) * warps * 2 * 32 * ndof + sizeof(float) * n_scene * kSceneStrideF +
         sizeof(float) * warps * n_saved * 12 * 32 + sizeof(float) * kMaxSaved * 4 +
         sizeof(uint16_t) * warps * kItemCap;
}

// resident CTAs of the persistent filter kernel on the current device
template <typename K>
static int persistent_blocks(K kern, size_t smem) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFThreads, smem);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

cudaError_t launch_fk_cc_filter(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                const GridView& grid, int mask_bits, SceneView<float> sc32,
                                SceneView<double> sc64, const double* q, int64_t n, int layout,
                                uint32_t* mask, uint32_t* amb_mask, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kFThreads - 1) / kFThreads;
  int n_saved = 0;
  for (int b = 0; b < rb32.n_boxes; ++b)
    if (rb32.boxes[b].save_slot + 1 > n_saved) n_saved = rb32.boxes[b].save_slot + 1;
  if (!rb32.check_self) n_saved = 0;
  const int64_t n_words = (n + 31) / 32;
  cudaError_t e = cudaSuccess;
  ACRO_DISPATCH_NDOF(rb32.ndof, {
    const size_t smem = filter_smem_bytes(NDOF, sc32.n, n_saved);
    if (mask_bits == 32) {
      auto kern = fk_cc_filter_kernel<NDOF, uint32_t>;
      if (smem > 48 * 1024)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      const int64_t blocks = std::min<int64_t>(tiles, persistent_blocks(kern, smem));
      kern<<<(unsigned)blocks, kFThreads, smem, st>>>(rb32, grid, sc32, n_saved, q, n, layout, mask,
                                                      amb_mask);
    } else {
      auto kern = fk_cc_filter_kernel<NDOF, uint64_t>;
      if (smem > 48 * 1024)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      const int64_t blocks = std::min<int64_t>(tiles, persistent_blocks(kern, smem));
      kern<<<(unsigned)blocks, kFThreads, smem, st>>>(rb32, grid, sc32, n_saved, q, n, layout, mask,
                                                      amb_mask);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t fsmem = sizeof(double) * sc64.n * kSceneStride + sizeof(uint16_t) * kFixThreads * 32;
    const int64_t fblocks = (n_words + kFixThreads - 1) / kFixThreads;
    fk_cc_fixup_kernel<NDOF><<<(unsigned)fblocks, kFixThreads, fsmem, st>>>(
        rb64, sc64, q, n, layout, mask, amb_mask, n_words);
    count_launch();
  });
  return cudaGetLastError();
}

}  // namespace acro
