// fk_cc_sm.cu - K2, production variant: fused FK -> collision as a per-lane state
// machine with work refill.
//
// Same contract as fk_cc.cu (reference Robot.is_in_collision,
// src/acrobotics/robot.py:244-273 + :206-221, geometry.py:51-66, FCL boxBox2 via
// shapes.py:50-58) and bit-identical verdicts, but organised for the statistics
// of the workload (SURVEY.md 8d / DESIGN.md): ~70 % of random configurations
// collide, half of them already at the second link, so "one thread = one
// configuration from start to end" leaves most lanes idle most of the time.
//
//  * A CTA owns a tile of kSmTile consecutive configurations.  Every lane runs a
//    small state machine (configuration, link index, running transform).  One
//    loop iteration = one link: DH step, that link's boxes against the scene,
//    self pairs.  A lane whose configuration is decided (hit, or chain finished)
//    immediately claims the next configuration of the tile from a shared
//    counter (one warp-aggregated atomic), so lanes at different links execute
//    the same instructions - the link index is data, not control flow.
//  * The robot descriptor is copied to shared memory (its arrays are laid out so
//    that different link / box indices fall into different banks).
//  * Box-vs-scene culling runs in FP32 on the FMA pipe (the FP64 pipe is the
//    bottleneck): each scene box's three face axes against the link box's
//    bounding sphere, with thresholds padded by a rigorous bound on the FP32
//    error, so a culled pair is one FCL's own FP64 face-axis test separates.
//    Everything that decides a verdict (FK, the 15-axis test) stays FP64.
//  * Verdict bits are OR-ed into a shared tile mask and flushed coalesced.
#include <cstdlib>

#include "acro_collide.cuh"
#include "acro_launch.h"

namespace acro {

constexpr int kSmTile = 4096;     // configurations per CTA
constexpr int kSmThreads = 128;
constexpr int kCullStride = 16;   // floats per scene box in the FP32 cull table
constexpr int kQPad = kMaxDof + 1;  // per-lane joint buffer stride (doubles), conflict free

// FP32 cull table entry of scene box j (built in shared memory from the FP64 box):
//  [4k..4k+3], k = 0..2 : column k of R2 and t'_k = -(R2^T T2)_k, so that the
//                         link-box centre in the box frame is y_k = col_k . c + t'_k
//  [12..14]             : B_k (1 + 1e-6) + 1e-6 |t'|_1   (padded half extents)
//  [15]                 : unused
__device__ __forceinline__ void build_cull_entry(float* e, const double* b) {
  double t[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t[k] = -(b[kOB_R + k] * b[kOB_T] + b[kOB_R + 3 + k] * b[kOB_T + 1] +
             b[kOB_R + 6 + k] * b[kOB_T + 2]);
  }
  const double tn = fabs(t[0]) + fabs(t[1]) + fabs(t[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    e[4 * k + 0] = (float)b[kOB_R + k];
    e[4 * k + 1] = (float)b[kOB_R + 3 + k];
    e[4 * k + 2] = (float)b[kOB_R + 6 + k];
    e[4 * k + 3] = (float)t[k];
    // round the padded half extent up
    e[12 + k] = __double2float_ru(b[kOB_H + k] * (1.0 + 1e-6) + 1e-6 * tn);
  }
  e[15] = 0.f;
}

// 15-axis test with a cheap first exit: box 1's face axes against box 2's
// bounding sphere (implies FCL's u_k test separates).
__device__ __forceinline__ bool pair_collides(const double A0, const double A1, const double A2,
                                              const double* R1, const double* c1,
                                              const double* b2) {
  const double p0 = b2[kOB_T + 0] - c1[0];
  const double p1 = b2[kOB_T + 1] - c1[1];
  const double p2 = b2[kOB_T + 2] - c1[2];
  const double rb = b2[kOB_RAD];
  const double z0 = fma(R1[6], p2, fma(R1[3], p1, R1[0] * p0));
  const double z1 = fma(R1[7], p2, fma(R1[4], p1, R1[1] * p0));
  const double z2 = fma(R1[8], p2, fma(R1[5], p1, R1[2] * p0));
  if ((fabs(z0) > A0 + rb) | (fabs(z1) > A1 + rb) | (fabs(z2) > A2 + rb)) return false;
  double g;
  return !sat_separated<double>(A0, A1, A2, R1, c1, b2, 0.0, g);
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(kSmThreads, MIN_CTAS)
    fk_cc_sm_kernel(const __grid_constant__ RobotDev<double> rb_param,
                    const SceneView<double> sc, const double* __restrict__ q, int64_t n,
                    int layout, uint32_t* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RobotDev<double>* rb = reinterpret_cast<RobotDev<double>*>(smem_raw);
  double* s_scene = reinterpret_cast<double*>(smem_raw + ((sizeof(RobotDev<double>) + 15) & ~15));
  float* s_cull = reinterpret_cast<float*>(s_scene + sc.n * kSceneStride);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_cull + sc.n * kCullStride);
  double* s_ql = reinterpret_cast<double*>(s_mask + kSmTile / 32);  // per-lane joint values
  __shared__ int s_next;

  {  // stage descriptor, scene and the cull table
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&rb_param);
    uint32_t* dst = reinterpret_cast<uint32_t*>(rb);
    for (int k = threadIdx.x; k < (int)(sizeof(RobotDev<double>) / 4); k += blockDim.x)
      dst[k] = src[k];
    for (int k = threadIdx.x; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];
    for (int k = threadIdx.x; k < kSmTile / 32; k += blockDim.x) s_mask[k] = 0u;
    if (threadIdx.x == 0) s_next = 0;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < sc.n; j += blockDim.x)
    build_cull_entry(s_cull + j * kCullStride, s_scene + j * kSceneStride);
  __syncthreads();

  const int64_t tile0 = (int64_t)blockIdx.x * kSmTile;
  const int tile_n = (int)((n - tile0) < kSmTile ? (n - tile0) : kSmTile);
  const int n_scene = sc.n;
  const int ndof = rb->ndof;
  const bool do_self = rb->check_self != 0;
  const unsigned full = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31;

  // slots: 0 = base boxes (no DH step), 1..ndof = DH step of link slot-1 + its boxes,
  // ndof+1 = tool boxes on the last link frame; empty base / tool slots are skipped
  const int first_slot = rb->slot_begin[1] > rb->slot_begin[0] ? 0 : 1;
  const int last_slot = rb->slot_begin[ndof + 2] > rb->slot_begin[ndof + 1] ? ndof + 1 : ndof;
  double* my_q = s_ql + threadIdx.x * kQPad;

  double saved[kMaxSaved][16];
  Tf<double> T;
  int cfg = -1;
  int slot = first_slot;
  tf_load<double>(T, rb->base);

  for (;;) {
    // ---- refill: lanes without work claim the next configurations of the tile
    const unsigned need = __ballot_sync(full, cfg < 0);
    if (need) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_next, __popc(need));
      base = __shfl_sync(full, base, 0);
      if (cfg < 0) {
        const int c = base + __popc(need & ((1u << lane) - 1u));
        if (c < tile_n) {
          cfg = c;
          slot = first_slot;
          tf_load<double>(T, rb->base);
          const int64_t row = tile0 + c;
          // all joint values of the new configuration in one burst of loads
          for (int i = 0; i < ndof; ++i)
            my_q[i] = layout == ACRO_Q_AOS ? q[row * ndof + i] : q[(int64_t)i * n + row];
        }
      }
      if (__all_sync(full, cfg < 0)) break;
    }
    if (cfg >= 0) {
      bool hit = false;
      if (slot >= 1 && slot <= ndof) chain_step<double>(T, *rb, slot - 1, my_q[slot - 1]);
      for (int b = rb->slot_begin[slot]; b < rb->slot_begin[slot + 1] && !hit; ++b) {
        const BoxDev<double>& bx = rb->boxes[b];
        double Rw[9], cw[3];
        if (bx.rot_identity) {
#pragma unroll
          for (int k = 0; k < 9; ++k) Rw[k] = T.r[k];
#pragma unroll
          for (int r = 0; r < 3; ++r)
            cw[r] = fma(T.r[3 * r + 2], bx.off[11],
                        fma(T.r[3 * r + 1], bx.off[7], fma(T.r[3 * r], bx.off[3], T.p[r])));
        } else {
          Tf<double> W;
          tf_mul<double>(W, T, bx.off);
#pragma unroll
          for (int k = 0; k < 9; ++k) Rw[k] = W.r[k];
#pragma unroll
          for (int r = 0; r < 3; ++r) cw[r] = W.p[r];
        }
        const double A0 = bx.half[0], A1 = bx.half[1], A2 = bx.half[2];
        // ---- FP32 cull against every scene box (warp-uniform j, broadcast reads)
        const float cx = (float)cw[0], cy = (float)cw[1], cz = (float)cw[2];
        // padded bounding radius: FP32 rounding of y_k is below 4 * 2^-24 * (|c|_1 + |t'|_1)
        const float ra = __double2float_ru(bx.radius * (1.0 + 1e-6)) +
                         1e-6f * (fabsf(cx) + fabsf(cy) + fabsf(cz));
        uint64_t cand = 0;
        for (int j = 0; j < n_scene; ++j) {
          const float4* e = reinterpret_cast<const float4*>(s_cull + j * kCullStride);
          const float4 c0 = e[0], c1 = e[1], c2 = e[2], bp = e[3];
          const float y0 = fmaf(c0.x, cx, fmaf(c0.y, cy, fmaf(c0.z, cz, c0.w)));
          const float y1 = fmaf(c1.x, cx, fmaf(c1.y, cy, fmaf(c1.z, cz, c1.w)));
          const float y2 = fmaf(c2.x, cx, fmaf(c2.y, cy, fmaf(c2.z, cz, c2.w)));
          const float m = fmaxf(fmaxf(fabsf(y0) - bp.x, fabsf(y1) - bp.y), fabsf(y2) - bp.z);
          if (!(m > ra)) cand |= (1ull << j);
        }
        // ---- exact FP64 test: surviving scene pairs, then the self pairs whose later
        //      member is this box (one loop, one copy of the 15-axis code)
        int pk = do_self ? rb->pair_begin[b] : 0;
        const int pend = do_self ? rb->pair_begin[b + 1] : 0;
        for (;;) {
          const double* ob;
          if (cand != 0) {
            const int j = __ffsll((long long)cand) - 1;
            cand &= cand - 1;
            ob = s_scene + j * kSceneStride;
          } else if (pk < pend) {
            ob = saved[rb->boxes[rb->partner[pk]].save_slot];
            ++pk;
          } else {
            break;
          }
          if (pair_collides(A0, A1, A2, Rw, cw, ob)) {
            hit = true;
            break;
          }
        }
        if (do_self && bx.save_slot >= 0) {
          double* ob = saved[bx.save_slot];
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
      if (hit) {
        atomicOr(&s_mask[cfg >> 5], 1u << (cfg & 31));
        cfg = -1;
      } else if (slot == last_slot) {
        cfg = -1;
      }
      ++slot;
    }
    __syncwarp(full);
  }

  __syncthreads();
  const int words = (tile_n + 31) / 32;
  for (int k = threadIdx.x; k < words; k += blockDim.x) mask[(tile0 >> 5) + k] = s_mask[k];
}

cudaError_t launch_fk_cc_sm(const RobotDev<double>& rb, SceneView<double> sc, const double* q,
                            int64_t n, int layout, uint32_t* mask, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kSmTile - 1) / kSmTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  const size_t smem = ((sizeof(RobotDev<double>) + 15) & ~size_t(15)) +
                      sizeof(double) * sc.n * kSceneStride + sizeof(float) * sc.n * kCullStride +
                      sizeof(uint32_t) * (kSmTile / 32) + sizeof(double) * kSmThreads * kQPad;
  // ACRO_K2_OCC=3 keeps all state in registers (136 regs, 12 warps/SM); the default caps the
  // kernel at 128 registers (16 warps/SM) at the price of a few spilled values
  static const bool occ3 = [] {
    const char* v = std::getenv("ACRO_K2_OCC");
    return v && v[0] == '3';
  }();
  auto kern = occ3 ? fk_cc_sm_kernel<3> : fk_cc_sm_kernel<4>;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)tiles, kSmThreads, smem, st>>>(rb, sc, q, n, layout, mask);
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
