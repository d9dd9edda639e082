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
//    the verdict.  For random inputs ~0.1 % of the configurations take that path.
//  * Scene candidates of a robot box come from one lookup in a precomputed grid
//    (GridView) instead of a loop over the scene; a pair the grid drops is one
//    that FCL's own face-axis test separates.
//  * One thread = one configuration (FP32 state fits in registers); the exact
//    tests of a warp's candidate pairs are redistributed over its lanes (items in
//    a small shared-memory list, poses fetched with warp shuffles), so the
//    15-axis test runs with dense warps although candidates are unevenly spread.
//  * Two passes: the slots up to `split_slot` for a fresh chunk, the rest for 32 still
//    undecided configurations regrouped from several chunks (see the kernel).
#include <algorithm>
#include <cstdlib>
#include <string>

#include "acro_filter.cuh"
#include "acro_launch.h"

namespace acro {

namespace {

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

// Two passes per configuration (split_slot > 0).  Most decisions fall in the first links (52 %
// of the benchmark's configurations collide at the second one), after which a warp that keeps
// its 32 configurations together runs the remaining links with a third of its lanes.  So:
//   pass 1  the chain up to slot split_slot for a fresh chunk, testing the slots of early_mask;
//           the chunk's mask words are written; the indices of the undecided configurations
//           go to a per-warp stash;
//   pass 2  whenever the stash holds 32 entries: those 32 configurations (from several chunks)
//           redo the cheap FP32 chain from the start, test every slot pass 1 did not (early
//           slots outside early_mask - by default the base and the first link, which hardly
//           ever decide a configuration - and everything after split_slot) with full lanes,
//           and OR late verdicts into the mask words with atomics.
// split_slot = 0 keeps the single pass (slot windows of filter_chunk).
// CTA size.  Each warp needs ~8 KB of shared memory (prefetch buffers, stash, saved poses) and
// each CTA another 3.3 KB (scene, half extents, the driver's 1 KB), so with 128-thread CTAs six
// fit an SM (24 warps, although the 72-register build allows 28): fewer, larger CTAs amortise
// the per-CTA part - 9 warps x 3 CTAs = 27 warps, 14 warps x 2 CTAs = 28 warps for Kuka against
// a 16-box scene (+5.6 % throughput).  The launcher instantiates all three sizes and takes the
// one with the most resident warps for the robot and scene at hand.
#ifndef ACRO_K2_WARPS
#define ACRO_K2_WARPS 0  // 0 = choose at launch; 4 / 9 / 14 = force (tuning)
#endif
typedef unsigned k2_stash_t;  // configuration indices below 2^32 (checked by the launcher)
constexpr int k2_min_ctas(int warps) { return warps <= 4 ? 7 : (warps <= 9 ? 3 : 2); }
// do the per-warp static arrays (prefetch buffers, stash, list) fit the 48 KB static limit?
__host__ __device__ constexpr bool k2_static_q(int ndof, int warps) {
  return (size_t)warps * (sizeof(double) * 2 * 32 * ndof + sizeof(k2_stash_t) * 64 +
                          sizeof(uint16_t) * kItemCap) <= 48 * 1024;
}

template <int NDOF, typename MaskT, bool PRISM, bool ROTID, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, k2_min_ctas(WARPS))
    fk_cc_filter_kernel(const __grid_constant__ RobotDev<float> rb, const GridView grid,
                        const SceneView<float> sc, const int n_saved, const int split_slot,
                        const unsigned early_mask, const double* __restrict__ q, const int64_t n, const int layout,
                        uint32_t* __restrict__ mask, uint32_t* __restrict__ amb_mask,
                        unsigned long long* __restrict__ work_counter) {
  constexpr int kK2Threads = WARPS * 32;
  // Shared memory.  The small per-warp arrays are STATIC, so that their addresses are an
  // immediate plus the warp / lane offset: with one dynamic block carved by hand the compiler,
  // short of registers, rematerialised the base pointers 19 times per chunk (S2R + 8 uniform
  // instructions each: 8 % of all issued instructions).  The prefetch buffers sit at a
  // compile-time offset of the dynamic block (they would exceed the 48 KB static limit in a
  // 14-warp CTA); the saved poses of the self-collision partners (n_saved is a property of
  // the robot) follow them.  Kuka x 16 boxes: 8 KB per warp + 2.3 KB per CTA.  (Keeping the joint rows of the stashed
  // configurations in shared memory instead of re-reading them from L2 in pass 2 drops
  // residency by one CTA per SM and was measured 9 % slower.)
  constexpr int kWarps = WARPS;
  constexpr int kSceneCap = sizeof(MaskT) == 4 ? 32 : kMaxScene;  // a 32-bit mask means <= 32 boxes
  // static while it fits the 48 KB static limit: the per-warp prefetch double buffers, stashes
  // and candidate lists (14 warps x 3 456 B for six joints).  Dynamic, in this order: the
  // half extents of the saved slots and the scene (CTA wide, at compile-time offsets), the
  // prefetch buffers when they did not fit above, the saved poses of the self-collision
  // partners ([kWarps][n_saved][12 * 32] floats).
  constexpr bool kStaticQ = k2_static_q(NDOF, WARPS);
  __shared__ __align__(16) double s_q_all[kStaticQ ? kWarps : 1][kStaticQ ? 2 * 32 * NDOF : 2];
  __shared__ k2_stash_t s_stash_all[kWarps][64];
  __shared__ uint16_t s_list_all[kWarps][kItemCap];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_half_st = reinterpret_cast<float*>(smem_raw);
  float* s_scene_st = s_half_st + kMaxSaved * 4;
  constexpr size_t kDynQ = sizeof(float) * (kMaxSaved * 4 + kSceneCap * kSceneStrideF);
  constexpr size_t kDynSaved = kDynQ + (kStaticQ ? 0 : sizeof(double) * kWarps * 2 * 32 * NDOF);
  const int t = threadIdx.x;
#if ACRO_FILTER_PIN_IDS
  // opaque to the optimiser: keeps the lane / warp index in a register instead of re-reading
  // SR_TID and re-deriving every shared-memory address from it inside the loops
  unsigned lane = t & 31;
  int warp = t >> 5;
  asm volatile("" : "+r"(lane));
  asm volatile("" : "+r"(warp));
#else
  const unsigned lane = t & 31;
  const int warp = t >> 5;
#endif
  for (int k = t; k < sc.n * kSceneStrideF; k += kK2Threads) {
    const int j = k >> 4, e = k & 15;
    s_scene_st[k] = sc.boxes[j * kSceneStride + kScenePerm[e]];
  }
  for (int b = t; b < rb.n_boxes; b += kK2Threads) {
    const int sl = rb.boxes[b].save_slot;
    if (sl >= 0) {
      s_half_st[4 * sl + 0] = rb.boxes[b].half[0];
      s_half_st[4 * sl + 1] = rb.boxes[b].half[1];
      s_half_st[4 * sl + 2] = rb.boxes[b].half[2];
      s_half_st[4 * sl + 3] = rb.boxes[b].radius;
    }
  }
  FilterCtx ctx;
  ctx.s_scene = s_scene_st;
  ctx.s_half = s_half_st;
  ctx.list = s_list_all[warp];
  ctx.s_saved = reinterpret_cast<float*>(smem_raw + kDynSaved) + warp * n_saved * 12 * 32;
  ctx.extra = nullptr;
  ctx.n_scene = sc.n;
  double* s_q = kStaticQ ? s_q_all[kStaticQ ? warp : 0]
                         : reinterpret_cast<double*>(smem_raw + kDynQ) + warp * (2 * 32 * NDOF);
  k2_stash_t* s_stash = s_stash_all[warp];
  __syncthreads();

  const int64_t n_chunks = (n + 31) >> 5;
#if !ACRO_FILTER_DYNAMIC
  const int64_t stride = (int64_t)gridDim.x * kWarps;
#endif

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

  // One loop, one call site of filter_chunk (the kernel body must stay small enough for the
  // instruction cache): an iteration is either pass 1 over the next chunk or pass 2 over 32
  // stashed configurations (or over the remainder once the chunks are exhausted).
  // pass 2 keeps its joint values (joint-major [NDOF][32]) in the prefetch buffer of the chunk
  // pass 1 has just finished with: that buffer is only refilled by the next pass-1 iteration
  int stashed = 0;                   // warp-uniform
  int it = 0;
#if ACRO_FILTER_DYNAMIC
  // Chunks are handed out by a grid-wide counter (one atomic per chunk, fetched one iteration
  // ahead so that its latency hides behind the chunk being processed): the cost of a chunk
  // varies a lot (pass 2, candidate rounds), and with a static stride the slowest warp of a
  // small batch finishes long after the average one.
  auto grab = [&]() -> int64_t {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(work_counter, 1ull);
    return (int64_t)__shfl_sync(kFull, v, 0);
  };
  int64_t chunk = grab();
  int64_t chunk_next = grab();
#else
  int64_t chunk = (int64_t)blockIdx.x * kWarps + warp;
#endif
  prefetch(chunk, 0);
  for (;;) {
    const bool second = split_slot > 0 && (stashed >= 32 || (chunk >= n_chunks && stashed > 0));
    if (!second && chunk >= n_chunks) break;
    int64_t idx;
    bool act;
    const double* qp;
    int qs, slot_end, m = 0;
    unsigned check_mask;
    if (second) {
      m = stashed < 32 ? stashed : 32;
      act = (int)lane < m;
      idx = act ? (int64_t)s_stash[stashed - m + lane] : 0;
      double* q2 = s_q + ((it + 1) & 1) * 32 * NDOF;
      if (act) {
#pragma unroll
        for (int i = 0; i < NDOF; ++i)
          q2[i * 32 + lane] = layout == ACRO_Q_AOS ? __ldg(q + idx * NDOF + i)
                                                   : __ldg(q + (int64_t)i * n + idx);
      }
      qp = q2 + lane;
      qs = 32;
      check_mask = ~early_mask;
      slot_end = -1;
    } else {
      const int buf = it & 1;
#if ACRO_FILTER_DYNAMIC
      prefetch(chunk_next, buf ^ 1);
#else
      prefetch(chunk + stride, buf ^ 1);
#endif
      cp_async_wait<1>();
      idx = (chunk << 5) + lane;
      act = idx < n;
      // this lane's joint i sits at qp[i * qs] (row of the flat AoS copy, column of the SoA one)
      qp = s_q + buf * 32 * NDOF + (layout == ACRO_Q_AOS ? lane * NDOF : lane);
      qs = layout == ACRO_Q_AOS ? 1 : 32;
      check_mask = split_slot > 0 ? early_mask : ~0u;
      slot_end = split_slot > 0 ? split_slot : -1;
    }
    __syncwarp();

    bool hit, amb, retired;
    filter_chunk<NDOF, MaskT, PRISM, ROTID>(rb, grid, ctx, lane, act, [=](int i) { return qp[i * qs]; },
                                     hit, amb, retired, check_mask, slot_end);

    if (second) {
      if (act && hit) atomicOr(&mask[idx >> 5], 1u << (idx & 31));
      if (act && amb && !hit) atomicOr(&amb_mask[idx >> 5], 1u << (idx & 31));
      stashed -= m;
    } else {
      const unsigned word = __ballot_sync(kFull, hit);
      const unsigned aword = __ballot_sync(kFull, amb && !hit);
      if (lane == 0) {
        mask[chunk] = word;
        amb_mask[chunk] = aword;
      }
      if (split_slot > 0) {
        // undecided and unambiguous configurations continue in pass 2
        const bool cont = act && !retired && !amb;
        const unsigned cm = __ballot_sync(kFull, cont);
        if (cont) s_stash[stashed + __popc(cm & ((1u << lane) - 1u))] = (k2_stash_t)idx;
        stashed += __popc(cm);
      }
#if ACRO_FILTER_DYNAMIC
      chunk = chunk_next;
      chunk_next = chunk_next < n_chunks ? grab() : chunk_next;
#else
      chunk += stride;
#endif
      ++it;
    }
    __syncwarp();  // buffers, stash entries and this chunk's mask words are settled
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------
// exact pass: recompute the ambiguous configurations in FP64
// ---------------------------------------------------------------------------
constexpr int kFixThreads = 256;
constexpr int kFixWords = 4;  // mask words per thread: a CTA compacts 32768 configurations, so
                              // that even at 0.1 % ambiguity its FP64 pass runs on a full warp

// MARGIN: also report the configurations whose decisive separation is within +-margin of
// contact (the K2' margin evaluation); the filter then ran with tolerance delta + margin, so
// every such configuration is among the flagged ones.
template <int NDOF, bool MARGIN>
__global__ void __launch_bounds__(kFixThreads)
    fk_cc_fixup_kernel(const __grid_constant__ RobotDev<double> rb, const SceneView<double> sc,
                       const double* __restrict__ q, const int64_t n, const int layout,
                       uint32_t* __restrict__ mask, const uint32_t* __restrict__ amb_mask,
                       const int64_t n_words, uint32_t* __restrict__ near_mask,
                       const double margin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_scene = reinterpret_cast<double*>(smem_raw);
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_scene + sc.n * kSceneStride);  // [256*4*32]
  __shared__ int s_warp_tot[kFixThreads / 32];
  __shared__ int s_total;
  const int t = threadIdx.x;
  const unsigned lane = t & 31;
  // thread t owns words w0 + t + 256 j (coalesced reads); list order does not matter
  const int64_t w0 = (int64_t)blockIdx.x * kFixThreads * kFixWords;
  unsigned bits[kFixWords];
  int c = 0;
#pragma unroll
  for (int j = 0; j < kFixWords; ++j) {
    const int64_t w = w0 + t + j * kFixThreads;
    bits[j] = w < n_words ? amb_mask[w] : 0u;
    c += __popc(bits[j]);
  }
  // CTA-wide exclusive scan of the bit counts
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
#pragma unroll
  for (int j = 0; j < kFixWords; ++j) {
    unsigned bj = bits[j];
    while (bj) {
      const int b = __ffs((int)bj) - 1;
      bj &= bj - 1;
      s_list[pos++] = (uint16_t)((t + j * kFixThreads) * 32 + b);
    }
  }
  for (int k = t; k < sc.n * kSceneStride; k += kFixThreads) s_scene[k] = sc.boxes[k];
  __syncthreads();
  const int64_t cfg0 = w0 * 32;
  for (int base = 0; base < total; base += kFixThreads) {
    const int k = base + t;
    const bool act = k < total;
    const int64_t idx = cfg0 + (act ? s_list[k] : 0);
    double qv[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i)
      qv[i] = act ? (layout == ACRO_Q_AOS ? q[idx * NDOF + i] : q[(int64_t)i * n + idx]) : 0.0;
    Verdict<double, MARGIN> v;
    if (__any_sync(kFull, act))
      config_collision<double, NDOF, MARGIN>(rb, qv, s_scene, sc.n, margin, act, v);
    if (act && v.hit) atomicOr(&mask[idx >> 5], 1u << (idx & 31));
    if (MARGIN && act && v.gmin <= margin && v.gmin >= -margin)
      atomicOr(&near_mask[idx >> 5], 1u << (idx & 31));
  }
}

// ---------------------------------------------------------------------------
// candidate grid construction
// ---------------------------------------------------------------------------
template <typename MaskT>
__global__ void __launch_bounds__(256)
    build_grid_kernel(const SceneView<double> sc, MaskT* __restrict__ cells, int dim,
                      int n_classes, double ox, double oy, double oz, double cell,
                      const GridClassRadii class_radius, double pad) {
  const int64_t per = (int64_t)dim * dim * dim;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= per * n_classes) return;
  const int cls = (int)(gid / per);
  const int64_t r = gid - cls * per;
  const int ix = (int)(r % dim), iy = (int)((r / dim) % dim), iz = (int)(r / ((int64_t)dim * dim));
  const double px = ox + (ix + 0.5) * cell, py = oy + (iy + 0.5) * cell, pz = oz + (iz + 0.5) * cell;
  const double reach = class_radius.r[cls] + pad;
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
                              const double* class_radius_host, double pad, cudaStream_t st) {
  GridClassRadii radii;
  for (int c = 0; c < kMaxBoxes; ++c) radii.r[c] = c < n_classes ? class_radius_host[c] : 0.0;
  const int64_t total = (int64_t)dim * dim * dim * n_classes;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (mask_bits == 32)
    build_grid_kernel<uint32_t><<<blocks, 256, 0, st>>>(sc, (uint32_t*)cells, dim, n_classes,
                                                       origin[0], origin[1], origin[2], cell,
                                                       radii, pad);
  else
    build_grid_kernel<uint64_t><<<blocks, 256, 0, st>>>(sc, (uint64_t*)cells, dim, n_classes,
                                                       origin[0], origin[1], origin[2], cell,
                                                       radii, pad);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------
size_t filter_scratch_bytes(int64_t n) {
  const size_t words = (size_t)((n + 31) / 32);
  return ((words * sizeof(uint32_t) + 7) / 8) * 8 + 8;
}

size_t filter_smem_bytes(int ndof, int mask_bits, int n_saved, int warps) {
  // the dynamic part: half extents of the saved slots, the scene, the prefetch buffers when
  // they do not fit the static limit, the saved poses of the self-collision partners
  const int scene_cap = mask_bits == 32 ? 32 : kMaxScene;
  return sizeof(float) * (kMaxSaved * 4 + (size_t)scene_cap * kSceneStrideF) +
         (k2_static_q(ndof, warps) ? 0 : (size_t)warps * sizeof(double) * 2 * 32 * ndof) +
         (size_t)warps * sizeof(float) * (size_t)n_saved * 12 * 32;
}

// resident CTAs per SM of a filter kernel instantiation with `smem` dynamic bytes (0: does not fit)
template <typename K>
static int resident_ctas(K kern, int threads, size_t smem) {
  int per_sm = 0;
  // static + dynamic shared memory beyond 48 KB needs the opt-in (the static part alone is 47 KB
  // in the 14-warp instantiation)
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return per_sm;
}

cudaError_t launch_fk_cc_filter(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                const GridView& grid, int mask_bits, SceneView<float> sc32,
                                SceneView<double> sc64, const double* q, int64_t n, int layout,
                                uint32_t* mask, uint32_t* scratch, uint32_t* near_mask,
                                double margin, int split_slot, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  // scratch: ambiguity mask words, then (8-byte aligned) the chunk counter of the filter kernel
  uint32_t* amb_mask = scratch;
  unsigned long long* work_counter = reinterpret_cast<unsigned long long*>(
      reinterpret_cast<unsigned char*>(scratch) + filter_scratch_bytes(n) - 8);
  {
    cudaError_t e0 = cudaMemsetAsync(work_counter, 0, 8, st);
    if (e0 != cudaSuccess) return e0;
  }
  if (n > 0xffffffffLL) return cudaErrorInvalidValue;  // stash entries are 32-bit indices
  int n_saved = 0;
  for (int b = 0; b < rb32.n_boxes; ++b)
    if (rb32.boxes[b].save_slot + 1 > n_saved) n_saved = rb32.boxes[b].save_slot + 1;
  if (!rb32.check_self) n_saved = 0;
  cudaError_t e = cudaSuccess;
  // pass 1 tests slots 2 .. split_slot; the base (slot 0) and the first link (slot 1) wait for
  // pass 2 unless the split is so early that nothing else is left for pass 1
  unsigned early_mask = split_slot > 0 ? ((2u << split_slot) - 1u) : ~0u;
  if (split_slot >= 2) early_mask &= ~3u;
  int sms = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const bool prism = rb32.prismatic_mask != 0;
  bool rotid = !prism;  // the specialised instantiation exists for all-revolute chains
  for (int b = 0; b < rb32.n_boxes; ++b) rotid = rotid && rb32.boxes[b].rot_identity != 0;
#define ACRO_K2_PICK(M, P, R)                                                             \
  pick(fk_cc_filter_kernel<NDOF, M, P, R, 4>, fk_cc_filter_kernel<NDOF, M, P, R, 9>,      \
       fk_cc_filter_kernel<NDOF, M, P, R, 14>)
  ACRO_DISPATCH_NDOF(rb32.ndof, {
    // one CTA size: resident CTAs per SM (0 = does not fit), launch
    auto resident = [&](auto kern, int warps) {
      return resident_ctas(kern, warps * 32, filter_smem_bytes(NDOF, mask_bits, n_saved, warps));
    };
    auto launch = [&](auto kern, int warps, int per_sm) -> cudaError_t {
      const size_t smem = filter_smem_bytes(NDOF, mask_bits, n_saved, warps);
      const int64_t tiles = (n + warps * 32 - 1) / (warps * 32);
      const int64_t blocks = std::min<int64_t>(tiles, (int64_t)sms * per_sm);
      kern<<<(unsigned)blocks, warps * 32, smem, st>>>(rb32, grid, sc32, n_saved, split_slot,
                                                       early_mask, q, n, layout, mask, amb_mask,
                                                       work_counter);
      return cudaGetLastError();
    };
    // the CTA size with the most resident warps (ties: the smaller CTA)
    auto pick = [&](auto k4, auto k9, auto k14) -> cudaError_t {
      const int forced = ACRO_K2_WARPS;
      const int r4 = (forced == 0 || forced == 4) ? resident(k4, 4) : 0;
      const int r9 = (forced == 0 || forced == 9) ? resident(k9, 9) : 0;
      const int r14 = (forced == 0 || forced == 14) ? resident(k14, 14) : 0;
      const int w4 = 4 * r4, w9 = 9 * r9, w14 = 14 * r14;
      if (w4 == 0 && w9 == 0 && w14 == 0) return cudaErrorInvalidConfiguration;
      if (w14 > w9 && w14 > w4) return launch(k14, 14, r14);
      if (w9 > w4) return launch(k9, 9, r9);
      return launch(k4, 4, r4);
    };
    if (mask_bits == 32)
      e = prism   ? ACRO_K2_PICK(uint32_t, true, false)
          : rotid ? ACRO_K2_PICK(uint32_t, false, true)
                  : ACRO_K2_PICK(uint32_t, false, false);
    else
      e = prism   ? ACRO_K2_PICK(uint64_t, true, false)
          : rotid ? ACRO_K2_PICK(uint64_t, false, true)
                  : ACRO_K2_PICK(uint64_t, false, false);
    if (e != cudaSuccess) return e;
    count_launch();
  });
#undef ACRO_K2_PICK
  if (e != cudaSuccess) return e;
  return launch_fk_cc_fixup(rb64, sc64, q, n, layout, mask, amb_mask, near_mask, margin, st);
}

// exact FP64 pass over the configurations flagged in amb_mask (near_mask != nullptr: margin
// mode, near_mask must be zeroed by the caller)
cudaError_t launch_fk_cc_fixup(const RobotDev<double>& rb64, SceneView<double> sc64, const double* q,
                               int64_t n, int layout, uint32_t* mask, const uint32_t* amb_mask,
                               uint32_t* near_mask, double margin, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t n_words = (n + 31) / 32;
  ACRO_DISPATCH_NDOF(rb64.ndof, {
    const size_t fsmem = sizeof(double) * sc64.n * kSceneStride +
                         sizeof(uint16_t) * kFixThreads * kFixWords * 32;
    const int64_t per_cta = (int64_t)kFixThreads * kFixWords;
    const int64_t fblocks = (n_words + per_cta - 1) / per_cta;
    if (near_mask) {
      auto kern = fk_cc_fixup_kernel<NDOF, true>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      kern<<<(unsigned)fblocks, kFixThreads, fsmem, st>>>(rb64, sc64, q, n, layout, mask, amb_mask,
                                                          n_words, near_mask, margin);
    } else {
      auto kern = fk_cc_fixup_kernel<NDOF, false>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
      kern<<<(unsigned)fblocks, kFixThreads, fsmem, st>>>(rb64, sc64, q, n, layout, mask, amb_mask,
                                                          n_words, nullptr, 0.0);
    }
    count_launch();
  });
  return cudaGetLastError();
}

}  // namespace acro

// ---------------------------------------------------------------------------
// measurement: how far are the filter's FP32 decision quantities from the FP64 ones?
// ---------------------------------------------------------------------------
namespace acro {

// the 15 signed separations of FCL's test (face axes of box 2, of box 1, nine edge axes)
template <typename real>
__device__ __forceinline__ void sat_axes(const real* A, const real* R1, const real* c1,
                                         const real* R2, const real* T2, const real* B,
                                         real* s_out) {
  const real p0 = T2[0] - c1[0], p1 = T2[1] - c1[1], p2 = T2[2] - c1[2];
  real R[9], Q[9], pp[3];
  for (int i = 0; i < 3; ++i) {
    pp[i] = fma_t(R1[6 + i], p2, fma_t(R1[3 + i], p1, R1[i] * p0));
    for (int j = 0; j < 3; ++j) {
      R[3 * i + j] = fma_t(R1[6 + i], R2[6 + j], fma_t(R1[3 + i], R2[3 + j], R1[i] * R2[j]));
      Q[3 * i + j] = abs_t(R[3 * i + j]);
    }
  }
  for (int j = 0; j < 3; ++j) {
    const real t = fma_t(R2[6 + j], p2, fma_t(R2[3 + j], p1, R2[j] * p0));
    s_out[j] = abs_t(t) - (fma_t(Q[6 + j], A[2], fma_t(Q[3 + j], A[1], Q[j] * A[0])) + B[j]);
  }
  for (int i = 0; i < 3; ++i)
    s_out[3 + i] = abs_t(pp[i]) -
                   (fma_t(Q[3 * i + 2], B[2], fma_t(Q[3 * i + 1], B[1], Q[3 * i] * B[0])) + A[i]);
  for (int k = 0; k < 9; ++k) Q[k] += real(1.0e-6);
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const real t = fma_t(pp[i2], R[3 * i1 + j], -(pp[i1] * R[3 * i2 + j]));
      s_out[6 + 3 * i + j] =
          abs_t(t) - fma_t(A[i1], Q[3 * i2 + j],
                           fma_t(A[i2], Q[3 * i1 + j],
                                 fma_t(B[j1], Q[3 * i + j2], B[j2] * Q[3 * i + j1])));
    }
  }
}

// out[0] = max over all (configuration, robot box, scene box, axis) of |s_fp32 - s_fp64|,
// out[1] = max |box centre fp32 - fp64| (inf norm); the FP32 side uses exactly the filter's
// arithmetic (sincos_filter + tf_mul_dh<float>), the FP64 side the exact kernels' chain_step.
template <int NDOF>
__global__ void __launch_bounds__(128)
    filter_error_kernel(const __grid_constant__ RobotDev<float> rb32,
                        const __grid_constant__ RobotDev<double> rb64, const SceneView<float> sc32,
                        const SceneView<double> sc64, const double* __restrict__ q, int64_t n,
                        double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
  double worst_s = 0.0, worst_p = 0.0;
  if (idx < n) {
    Tf<float> Tf32;
    Tf<double> Tf64;
    tf_load<float>(Tf32, rb32.base);
    tf_load<double>(Tf64, rb64.base);
    for (int slot = 0; slot <= NDOF + 1; ++slot) {
      if (slot >= 1 && slot <= NDOF) {
        const int i = slot - 1;
        const double qi = q[idx * NDOF + i];
        const bool prismatic = (rb32.prismatic_mask >> i) & 1;
        float s, c;
        sincos_filter(prismatic ? (double)rb32.theta[i] : qi, s, c);
        tf_mul_dh<float>(Tf32, c, s, rb32.ca[i], rb32.sa[i], rb32.a[i],
                         prismatic ? (float)qi : rb32.d[i]);
        chain_step<double>(Tf64, rb64, i, qi);
      }
      for (int b = rb32.slot_begin[slot]; b < rb32.slot_begin[slot + 1]; ++b) {
        Tf<float> W32;
        Tf<double> W64;
        tf_mul<float>(W32, Tf32, rb32.boxes[b].off);
        tf_mul<double>(W64, Tf64, rb64.boxes[b].off);
        for (int k = 0; k < 3; ++k)
          worst_p = fmax(worst_p, fabs((double)W32.p[k] - W64.p[k]));
        for (int j = 0; j < sc64.n; ++j) {
          const float* f = sc32.boxes + (size_t)j * kSceneStride;
          const double* d = sc64.boxes + (size_t)j * kSceneStride;
          float s32[15];
          double s64[15];
          sat_axes<float>(rb32.boxes[b].half, W32.r, W32.p, f, f + 9, f + 12, s32);
          sat_axes<double>(rb64.boxes[b].half, W64.r, W64.p, d, d + 9, d + 12, s64);
          for (int k = 0; k < 15; ++k) worst_s = fmax(worst_s, fabs((double)s32[k] - s64[k]));
        }
      }
    }
  }
  // non-negative doubles order like their bit patterns
  atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(worst_s));
  atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(worst_p));
}

cudaError_t launch_filter_error(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                SceneView<float> sc32, SceneView<double> sc64, const double* q,
                                int64_t n, double* out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess || n <= 0) return e;
  ACRO_DISPATCH_NDOF(rb32.ndof, {
    filter_error_kernel<NDOF><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rb32, rb64, sc32, sc64, q,
                                                                           n, out);
  });
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
