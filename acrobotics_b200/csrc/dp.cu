// dp.cu - shortest path through the planner's layered graph.
//
// Reference call sites: find_shortest_joint_path (src/acrobotics/planning/sampling_based.py:
// 109-142) -> acrolib.dynamic_programming.shortest_path / shortest_path_with_state_cost with one
// of acrolib.cost_functions.{norm_l1, norm_l2, sum_squared, weighted_sum_squared}
// (sampling_based.py:28-44).  acrolib is an un-vendored dependency (requirements.txt:5): the
// algorithm is restated from its call sites - layer i holds the joint solutions of path point
// i, every node of layer i is connected to every node of layer i+1 with cost c(q_a, q_b), a
// node may carry a state cost, and the result is the cheapest source-to-sink chain, its nodes
// and its length.  Ties go to the lowest node index.  Parity with acrolib is unpinned.
//
// One launch per layer transition (a min-plus product V' = min_j (V_j + c(j, k))): a warp owns
// one target node, its lanes stride over the source nodes staged through shared memory, then
// an arg-min reduction across the warp.  FP64 throughout.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "acro_launch.h"

namespace acro {

constexpr int kDpThreads = 256;   // 8 warps = 8 target nodes per CTA pass
constexpr int kDpTile = 256;      // source nodes staged per pass
constexpr int kDpMaxDof = ACRO_MAX_DOF;

struct DpWeights {
  double w[kDpMaxDof];
};

// components beyond ndof are zero in both operands, so the loop runs over the full width
template <int COST>
__device__ __forceinline__ double dp_edge_cost(const double* a, const double* b, const DpWeights& w,
                                               int /*ndof*/) {
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < kDpMaxDof; ++i) {
    const double d = a[i] - b[i];
    if (COST == 0) acc += fabs(d);               // norm_l1
    if (COST == 1 || COST == 2) acc = fma(d, d, acc);  // norm_l2 / sum_squared
    if (COST == 3) acc = fma(w.w[i] * d, d, acc);      // weighted_sum_squared
  }
  return COST == 1 ? sqrt(acc) : acc;
}

// TPW = target nodes per warp (register blocking: with 4, sources are staged once per 32
// targets; with 1, small layers still spread over enough warps to fill the machine)
template <int COST, int kDpTpw>
__global__ void __launch_bounds__(kDpThreads)
    dp_layer_kernel(const double* __restrict__ q, int ndof, int64_t src0, int n_src, int64_t dst0,
                    int n_dst, const double* __restrict__ state_cost, double* __restrict__ value,
                    int32_t* __restrict__ pred, const DpWeights w) {
  __shared__ double s_q[kDpTile * kDpMaxDof];
  __shared__ double s_v[kDpTile];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int per_cta = (kDpThreads / 32) * kDpTpw;
  for (int k0 = blockIdx.x * per_cta; k0 < n_dst; k0 += gridDim.x * per_cta) {
    const int kb = k0 + warp * kDpTpw;
    double tq[kDpTpw][kDpMaxDof];
    double best[kDpTpw];
    int best_j[kDpTpw];
#pragma unroll
    for (int r = 0; r < kDpTpw; ++r) {
      best[r] = 1.0 / 0.0;
      best_j[r] = -1;
#pragma unroll
      for (int i = 0; i < kDpMaxDof; ++i)
        tq[r][i] = (kb + r < n_dst && i < ndof) ? q[(dst0 + kb + r) * ndof + i] : 0.0;
    }
    for (int j0 = 0; j0 < n_src; j0 += kDpTile) {
      const int m = n_src - j0 < kDpTile ? n_src - j0 : kDpTile;
      __syncthreads();
      for (int e = threadIdx.x; e < m * ndof; e += kDpThreads) s_q[e] = q[(src0 + j0) * ndof + e];
      for (int e = threadIdx.x; e < m; e += kDpThreads) s_v[e] = value[src0 + j0 + e];
      __syncthreads();
      for (int j = lane; j < m; j += 32) {
        double a[kDpMaxDof];
#pragma unroll
        for (int i = 0; i < kDpMaxDof; ++i) a[i] = i < ndof ? s_q[j * ndof + i] : 0.0;
        const double v = s_v[j];
#pragma unroll
        for (int r = 0; r < kDpTpw; ++r) {
          const double c = v + dp_edge_cost<COST>(a, tq[r], w, ndof);
          if (c < best[r]) {  // strict: the lowest index wins ties within a lane
            best[r] = c;
            best_j[r] = j0 + j;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kDpTpw; ++r) {
      double b = best[r];
      int bj = best_j[r];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {  // arg-min across the warp, lowest index on ties
        const double ob = __shfl_xor_sync(0xffffffffu, b, d);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, d);
        if (ob < b || (ob == b && oj >= 0 && (bj < 0 || oj < bj))) {
          b = ob;
          bj = oj;
        }
      }
      if (lane == 0 && kb + r < n_dst) {
        value[dst0 + kb + r] = b + (state_cost ? state_cost[dst0 + kb + r] : 0.0);
        pred[dst0 + kb + r] = bj;
      }
    }
  }
}

// Small layers (the common case: a few hundred to a few thousand nodes): one CTA per target
// node, one or a few source nodes per thread - the whole transition is one load round trip, a
// handful of FP64 operations and a CTA-wide arg-min, so a layer costs little more than its
// launch.  Lowest index wins ties, like the tiled kernel.
constexpr int kDpCtaThreads = 128;  // 8 CTAs per SM: ~1200 target nodes per wave

template <int COST>
__global__ void __launch_bounds__(kDpCtaThreads, 8)
    dp_layer_cta_kernel(const double* __restrict__ q, int ndof, int64_t src0, int n_src,
                        int64_t dst0, const double* __restrict__ state_cost,
                        double* __restrict__ value, int32_t* __restrict__ pred,
                        const DpWeights w) {
  __shared__ double s_best[kDpCtaThreads / 32];
  __shared__ int s_idx[kDpCtaThreads / 32];
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double tq[kDpMaxDof];
#pragma unroll
  for (int i = 0; i < kDpMaxDof; ++i) tq[i] = i < ndof ? __ldg(q + (dst0 + k) * ndof + i) : 0.0;
  double best = 1.0 / 0.0;
  int best_j = -1;
  // two source nodes per thread in flight
  for (int j0 = threadIdx.x; j0 < n_src; j0 += 2 * kDpCtaThreads) {
    double a[2][kDpMaxDof], v[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = j0 + u * kDpCtaThreads;
      const bool ok = j < n_src;
#pragma unroll
      for (int i = 0; i < kDpMaxDof; ++i)
        a[u][i] = (ok && i < ndof) ? __ldg(q + (src0 + j) * ndof + i) : 0.0;
      v[u] = ok ? value[src0 + j] : 1.0 / 0.0;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const double c = v[u] + dp_edge_cost<COST>(a[u], tq, w, ndof);
      if (c < best) {  // ascending index within a thread: the lowest index wins ties
        best = c;
        best_j = j0 + u * kDpCtaThreads;
      }
    }
  }
  auto better = [](double ob, int oj, double b, int bj) {
    return ob < b || (ob == b && oj >= 0 && (bj < 0 || oj < bj));
  };
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, d);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, d);
    if (better(ob, oj, best, best_j)) {
      best = ob;
      best_j = oj;
    }
  }
  if (lane == 0) {
    s_best[warp] = best;
    s_idx[warp] = best_j;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int v2 = 1; v2 < kDpCtaThreads / 32; ++v2)
      if (better(s_best[v2], s_idx[v2], best, best_j)) {
        best = s_best[v2];
        best_j = s_idx[v2];
      }
    value[dst0 + k] = best + (state_cost ? state_cost[dst0 + k] : 0.0);
    pred[dst0 + k] = best_j;
  }
}

__global__ void dp_init_kernel(const double* __restrict__ state_cost, int64_t n0,
                               double* __restrict__ value, int32_t* __restrict__ pred) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n0) {
    value[i] = state_cost ? state_cost[i] : 0.0;
    pred[i] = -1;
  }
}

// single warp: arg-min over the last layer, then walk the predecessors back to layer 0
__global__ void dp_backtrack_kernel(const double* __restrict__ value,
                                    const int32_t* __restrict__ pred,
                                    const int64_t* __restrict__ offsets, int64_t n_layers,
                                    int64_t* __restrict__ path_rows, double* __restrict__ length,
                                    const unsigned* __restrict__ sync) {
  const int lane = threadIdx.x;
  const int64_t lo = offsets[n_layers - 1], hi = offsets[n_layers];
  double best = 1.0 / 0.0;
  int64_t best_k = -1;
  for (int64_t k = lo + lane; k < hi; k += 32)
    if (value[k] < best) {
      best = value[k];
      best_k = k;
    }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, d);
    const long long ok = __shfl_xor_sync(0xffffffffu, (long long)best_k, d);
    if (ob < best || (ob == best && ok >= 0 && (best_k < 0 || ok < best_k))) {
      best = ob;
      best_k = ok;
    }
  }
  if (lane == 0) {
    // a grid barrier of the persistent kernel gave up: the values are not trustworthy
    *length = (sync && sync[1] != 0u) ? __longlong_as_double(0x7ff8000000000000LL) : best;
    int64_t node = best_k;
    for (int64_t layer = n_layers - 1; layer >= 0; --layer) {
      path_rows[layer] = node;
      if (layer > 0 && node >= 0) {
        const int32_t p = pred[node];
        node = p < 0 ? -1 : offsets[layer - 1] + p;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Whole graph in ONE launch (the common planner case: thousands of layers of a few hundred to
// a couple of thousand nodes).  A layer transition is ~0.5 us of FP64 work for the whole chip,
// but the layers depend on each other, so with one launch per layer the chain is bound by the
// kernel boundary (~9 us per layer measured).  Here one persistent CTA per SM (cooperative
// launch: all CTAs are co-resident) walks the layers; target node k of a layer belongs to warp
// k / G of CTA k % G, so a layer spreads over every SM's FP64 pipe.
//
// Synchronisation is by DATA FLOW, not by a grid barrier (a counter barrier over 148 CTAs was
// measured at ~2.5 us per layer): the value array is pre-filled with a NaN sentinel, a consumer
// spins on the very values it needs (ld.volatile from L2) until they are no longer NaN, and a
// producer simply stores its value (st.cg).  Nothing else produced during the walk is read
// during the walk (the joint rows are input, `pred` is only read by the back-tracking kernel),
// so no fences are needed; a CTA can run at most one layer ahead of the slowest one.  The next
// layer's joint rows (independent of the values) are copied into shared memory with cp.async
// while the current layer is evaluated.  A spin that does not complete within ~2 s sets a
// sticky flag, after which nobody waits any more - the kernel can not hang the device; the
// back-tracking kernel then reports a NaN length.
// ---------------------------------------------------------------------------
#ifndef ACRO_DP_THREADS
#define ACRO_DP_THREADS 256
#endif
constexpr int kDpPThreads = ACRO_DP_THREADS;
constexpr int kDpPWarps = kDpPThreads / 32;
constexpr int kDpPMaxNodes = 1920;  // two padded row buffers of 6-dof layers fit in 227 KB
#ifndef ACRO_DP_UNROLL
#define ACRO_DP_UNROLL 8
#endif
constexpr int kDpPUnroll = ACRO_DP_UNROLL;
#ifndef ACRO_DP_BATCH
#define ACRO_DP_BATCH 6
#endif
constexpr int kDpPBatch = ACRO_DP_BATCH;  // targets scored per pass over the source rows
static_assert(kDpPBatch <= kDpPWarps, "one warp finishes one target of a batch");
#ifndef ACRO_DP_TIMING
#define ACRO_DP_TIMING 0
#endif
#ifndef ACRO_DP_FENCE
#define ACRO_DP_FENCE 0
#endif
#ifndef ACRO_DP_WAR_FENCE
#define ACRO_DP_WAR_FENCE 0
#endif

__device__ __forceinline__ void dp_cp_async_16(void* smem_dst, const void* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void dp_cp_async_8(void* smem_dst, const void* gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}

// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void dp_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(count)
               : "memory");
}
__device__ __forceinline__ void dp_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool dp_mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void dp_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          (unsigned)__cvta_generic_to_shared(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}

__device__ __forceinline__ double dp_ld_volatile(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// value of a source node, waiting for its producer if necessary
__device__ __forceinline__ double dp_wait_value(const double* p, unsigned* flag) {
  double v = dp_ld_volatile(p);
  if (v == v) return v;
  long long t0 = 0;
  for (unsigned spins = 1;; ++spins) {
    v = dp_ld_volatile(p);
    if (v == v) return v;
    if ((spins & 1023u) == 0u) {
      if (t0 == 0) t0 = clock64();
      if (*reinterpret_cast<volatile unsigned*>(flag) != 0u) return v;
      if (clock64() - t0 > 4000000000LL) {  // ~2 s
        atomicExch(flag, 1u);
        return v;
      }
    }
  }
}

// Minimum of (value, index) pairs over the lanes of `mask` (lowest non-negative index among
// equal values, -1 when nobody holds an index): the value through a monotone 64-bit key and two
// 32-bit hardware reductions, the index through a third.  Values are never NaN here.
__device__ __forceinline__ void dp_warp_argmin(unsigned mask, double& b, int& bj) {
  const long long bits = __double_as_longlong(b);
  // IEEE doubles order like sign-magnitude integers: flip all bits of negatives, the sign bit
  // of non-negatives
  const unsigned long long key =
      (unsigned long long)bits ^ ((bits >> 63) | (long long)0x8000000000000000ULL);
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_min_sync(mask, hi);
  const unsigned mlo = __reduce_min_sync(mask, hi == mhi ? lo : 0xffffffffu);
  const bool mine = hi == mhi && lo == mlo;
  const unsigned mj = __reduce_min_sync(mask, mine ? (unsigned)bj : 0xffffffffu);
  const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
  const long long mbits =
      (long long)(mkey ^ (((long long)~mkey >> 63) | (long long)0x8000000000000000ULL));
  b = __longlong_as_double(mbits);
  bj = (int)mj;
}

__device__ __forceinline__ bool dp_better(double ob, int oj, double b, int bj) {
  return ob < b || (ob == b && oj >= 0 && (bj < 0 || oj < bj));
}

// One persistent CTA per SM; in every layer target node k belongs to warp k / G of CTA k % G
// (a layer of ~840 nodes is one target per warp on ~6 warps of every SM).
template <int COST, int NDOF>
__global__ void __launch_bounds__(kDpPThreads, 1)
    dp_persistent_kernel(const double* __restrict__ q, const int64_t* __restrict__ offsets,
                         const int64_t n_layers, const double* __restrict__ state_cost,
                         double* value, int32_t* __restrict__ pred, const DpWeights w,
                         unsigned* sync) {
  extern __shared__ __align__(16) unsigned char dp_smem[];
  // The rows of a layer lie in shared memory as they lie in global memory (pitch NDOF), so a
  // layer arrives with ONE TMA bulk copy issued by one thread (round 2 first version: rows
  // padded to an odd pitch and copied element-wise with 8-byte cp.async - 20 copies per thread
  // and layer, 2 600 cycles of issue on the critical path of every transition).  Lane j reads
  // row j: conflict free with 16-byte reads when NDOF is even (a 48-byte lane stride covers
  // every bank once per quarter warp), with 8-byte reads when NDOF is odd.
  constexpr int kPitch = NDOF;
  constexpr int kRowCap = kDpPMaxNodes * kPitch;     // doubles per row buffer
  double* s_q = reinterpret_cast<double*>(dp_smem);  // [2][kRowCap]
  double* s_v = s_q + 2 * kRowCap;                   // [kDpPMaxNodes]
  __shared__ __align__(8) uint64_t s_bar[2];         // one mbarrier per row buffer
  __shared__ double s_pb[kDpPBatch][kDpPWarps];      // per-warp partial minima of a batch
  __shared__ int s_pj[kDpPBatch][kDpPWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  unsigned* flag = sync + 1;
  if (tid == 0) {
    dp_mbar_init(&s_bar[0], 1);
    dp_mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // the rows of every layer go by bulk copy when source addresses and sizes are multiples of 16
  // bytes: even NDOF and a 16-byte aligned q (one decision for the whole walk - the barrier
  // phases below count one copy per layer)
  const bool bulk = (NDOF % 2 == 0) && (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
  // asynchronous copy of a layer's joint rows into buffer `buf`.  The buffer was last read in
  // the previous transition, which ended with __syncthreads.
  auto prefetch_rows = [&](int64_t layer, int buf) {
    const int64_t r0 = __ldg(offsets + layer);
    const int cnt = (int)(__ldg(offsets + layer + 1) - r0) * NDOF;  // doubles
    const double* src = q + r0 * NDOF;
    double* dst = s_q + buf * kRowCap;
    if (bulk) {
      if (tid == 0) {
#if ACRO_DP_WAR_FENCE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        // (no proxy fence by default: the buffer's last generic-proxy accesses were reads,
        // complete before the __syncthreads that ended the previous transition)
        dp_mbar_expect_tx(&s_bar[buf], (unsigned)cnt * 8u);
        dp_bulk_g2s(dst, src, (unsigned)cnt * 8u, &s_bar[buf]);
      }
    } else {
      for (int e = tid; e < cnt; e += kDpPThreads) dp_cp_async_8(dst + e, src + e);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");  // (empty after a bulk copy)
  };
  // the joint rows are read once, i.e. from DRAM: pull them into L2 two layers ahead
  auto prefetch_l2 = [&](int64_t layer) {
    if (layer >= n_layers) return;
    const int64_t r0 = __ldg(offsets + layer);
    const int64_t bytes = (__ldg(offsets + layer + 1) - r0) * NDOF * (int64_t)sizeof(double);
    const char* src = reinterpret_cast<const char*>(q + r0 * NDOF);
    for (int64_t b = (int64_t)tid * 128; b < bytes; b += (int64_t)kDpPThreads * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(src + b));
  };
  prefetch_l2(1);
  prefetch_l2(2);
  prefetch_rows(0, 0);
  {
    const int64_t n0 = __ldg(offsets + 1) - __ldg(offsets);
    for (int64_t i = (int64_t)blockIdx.x * kDpPThreads + tid; i < n0; i += (int64_t)G * kDpPThreads) {
      pred[i] = -1;
      __stcg(value + i, state_cost ? state_cost[i] : 0.0);
    }
  }
#if ACRO_DP_TIMING
  long long t_pre = 0, t_wait = 0, t_comp = 0, t_tail = 0, tt;
#define ACRO_DP_T(acc) { const long long now = clock64(); acc += now - tt; tt = now; }
  tt = clock64();
#else
#define ACRO_DP_T(acc)
#endif
  for (int64_t l = 1; l < n_layers; ++l) {
    const int64_t s0 = __ldg(offsets + l - 1), d0 = __ldg(offsets + l);
    const int ns = (int)(d0 - s0), nd = (int)(__ldg(offsets + l + 1) - d0);
    const double* sq = s_q + ((l - 1) & 1) * kRowCap;
    // this layer's rows are the source rows of the next transition: start copying them now
    if (l + 1 < n_layers) prefetch_rows(l, l & 1);
    prefetch_l2(l + 2);
    // The CTA owns the targets k = blockIdx.x + G t, t < tpc (6 of a layer of 840 nodes), and
    // evaluates them kDpPBatch at a time: every thread takes a share of the SOURCE rows and
    // scores each of its rows against all targets of the batch, so that a row is read from
    // shared memory once per batch instead of once per target (with one warp per target the
    // transition was bound by 6 x 40 KB of shared-memory reads, and two of the eight warps had
    // nothing to do).  The first batch's target rows (read-only input, the same addresses for
    // every thread) are in flight while the values arrive.
    const int tpc = (nd + G - 1) / G;
    double tq[kDpPBatch][NDOF];
#pragma unroll
    for (int t = 0; t < kDpPBatch; ++t) {
      const int k = blockIdx.x + G * t;
#pragma unroll
      for (int i = 0; i < NDOF; ++i) tq[t][i] = k < nd ? __ldg(q + (d0 + k) * NDOF + i) : 0.0;
    }
    ACRO_DP_T(t_pre)
    {
      // every thread fetches all of its source values at once (one L2 round trip instead of one
      // per value) and re-fetches, again all at once, those that had not been produced yet
      constexpr int kPer = (kDpPMaxNodes + kDpPThreads - 1) / kDpPThreads;
      double v[kPer];
      unsigned pending = 0;
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int j = tid + u * kDpPThreads;
        v[u] = j < ns ? dp_ld_volatile(value + s0 + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kPer; ++u)
        if (!(v[u] == v[u])) pending |= 1u << u;
      long long t0 = 0;
      for (unsigned spins = 1; pending != 0u; ++spins) {
#pragma unroll
        for (int u = 0; u < kPer; ++u)
          if ((pending >> u) & 1u) v[u] = dp_ld_volatile(value + s0 + tid + u * kDpPThreads);
#pragma unroll
        for (int u = 0; u < kPer; ++u)
          if (v[u] == v[u]) pending &= ~(1u << u);
        if ((spins & 1023u) == 0u) {  // give up after ~2 s (sticky flag: nobody waits any more)
          if (t0 == 0) t0 = clock64();
          if (*reinterpret_cast<volatile unsigned*>(flag) != 0u) break;
          if (clock64() - t0 > 4000000000LL) {
            atomicExch(flag, 1u);
            break;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int j = tid + u * kDpPThreads;
        if (j < ns) s_v[j] = v[u];
      }
    }
    if (l + 1 < n_layers)
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    else
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (bulk) {
      // rows of layer l - 1: the ((l - 1) >> 1)-th copy into buffer (l - 1) & 1
      const unsigned parity = (unsigned)((l - 1) >> 1) & 1u;
      while (!dp_mbar_try_wait(&s_bar[(l - 1) & 1], parity)) {
      }
    }
    __syncthreads();
    ACRO_DP_T(t_wait)
    for (int t0 = 0; t0 < tpc; t0 += kDpPBatch) {  // CTA-uniform trip count (one batch as a rule)
      if (t0 > 0) {
#pragma unroll
        for (int t = 0; t < kDpPBatch; ++t) {
          const int k = blockIdx.x + G * (t0 + t);
#pragma unroll
          for (int i = 0; i < NDOF; ++i) tq[t][i] = k < nd ? __ldg(q + (d0 + k) * NDOF + i) : 0.0;
        }
      }
      double best[kDpPBatch];
      int best_j[kDpPBatch];
#pragma unroll
      for (int t = 0; t < kDpPBatch; ++t) {
        best[t] = 1.0 / 0.0;
        best_j[t] = -1;
      }
      // ascending source index within a thread and a strict comparison: the lowest index wins
      // ties, like the per-layer kernels.  The FP64 accumulation of one edge is a dependent
      // chain (its order is part of the result); the kDpPBatch edges of a row are independent.
      for (int j = tid; j < ns; j += kDpPThreads) {
        double sr[NDOF];
        const double* rowp = sq + j * kPitch;
        if (NDOF % 2 == 0) {
#pragma unroll
          for (int i = 0; i < NDOF / 2; ++i) {
            const double2 v2 = reinterpret_cast<const double2*>(rowp)[i];
            sr[2 * i] = v2.x;
            sr[2 * i + 1] = v2.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NDOF; ++i) sr[i] = rowp[i];
        }
        const double sv = s_v[j];
#pragma unroll
        for (int t = 0; t < kDpPBatch; ++t) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < NDOF; ++i) {
            const double d = sr[i] - tq[t][i];
            if (COST == 0) acc += fabs(d);
            if (COST == 1 || COST == 2) acc = fma(d, d, acc);
            if (COST == 3) acc = fma(w.w[i] * d, d, acc);
          }
          if (COST == 1) acc = sqrt(acc);
          const double c = sv + acc;
          if (c < best[t]) {
            best[t] = c;
            best_j[t] = j;
          }
        }
      }
      // minimum over the warp, then over the CTA's warps through shared memory
#pragma unroll
      for (int t = 0; t < kDpPBatch; ++t) dp_warp_argmin(0xffffffffu, best[t], best_j[t]);
      if (lane == 0) {
#pragma unroll
        for (int t = 0; t < kDpPBatch; ++t) {
          s_pb[t][warp] = best[t];
          s_pj[t][warp] = best_j[t];
        }
      }
      __syncthreads();
      if (warp < kDpPBatch) {
        const int k = blockIdx.x + G * (t0 + warp);
        double b = lane < kDpPWarps ? s_pb[warp][lane] : 1.0 / 0.0;
        int bj = lane < kDpPWarps ? s_pj[warp][lane] : -1;
        dp_warp_argmin(0xffffffffu, b, bj);
        if (lane == 0 && k < nd) {
          pred[d0 + k] = bj;
          __stcg(value + d0 + k, b + (state_cost ? state_cost[d0 + k] : 0.0));
#if ACRO_DP_FENCE
          __threadfence();
#endif
        }
      }
      if (t0 + kDpPBatch < tpc) __syncthreads();  // the partials are reused by the next batch
    }
    ACRO_DP_T(t_comp)
    __syncthreads();  // s_v and the row buffer of this transition are free again
    ACRO_DP_T(t_tail)
  }
#if ACRO_DP_TIMING
  if ((blockIdx.x == 0 || blockIdx.x == 100) && (tid == 0 || tid == 224))
    printf("cta %d thr %d cycles/layer: pre %lld wait %lld comp %lld tail %lld\n", blockIdx.x, tid,
           t_pre / n_layers, t_wait / n_layers, t_comp / n_layers, t_tail / n_layers);
#endif
}

// launches the persistent kernel when the graph is eligible (*launched), else leaves the
// per-layer path to the caller
static cudaError_t try_persistent(const double* q, int ndof, const int64_t* offsets_host,
                                  const int64_t* offsets_dev, int64_t n_layers, int cost_type,
                                  const DpWeights& w, const double* state_cost, double* value,
                                  int32_t* pred, unsigned* sync, cudaStream_t st, bool* launched) {
  *launched = false;
  static const bool disabled = std::getenv("ACRO_DP_PERSISTENT") &&
                               std::string(std::getenv("ACRO_DP_PERSISTENT")) == "0";
  if (disabled || n_layers < 32) return cudaSuccess;
  for (int64_t l = 0; l < n_layers; ++l)
    if (offsets_host[l + 1] - offsets_host[l] > kDpPMaxNodes) return cudaSuccess;
  // two row buffers + the source values
  const size_t smem = sizeof(double) * ((size_t)2 * kDpPMaxNodes * ndof + kDpPMaxNodes);
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop || sms < 1 || smem > 227 * 1024) return cudaSuccess;
  const void* fn = nullptr;
#define ACRO_DP_PICK(C)                                                       \
  ACRO_DISPATCH_NDOF(ndof, { fn = (const void*)dp_persistent_kernel<C, NDOF>; })
  switch (cost_type) {
    case 0: ACRO_DP_PICK(0); break;
    case 1: ACRO_DP_PICK(1); break;
    case 2: ACRO_DP_PICK(2); break;
    default: ACRO_DP_PICK(3); break;
  }
#undef ACRO_DP_PICK
  int per_sm = 0;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kDpPThreads, smem);
  if (e != cudaSuccess || per_sm < 1) return e;
  // NaN sentinel in every value: "not produced yet"
  const int64_t total = offsets_host[n_layers];
  e = cudaMemsetAsync(value, 0xFF, sizeof(double) * (size_t)total, st);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&q,     (void*)&offsets_dev, (void*)&n_layers, (void*)&state_cost,
                  (void*)&value, (void*)&pred,        (void*)&w,        (void*)&sync};
  e = cudaLaunchCooperativeKernel(fn, dim3((unsigned)sms), dim3(kDpPThreads), args, smem, st);
  if (e == cudaSuccess) {
    count_launch();
    *launched = true;
  }
  return e;
}

cudaError_t launch_shortest_path(const double* q, int ndof, const int64_t* offsets_host,
                                 const int64_t* offsets_dev, int64_t n_layers, int cost_type,
                                 const double* weights_host, const double* state_cost,
                                 double* value, int32_t* pred, int64_t* path_rows, double* length,
                                 cudaStream_t st) {
  DpWeights w;
  for (int i = 0; i < kDpMaxDof; ++i) w.w[i] = (weights_host && i < ndof) ? weights_host[i] : 1.0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  bool persistent = false;
  unsigned* sync = nullptr;  // [0] barrier counter, [1] sticky "a barrier timed out" flag
  {
    cudaError_t ep = cudaMallocAsync(&sync, 2 * sizeof(unsigned), st);
    if (ep == cudaSuccess) ep = cudaMemsetAsync(sync, 0, 2 * sizeof(unsigned), st);
    if (ep == cudaSuccess)
      ep = try_persistent(q, ndof, offsets_host, offsets_dev, n_layers, cost_type, w, state_cost,
                          value, pred, sync, st, &persistent);
    if (ep != cudaSuccess) {
      if (sync) cudaFreeAsync(sync, st);
      return ep;
    }
  }
  const int64_t n0 = offsets_host[1] - offsets_host[0];
  if (!persistent) {
    dp_init_kernel<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(state_cost, n0, value, pred);
    count_launch();
  }
  for (int64_t l = 1; l < n_layers && !persistent; ++l) {
    const int64_t s0 = offsets_host[l - 1], d0 = offsets_host[l];
    const int64_t ns = d0 - s0, nd = offsets_host[l + 1] - d0;
    if (ns > 0x7fffffff || nd > 0x7fffffff) return cudaErrorInvalidValue;
    const bool small = nd < (int64_t)sms * (kDpThreads / 32) * 4 * 2;
    const int per_cta = (kDpThreads / 32) * (small ? 1 : 4);
    int64_t blocks = (nd + per_cta - 1) / per_cta;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    if (blocks < 1) blocks = 1;
#define ACRO_DP_LAUNCH(C)                                                                         \
  if (ns <= 8192 && nd <= 65535)                                                                  \
    dp_layer_cta_kernel<C><<<(unsigned)nd, kDpCtaThreads, 0, st>>>(q, ndof, s0, (int)ns, d0,        \
                                                                state_cost, value, pred, w);     \
  else if (small)                                                                                 \
    dp_layer_kernel<C, 1><<<(unsigned)blocks, kDpThreads, 0, st>>>(q, ndof, s0, (int)ns, d0,     \
                                                                   (int)nd, state_cost, value,  \
                                                                   pred, w);                     \
  else                                                                                            \
    dp_layer_kernel<C, 4><<<(unsigned)blocks, kDpThreads, 0, st>>>(q, ndof, s0, (int)ns, d0,     \
                                                                   (int)nd, state_cost, value,  \
                                                                   pred, w)
    switch (cost_type) {
      case 0: ACRO_DP_LAUNCH(0); break;
      case 1: ACRO_DP_LAUNCH(1); break;
      case 2: ACRO_DP_LAUNCH(2); break;
      case 3: ACRO_DP_LAUNCH(3); break;
      default: return cudaErrorInvalidValue;
    }
#undef ACRO_DP_LAUNCH
    count_launch();
  }
  dp_backtrack_kernel<<<1, 32, 0, st>>>(value, pred, offsets_dev, n_layers, path_rows, length, sync);
  count_launch();
  cudaError_t e = cudaGetLastError();
  const cudaError_t e2 = cudaFreeAsync(sync, st);
  return e != cudaSuccess ? e : e2;
}

}  // namespace acro
