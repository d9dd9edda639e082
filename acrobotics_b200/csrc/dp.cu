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
                                    int64_t* __restrict__ path_rows, double* __restrict__ length) {
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
    *length = best;
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
  const int64_t n0 = offsets_host[1] - offsets_host[0];
  dp_init_kernel<<<(unsigned)((n0 + 255) / 256), 256, 0, st>>>(state_cost, n0, value, pred);
  count_launch();
  for (int64_t l = 1; l < n_layers; ++l) {
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
  dp_backtrack_kernel<<<1, 32, 0, st>>>(value, pred, offsets_dev, n_layers, path_rows, length);
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
