// path.cu - K6: discretised-path collision sweep.
//
// Replaces reference Robot.is_path_in_collision_discrete
// (src/acrobotics/robot.py:275-283) on top of Robot._linear_interpolation_path
// (robot.py:229-237) for a batch of joint-space segments.
//
// Two kernels:
//  * path_cc_kernel: one warp per segment, the lanes stride over the interpolants, each
//    interpolant runs the exact FP64 FK -> collision core of K2' (small batches, and the exact
//    pass of the filtered variant);
//  * path_cc_filter_kernel: the FP32 filter of the production K2 kernel (acro_filter.cuh)
//    driven by lanes that each own one segment at a time: every round evaluates the next
//    interpolant of 32 segments, a lane whose segment is decided (first sure collision, or
//    all interpolants sure-free) immediately takes the next segment, so short segments (the
//    planner's edges) and long ones keep the warps equally full.  Segments with an ambiguous
//    interpolant and no sure collision are recomputed by path_cc_kernel in FP64.
#include <algorithm>

#include "acro_filter.cuh"
#include "acro_launch.h"

namespace acro {

// robot.py:233-237: num_steps = ceil(|dq| / max_q_step); S = linspace(0, 1, num_steps)
__device__ __forceinline__ int path_steps(double sq, double max_q_step, double& step) {
  const double steps_f = ceil(sqrt(sq) / max_q_step);
  const int n_steps = steps_f > 2.0e9 ? 2000000000 : (int)steps_f;
  step = n_steps > 1 ? 1.0 / (double)(n_steps - 1) : 0.0;
  return n_steps;
}

// numpy.linspace: k * step, with the end point forced to exactly 1
__device__ __forceinline__ double path_param(int k, int n_steps, double step) {
  return (k == n_steps - 1 && n_steps > 1) ? 1.0 : (double)k * step;
}

// one segment, one warp, exact FP64
template <int NDOF>
__device__ __forceinline__ bool segment_collides_f64(const RobotDev<double>& rb,
                                                     const double* s_scene, int n_scene,
                                                     const double* __restrict__ q0,
                                                     const double* __restrict__ q1, int64_t seg,
                                                     double max_q_step, int lane) {
  double a[NDOF], b[NDOF];
  double sq = 0.0;
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    a[i] = q0[seg * NDOF + i];
    b[i] = q1[seg * NDOF + i];
    const double d = b[i] - a[i];
    sq = fma(d, d, sq);
  }
  double step;
  const int n_steps = path_steps(sq, max_q_step, step);
  bool any = false;
  for (int k0 = 0; k0 < n_steps && !any; k0 += 32) {
    const int k = k0 + lane;
    const bool act = k < n_steps;
    const double s = path_param(k, n_steps, step);
    double qv[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) qv[i] = (1.0 - s) * a[i] + s * b[i];
    Verdict<double, false> v;
    config_collision<double, NDOF, false>(rb, qv, s_scene, n_scene, 0.0, act, v);
    any = __any_sync(0xffffffffu, act && v.hit);
  }
  return any;
}

// only == nullptr: every segment, one warp each.  Otherwise: the segments whose bit is set in
// `only` (the ambiguous ones of the filter pass); a CTA compacts the set bits of 128 words and
// its warps walk over the list.
template <int NDOF>
__global__ void __launch_bounds__(kTile)
    path_cc_kernel(const __grid_constant__ RobotDev<double> rb, const SceneView<double> sc,
                   const double* __restrict__ q0, const double* __restrict__ q1, int64_t n_seg,
                   double max_q_step, uint32_t* __restrict__ mask,
                   const uint32_t* __restrict__ only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_scene = reinterpret_cast<double*>(smem_raw);
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_scene + sc.n * kSceneStride);  // [128*32]
  __shared__ int s_warp_tot[kTile / 32];
  __shared__ int s_total;
  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = t >> 5;

  if (only == nullptr) {
    for (int k = t; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];
    __syncthreads();
    const int64_t seg = (int64_t)blockIdx.x * (kTile / 32) + warp;
    if (seg >= n_seg) return;  // warp-uniform
    const bool any =
        segment_collides_f64<NDOF>(rb, s_scene, sc.n, q0, q1, seg, max_q_step, lane);
    if (lane == 0 && any) atomicOr(&mask[seg >> 5], 1u << (seg & 31));
    return;
  }

  const int64_t n_words = (n_seg + 31) / 32;
  const int64_t w = (int64_t)blockIdx.x * kTile + t;
  unsigned bits = w < n_words ? only[w] : 0u;
  const int c = __popc(bits);
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp_tot[warp] = incl;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int k = 0; k < kTile / 32; ++k) {
      const int v = s_warp_tot[k];
      s_warp_tot[k] = run;
      run += v;
    }
    s_total = run;
  }
  __syncthreads();
  const int total = s_total;
  if (total == 0) return;  // CTA-uniform
  int pos = s_warp_tot[warp] + incl - c;
  while (bits) {
    const int b = __ffs((int)bits) - 1;
    bits &= bits - 1;
    s_list[pos++] = (uint16_t)(t * 32 + b);
  }
  for (int k = t; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];
  __syncthreads();
  const int64_t seg0 = (int64_t)blockIdx.x * kTile * 32;
  for (int k = warp; k < total; k += kTile / 32) {
    const int64_t seg = seg0 + s_list[k];
    const bool any =
        segment_collides_f64<NDOF>(rb, s_scene, sc.n, q0, q1, seg, max_q_step, lane);
    if (lane == 0 && any) atomicOr(&mask[seg >> 5], 1u << (seg & 31));
  }
}

// ---------------------------------------------------------------------------
// filtered sweep
// ---------------------------------------------------------------------------
constexpr int kSegBlock = 128;  // segments a warp claims at a time

template <int NDOF, typename MaskT>
__global__ void __launch_bounds__(kFThreads, ACRO_FILTER_MIN_CTAS)
    path_cc_filter_kernel(const __grid_constant__ RobotDev<float> rb, const GridView grid,
                          const SceneView<float> sc, const int n_saved,
                          const double* __restrict__ q0, const double* __restrict__ q1,
                          const int64_t n_seg, const double max_q_step,
                          uint32_t* __restrict__ mask, uint32_t* __restrict__ amb_mask,
                          unsigned long long* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [segment end points: 4 warps x 2*NDOF x 32 doubles][filter context]
  const int t = threadIdx.x;
  const unsigned lane = t & 31;
  const int warp = t >> 5;
  double* s_ab = reinterpret_cast<double*>(smem_raw) + warp * 2 * NDOF * 32;
  unsigned char* ctx_base = smem_raw + sizeof(double) * (kFThreads / 32) * 2 * NDOF * 32;
  filter_ctx_stage(ctx_base, kFThreads / 32, rb, sc, n_saved);
  const FilterCtx ctx = filter_ctx_carve(ctx_base, kFThreads / 32, warp, sc.n, n_saved);
  __syncthreads();

  int64_t seg = -1;       // this lane's segment (-1: none)
  int k = 0, n_steps = 0;
  double step = 0.0;
  bool seg_amb = false;
  int64_t cur = 0, end = 0;  // the warp's claimed block of segments (warp-uniform)
  bool exhausted = false;

  for (;;) {
    // ---- lanes without a segment take the next ones
    for (;;) {
      const unsigned need = __ballot_sync(kFull, seg < 0);
      if (need == 0u || exhausted) break;
      if (cur >= end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)kSegBlock);
        base = __shfl_sync(kFull, base, 0);
        cur = (int64_t)base;
        end = cur + kSegBlock < n_seg ? cur + kSegBlock : n_seg;
        if (cur >= n_seg) {
          exhausted = true;
          break;
        }
      }
      const int rank = __popc(need & ((1u << lane) - 1u));
      const int64_t avail = end - cur;
      const int want = __popc(need);
      const int take = want < avail ? want : (int)avail;
      if (seg < 0 && rank < take) {
        const int64_t sg = cur + rank;
        double sq = 0.0;
#pragma unroll
        for (int i = 0; i < NDOF; ++i) {
          const double a = q0[sg * NDOF + i], b = q1[sg * NDOF + i];
          s_ab[i * 32 + lane] = a;
          s_ab[(NDOF + i) * 32 + lane] = b;
          const double d = b - a;
          sq = fma(d, d, sq);
        }
        n_steps = path_steps(sq, max_q_step, step);
        k = 0;
        seg_amb = false;
        if (n_steps > 0) seg = sg;  // an empty path (q_start == q_goal) never collides
      }
      cur += take;
    }
    const bool active = seg >= 0;
    if (!__any_sync(kFull, active)) break;

    const double s = path_param(k, n_steps, step);
    bool hit, amb, retired;
    filter_chunk<NDOF, MaskT, true>(
        rb, grid, ctx, lane, active,
        [&](int i) { return (1.0 - s) * s_ab[i * 32 + lane] + s * s_ab[(NDOF + i) * 32 + lane]; },
        hit, amb, retired);
    if (active) {
      if (hit) {
        atomicOr(&mask[seg >> 5], 1u << (seg & 31));
        seg = -1;
      } else {
        seg_amb |= amb;
        if (++k == n_steps) {
          if (seg_amb) atomicOr(&amb_mask[seg >> 5], 1u << (seg & 31));
          seg = -1;
        }
      }
    }
  }
}

cudaError_t launch_path_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* q0,
                           const double* q1, int64_t n_seg, double max_q_step, uint32_t* mask,
                           cudaStream_t st) {
  if (n_seg <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(mask, 0, sizeof(uint32_t) * ((n_seg + 31) / 32), st);
  if (e != cudaSuccess) return e;
  const int per_cta = kTile / 32;
  const int64_t ctas = (n_seg + per_cta - 1) / per_cta;
  if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
  ACRO_DISPATCH_NDOF(rb.ndof, {
    const size_t smem = sizeof(double) * (sc.n * kSceneStride);
    path_cc_kernel<NDOF><<<(unsigned)ctas, kTile, smem, st>>>(rb, sc, q0, q1, n_seg, max_q_step,
                                                               mask, nullptr);
  });
  count_launch();
  return cudaGetLastError();
}

template <typename K>
static int path_persistent_blocks(K kern, size_t smem) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFThreads, smem);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

// scratch: ceil(n_seg/32) words (ambiguity mask) followed by one 8-byte counter, zeroed here
cudaError_t launch_path_cc_filter(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                  const GridView& grid, int mask_bits, SceneView<float> sc32,
                                  SceneView<double> sc64, const double* q0, const double* q1,
                                  int64_t n_seg, double max_q_step, uint32_t* mask,
                                  void* scratch, cudaStream_t st) {
  if (n_seg <= 0) return cudaSuccess;
  const int64_t n_words = (n_seg + 31) / 32;
  const size_t amb_bytes = ((sizeof(uint32_t) * n_words + 7) / 8) * 8;
  uint32_t* amb = reinterpret_cast<uint32_t*>(scratch);
  unsigned long long* counter =
      reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(scratch) + amb_bytes);
  cudaError_t e = cudaMemsetAsync(mask, 0, sizeof(uint32_t) * n_words, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(scratch, 0, amb_bytes + 8, st);
  if (e != cudaSuccess) return e;
  int n_saved = 0;
  for (int b = 0; b < rb32.n_boxes; ++b)
    if (rb32.boxes[b].save_slot + 1 > n_saved) n_saved = rb32.boxes[b].save_slot + 1;
  if (!rb32.check_self) n_saved = 0;
  ACRO_DISPATCH_NDOF(rb32.ndof, {
    const size_t smem = sizeof(double) * (kFThreads / 32) * 2 * NDOF * 32 +
                        filter_ctx_bytes(kFThreads / 32, sc32.n, n_saved);
    const int64_t want = (n_seg + kFThreads - 1) / kFThreads;
    if (mask_bits == 32) {
      auto kern = path_cc_filter_kernel<NDOF, uint32_t>;
      if (smem > 48 * 1024)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      const int64_t blocks = std::min<int64_t>(want, path_persistent_blocks(kern, smem));
      kern<<<(unsigned)blocks, kFThreads, smem, st>>>(rb32, grid, sc32, n_saved, q0, q1, n_seg,
                                                      max_q_step, mask, amb, counter);
    } else {
      auto kern = path_cc_filter_kernel<NDOF, uint64_t>;
      if (smem > 48 * 1024)
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      const int64_t blocks = std::min<int64_t>(want, path_persistent_blocks(kern, smem));
      kern<<<(unsigned)blocks, kFThreads, smem, st>>>(rb32, grid, sc32, n_saved, q0, q1, n_seg,
                                                      max_q_step, mask, amb, counter);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t fsmem = sizeof(double) * (sc64.n * kSceneStride) + sizeof(uint16_t) * kTile * 32;
    const int64_t fblocks = (n_words + kTile - 1) / kTile;
    path_cc_kernel<NDOF><<<(unsigned)fblocks, kTile, fsmem, st>>>(rb64, sc64, q0, q1, n_seg,
                                                                  max_q_step, mask, amb);
    count_launch();
  });
  return cudaGetLastError();
}

}  // namespace acro
