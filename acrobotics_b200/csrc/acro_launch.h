// acro_launch.h - internal C++ launch interface between the kernel translation
// units and the C ABI (abi.cu).  Not part of the public interface.
#pragma once
#include <cstdint>

#include "acro_device.cuh"

namespace acro {

constexpr int kTile = 128;  // configurations per CTA (= threads per CTA)

void count_launch();

template <typename real>
cudaError_t launch_fk(const RobotDev<real>& rb, const real* q, int64_t n, int layout, real* out,
                      bool all_links, cudaStream_t st);
cudaError_t launch_fk_rpy(const RobotDev<double>& rb, const double* q, int64_t n, int layout,
                          double* out, cudaStream_t st);

template <typename real>
cudaError_t launch_fk_cc(const RobotDev<real>& rb, SceneView<real> sc, const real* q, int64_t n,
                         int layout, uint32_t* mask, uint32_t* near_mask, real margin,
                         cudaStream_t st);

// all-FP64 K2 variant (link-by-link wavefronts over a shared-memory work list + FP32 cull)
bool wave_supports(const RobotDev<double>& rb);
cudaError_t launch_fk_cc_wave(const RobotDev<double>& rb, const CullTable& ct,
                              SceneView<double> sc, const double* q, int64_t n, int layout,
                              uint32_t* mask, cudaStream_t st);

// production K2 (FP32 filter + candidate grid + exact FP64 fix-up pass); fk_cc_filt.cu
cudaError_t launch_build_grid(SceneView<double> sc, void* cells, int mask_bits, int dim,
                              int n_classes, const double* origin, double cell,
                              const double* class_radius_host, double pad, cudaStream_t st);
cudaError_t launch_fk_cc_filter(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                const GridView& grid, int mask_bits, SceneView<float> sc32,
                                SceneView<double> sc64, const double* q, int64_t n, int layout,
                                uint32_t* mask, uint32_t* scratch, uint32_t* near_mask,
                                double margin, int split_slot, cudaStream_t st);
// bytes of `scratch` for a batch of n configurations (ambiguity mask + the chunk counter)
size_t filter_scratch_bytes(int64_t n);

cudaError_t launch_fk_cc_fixup(const RobotDev<double>& rb64, SceneView<double> sc64, const double* q,
                               int64_t n, int layout, uint32_t* mask, const uint32_t* amb_mask,
                               uint32_t* near_mask, double margin, cudaStream_t st);
cudaError_t launch_filter_error(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                SceneView<float> sc32, SceneView<double> sc64, const double* q,
                                int64_t n, double* out, cudaStream_t st);

cudaError_t launch_widen(const float* in, double* out, int64_t count, cudaStream_t st);
cudaError_t launch_collide_boxes(const AcroBox* a, const AcroBox* b, int64_t n, uint32_t* mask,
                                 cudaStream_t st);

cudaError_t launch_jac(const RobotDev<double>& rb, const double* q, int64_t n, int layout,
                       double* J, bool rpy, cudaStream_t st);

cudaError_t launch_ik_kuka(const RobotDev<double>& rb, const double* T, int64_t n, double* q_out,
                           uint8_t* valid, cudaStream_t st);
cudaError_t launch_ik_kuka_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* T,
                              int64_t n, double* q_out, uint8_t* valid, uint8_t* free_mask,
                              cudaStream_t st);
// free[i] = valid[i] & ~hit[i]  (bytes: the eight IK branches of pose i)
cudaError_t launch_free_mask(const uint8_t* valid, const uint8_t* hit, int64_t n, uint8_t* free_mask,
                             cudaStream_t st);
cudaError_t launch_ik_arm2(double a1, double a2, double a3, const double* p, int64_t n,
                           double* q_out, uint8_t* valid, cudaStream_t st);
cudaError_t launch_ik_anthro(double l2, double l3, const double* p, int64_t n, double* q_out,
                             uint8_t* valid, cudaStream_t st);
cudaError_t launch_ik_wrist(const double* R, const double* Rb9, int64_t n, double* q_out,
                            cudaStream_t st);

cudaError_t launch_path_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* q0,
                           const double* q1, int64_t n_seg, double max_q_step, uint32_t* mask,
                           cudaStream_t st);

// K7 (swept.cu): FCL's continuous collision (InterpMotion, naive solver) of (q0, q1) pairs
cudaError_t launch_swept_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* q0,
                            const double* q1, int64_t n, int n_samples, uint32_t* mask,
                            uint32_t* near_mask, double margin, unsigned long long* counter,
                            cudaStream_t st);

cudaError_t launch_swept_boxes(const AcroBox* a0, const AcroBox* a1, const AcroBox* b, int64_t n,
                               int n_samples, uint32_t* mask, cudaStream_t st);

// filtered sweep (FP32 filter + exact FP64 pass over the ambiguous segments); `scratch` holds
// 8 * ceil(ceil(n_seg / 32) / 2) + 8 bytes
cudaError_t launch_path_cc_filter(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                  const GridView& grid, int mask_bits, SceneView<float> sc32,
                                  SceneView<double> sc64, const double* q0, const double* q1,
                                  int64_t n_seg, double max_q_step, uint32_t* mask,
                                  void* scratch, cudaStream_t st);

// CSR compaction of the collision-free IK solutions per path point (planner.cu)
size_t csr_temp_bytes(int64_t n_points);
cudaError_t launch_csr_compact(const double* q_all, const uint8_t* free_mask, int64_t n_points,
                               int samples, int64_t* offsets, double* q_free, int64_t capacity,
                               void* scratch, size_t temp_bytes, cudaStream_t st);

// layered-graph shortest path (dp.cu); cost_type: 0 l1, 1 l2, 2 sum squared, 3 weighted sum squared
cudaError_t launch_shortest_path(const double* q, int ndof, const int64_t* offsets_host,
                                 const int64_t* offsets_dev, int64_t n_layers, int cost_type,
                                 const double* weights_host, const double* state_cost,
                                 double* value, int32_t* pred, int64_t* path_rows, double* length,
                                 cudaStream_t st);

cudaError_t measure_fma_peak(bool fp64, int iters, double* tflops, double* ms);
cudaError_t launch_math_probe(int fn, const double* a, const double* b, double* out0, double* out1,
                              int64_t n, cudaStream_t st);

#define ACRO_DISPATCH_NDOF(ndof, ...)            \
  switch (ndof) {                                \
    case 1: { constexpr int NDOF = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int NDOF = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int NDOF = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int NDOF = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int NDOF = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int NDOF = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int NDOF = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int NDOF = 8; __VA_ARGS__; } break; \
    default: return cudaErrorInvalidValue;       \
  }

// read this thread's joint vector: AoS goes through a coalesced shared-memory
// tile, SoA is read directly (already coalesced)
template <typename real, int NDOF>
__device__ __forceinline__ void load_q(real* qv, real* s_q, const real* __restrict__ q,
                                       int64_t row0, int64_t n, int layout) {
  const int64_t idx = row0 + threadIdx.x;
  if (layout == ACRO_Q_AOS) {
    cta_load_rows<real>(s_q, q, row0, n, kTile, NDOF);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NDOF; ++i) qv[i] = idx < n ? s_q[threadIdx.x * NDOF + i] : real(0);
  } else {
#pragma unroll
    for (int i = 0; i < NDOF; ++i) qv[i] = idx < n ? q[(int64_t)i * n + idx] : real(0);
  }
}

// ---------------------------------------------------------------------------
// direct per-thread row I/O for the streaming (HBM-bound) kernels: every thread moves
// its own row of W reals with 16-byte accesses when the row is 16-byte aligned (even W
// for doubles, W % 4 == 0 for floats, 16-byte aligned base), else element by element.  A warp's
// accesses cover one contiguous span, so every 32-byte sector fetched is fully used.
// ---------------------------------------------------------------------------
template <typename real, int W>
__device__ __forceinline__ void load_row(real* dst, const real* __restrict__ src, bool vec) {
  constexpr int PER = 16 / sizeof(real);
  constexpr int PER8 = 8 / sizeof(real);
  if (W % PER == 0 && vec) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
#pragma unroll
    for (int k = 0; k < W / PER; ++k) {
      const int4 v = __ldg(s4 + k);
      reinterpret_cast<int4*>(dst)[k] = v;
    }
  } else if (W % PER8 == 0 && vec) {
    // rows that are only 8-byte aligned (six floats): 8-byte loads
    const int2* s2 = reinterpret_cast<const int2*>(src);
#pragma unroll
    for (int k = 0; k < W / PER8; ++k) {
      const int2 v = __ldg(s2 + k);
      reinterpret_cast<int2*>(dst)[k] = v;
    }
  } else {
#pragma unroll
    for (int k = 0; k < W; ++k) dst[k] = __ldg(src + k);
  }
}

template <typename real, int W>
__device__ __forceinline__ void store_row(real* __restrict__ dst, const real* src, bool vec) {
  constexpr int PER = 16 / sizeof(real);
  if (W % PER == 0 && vec) {
    int4* d4 = reinterpret_cast<int4*>(dst);
#pragma unroll
    for (int k = 0; k < W / PER; ++k) __stcs(d4 + k, reinterpret_cast<const int4*>(src)[k]);
  } else {
#pragma unroll
    for (int k = 0; k < W; ++k) __stcs(dst + k, src[k]);
  }
}

// Coalesced tile output for the streaming kernels: the CTA's threads park their result rows
// (W reals each) in shared memory and the CTA then writes the tile as one contiguous span with
// 16-byte stores - every 128-byte line leaves the SM whole.  Rows are laid out with an odd stride
// (in 16-byte units), which makes both the per-thread staging stores and the flat copy-out
// reads bank-conflict free.  `ld` = reals between consecutive rows in global memory, `col0` =
// first column (fk_all_links writes one 16-real block per link into 16*ndof-real rows).
template <typename real, int W>
struct RowStage {
  static constexpr int PER = 16 / sizeof(real);
  static constexpr bool kVec = (W % PER) == 0;
  static constexpr int UNITS = kVec ? W / PER : W;
  static constexpr int STRIDE = (UNITS % 2) ? UNITS : UNITS + 1;
  static constexpr size_t kBytes = (size_t)kTile * STRIDE * (kVec ? 16 : sizeof(real));
};

template <typename real, int W>
__device__ __forceinline__ void stage_store_rows(unsigned char* smem, const real* row,
                                                 real* __restrict__ out, int64_t row0, int64_t n,
                                                 int64_t ld, int64_t col0, bool vec) {
  using RS = RowStage<real, W>;
  const int t = threadIdx.x;
  const int64_t left = n - row0;
  const int rows = left < kTile ? (int)left : kTile;
  if (RS::kVec && vec) {
    int4* s = reinterpret_cast<int4*>(smem);
#pragma unroll
    for (int k = 0; k < RS::UNITS; ++k)
      s[t * RS::STRIDE + k] = reinterpret_cast<const int4*>(row)[k];
    __syncthreads();
    const int total = rows * RS::UNITS;
    for (int e = t; e < total; e += kTile) {
      const int r = e / RS::UNITS, c = e - r * RS::UNITS;
      __stcs(reinterpret_cast<int4*>(out + (row0 + r) * ld + col0) + c, s[r * RS::STRIDE + c]);
    }
  } else {
    constexpr int SS = W | 1;
    real* s = reinterpret_cast<real*>(smem);
#pragma unroll
    for (int k = 0; k < W; ++k) s[t * SS + k] = row[k];
    __syncthreads();
    const int total = rows * W;
    for (int e = t; e < total; e += kTile) {
      const int r = e / W, c = e - r * W;
      __stcs(out + (row0 + r) * ld + col0 + c, s[r * SS + c]);
    }
  }
}

template <typename real, int W>
constexpr size_t row_stage_bytes() {
  return RowStage<real, W>::kBytes > (size_t)kTile * (W | 1) * sizeof(real)
             ? RowStage<real, W>::kBytes
             : (size_t)kTile * (W | 1) * sizeof(real);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// joint vector of configuration idx (AoS: one row; SoA: NDOF coalesced loads)
template <typename real, int NDOF>
__device__ __forceinline__ void load_q_direct(real* qv, const real* __restrict__ q, int64_t idx,
                                              int64_t n, int layout, bool vec) {
  if (idx < n) {
    if (layout == ACRO_Q_AOS) {
      alignas(16) real tmp[NDOF % 2 ? NDOF + 1 : NDOF];
      load_row<real, NDOF>(tmp, q + idx * NDOF, vec);
#pragma unroll
      for (int i = 0; i < NDOF; ++i) qv[i] = tmp[i];
    } else {
#pragma unroll
      for (int i = 0; i < NDOF; ++i) qv[i] = __ldg(q + (int64_t)i * n + idx);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NDOF; ++i) qv[i] = real(0);
  }
}

}  // namespace acro
