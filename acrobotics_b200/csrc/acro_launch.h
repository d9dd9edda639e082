// acro_launch.h - internal C++ launch interface between the kernel translation
// units and the C ABI (abi.cu).  Not part of the public interface.
#pragma once
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

// production K2 (state machine + refill + FP32 cull); FP64, no near-margin output
cudaError_t launch_fk_cc_sm(const RobotDev<double>& rb, SceneView<double> sc, const double* q,
                            int64_t n, int layout, uint32_t* mask, cudaStream_t st);

// production K2 (link-by-link wavefronts over a shared-memory work list + FP32 cull)
bool wave_supports(const RobotDev<double>& rb);
cudaError_t launch_fk_cc_wave(const RobotDev<double>& rb, const CullTable& ct,
                              SceneView<double> sc, const double* q, int64_t n, int layout,
                              uint32_t* mask, cudaStream_t st);

// production K2 (FP32 filter + candidate grid + exact FP64 fix-up pass); fk_cc_filt.cu
cudaError_t launch_build_grid(SceneView<double> sc, void* cells, int mask_bits, int dim,
                              int n_classes, const double* origin, double cell,
                              const double* class_radius_dev, double pad, cudaStream_t st);
cudaError_t launch_fk_cc_filter(const RobotDev<float>& rb32, const RobotDev<double>& rb64,
                                const GridView& grid, int mask_bits, SceneView<float> sc32,
                                SceneView<double> sc64, const double* q, int64_t n, int layout,
                                uint32_t* mask, uint32_t* amb_mask, cudaStream_t st);

cudaError_t launch_collide_boxes(const AcroBox* a, const AcroBox* b, int64_t n, uint32_t* mask,
                                 cudaStream_t st);

cudaError_t launch_jac(const RobotDev<double>& rb, const double* q, int64_t n, int layout,
                       double* J, bool rpy, cudaStream_t st);

cudaError_t launch_ik_kuka(const RobotDev<double>& rb, const double* T, int64_t n, double* q_out,
                           uint8_t* valid, cudaStream_t st);
cudaError_t launch_ik_kuka_cc(const RobotDev<double>& rb, SceneView<double> sc, const double* T,
                              int64_t n, double* q_out, uint8_t* valid, uint8_t* free_mask,
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

cudaError_t measure_fma_peak(bool fp64, int iters, double* tflops, double* ms);

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

}  // namespace acro
