// path.cu - K6: discretised-path collision sweep.
//
// Replaces reference Robot.is_path_in_collision_discrete
// (src/acrobotics/robot.py:275-283) on top of Robot._linear_interpolation_path
// (robot.py:229-237) for a batch of joint-space segments: one warp per segment,
// the lanes stride over the interpolants, each interpolant runs the same fused
// FK -> collision core as K2, the warp votes.
#include "acro_collide.cuh"
#include "acro_launch.h"

namespace acro {

template <int NDOF>
__global__ void __launch_bounds__(kTile)
    path_cc_kernel(const __grid_constant__ RobotDev<double> rb, const SceneView<double> sc,
                   const double* __restrict__ q0, const double* __restrict__ q1, int64_t n_seg,
                   double max_q_step, uint32_t* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_scene = reinterpret_cast<double*>(smem_raw);
  for (int k = threadIdx.x; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t seg = (int64_t)blockIdx.x * (kTile / 32) + (threadIdx.x >> 5);
  if (seg >= n_seg) return;  // warp-uniform

  double a[NDOF], b[NDOF];
  double sq = 0.0;
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    a[i] = q0[seg * NDOF + i];
    b[i] = q1[seg * NDOF + i];
    const double d = b[i] - a[i];
    sq = fma(d, d, sq);
  }
  // robot.py:233-237: num_steps = ceil(|dq| / max_q_step); S = linspace(0, 1, num_steps)
  const double steps_f = ceil(sqrt(sq) / max_q_step);
  const int n_steps = steps_f > 2.0e9 ? 2000000000 : (int)steps_f;
  const double step = n_steps > 1 ? 1.0 / (double)(n_steps - 1) : 0.0;
  const unsigned full = 0xffffffffu;
  bool any = false;
  for (int k0 = 0; k0 < n_steps && !any; k0 += 32) {
    const int k = k0 + lane;
    const bool act = k < n_steps;
    // numpy.linspace: k * step, with the end point forced to exactly 1
    const double s = (k == n_steps - 1 && n_steps > 1) ? 1.0 : (double)k * step;
    double qv[NDOF];
#pragma unroll
    for (int i = 0; i < NDOF; ++i) qv[i] = (1.0 - s) * a[i] + s * b[i];
    Verdict<double, false> v;
    config_collision<double, NDOF, false>(rb, qv, s_scene, sc.n, 0.0, act, v);
    any = __any_sync(full, act && v.hit);
  }
  if (lane == 0 && any) atomicOr(&mask[seg >> 5], 1u << (seg & 31));
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
                                                               mask);
  });
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
