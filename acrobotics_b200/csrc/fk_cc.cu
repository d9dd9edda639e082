// fk_cc.cu - K2: fused forward kinematics -> collision check.
//
// Replaces reference Robot.is_in_collision (src/acrobotics/robot.py:244-273),
// Robot._check_self_collision (robot.py:206-221), Scene.is_in_collision
// (src/acrobotics/geometry.py:51-66) and the python-fcl Box-Box call
// (src/acrobotics/shapes.py:50-58) for a batch of configurations.
//
// One thread per configuration; scene boxes in shared memory; the verdicts of a
// warp are packed with __ballot_sync into one uint32 of the output bit mask.
// Compute bound (FP64 issue rate): 48 B in, 1 bit out per configuration.
#include "acro_collide.cuh"
#include "acro_launch.h"

namespace acro {

template <typename real, int NDOF, bool MARGIN>
__global__ void __launch_bounds__(kTile)
    fk_cc_kernel(const __grid_constant__ RobotDev<real> rb, const SceneView<real> sc,
                 const real* __restrict__ q, int64_t n, int layout, uint32_t* __restrict__ mask,
                 uint32_t* __restrict__ near_mask, real margin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* s_scene = reinterpret_cast<real*>(smem_raw);
  real* s_q = s_scene + sc.n * kSceneStride;
  for (int k = threadIdx.x; k < sc.n * kSceneStride; k += blockDim.x) s_scene[k] = sc.boxes[k];

  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t idx = row0 + threadIdx.x;
  real qv[NDOF];
  load_q<real, NDOF>(qv, s_q, q, row0, n, layout);
  __syncthreads();  // scene tile visible (load_q only syncs on the AoS path)

  Verdict<real, MARGIN> v;
  config_collision<real, NDOF, MARGIN>(rb, qv, s_scene, sc.n, margin, idx < n, v);

  const unsigned full = 0xffffffffu;
  const unsigned word = __ballot_sync(full, idx < n && v.hit);
  unsigned near_word = 0;
  if (MARGIN) {
    const bool near = idx < n && (v.gmin <= margin) && (v.gmin >= -margin);
    near_word = __ballot_sync(full, near);
  }
  if ((threadIdx.x & 31) == 0 && row0 + (threadIdx.x & ~31) < n) {
    mask[idx >> 5] = word;
    if (MARGIN) near_mask[idx >> 5] = near_word;
  }
}

// n independent box pairs: the fcl.collide seam itself (shapes.py:50-58)
__global__ void __launch_bounds__(kTile)
    collide_boxes_kernel(const AcroBox* __restrict__ a, const AcroBox* __restrict__ b, int64_t n,
                         uint32_t* __restrict__ mask) {
  const int64_t idx = (int64_t)blockIdx.x * kTile + threadIdx.x;
  bool hit = false;
  if (idx < n) {
    const AcroBox& A = a[idx];
    const AcroBox& B = b[idx];
    double R1[9], c1[3], ob[16];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        R1[3 * r + c] = A.tf[4 * r + c];
        ob[kOB_R + 3 * r + c] = B.tf[4 * r + c];
      }
      c1[r] = A.tf[4 * r + 3];
      ob[kOB_T + r] = B.tf[4 * r + 3];
      ob[kOB_H + r] = B.half[r];
    }
    double g;
    hit = !sat_separated<double>(A.half[0], A.half[1], A.half[2], R1, c1, ob, 0.0, g);
  }
  const unsigned word = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && idx < n) mask[idx >> 5] = word;
}

// widen joint values (exact): the FP32-input entry point feeds the FP64-exact K2 path
__global__ void __launch_bounds__(256)
    widen_kernel(const float* __restrict__ in, double* __restrict__ out, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < count) out[i] = (double)in[i];
}

cudaError_t launch_widen(const float* in, double* out, int64_t count, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  widen_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(in, out, count);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_collide_boxes(const AcroBox* a, const AcroBox* b, int64_t n, uint32_t* mask,
                                 cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  collide_boxes_kernel<<<(unsigned)tiles, kTile, 0, st>>>(a, b, n, mask);
  count_launch();
  return cudaGetLastError();
}

template <typename real>
cudaError_t launch_fk_cc(const RobotDev<real>& rb, SceneView<real> sc, const real* q, int64_t n,
                         int layout, uint32_t* mask, uint32_t* near_mask, real margin,
                         cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  ACRO_DISPATCH_NDOF(rb.ndof, {
    const size_t smem = sizeof(real) * (sc.n * kSceneStride + kTile * NDOF);
    if (near_mask)
      fk_cc_kernel<real, NDOF, true>
          <<<(unsigned)tiles, kTile, smem, st>>>(rb, sc, q, n, layout, mask, near_mask, margin);
    else
      fk_cc_kernel<real, NDOF, false>
          <<<(unsigned)tiles, kTile, smem, st>>>(rb, sc, q, n, layout, mask, nullptr, real(0));
  });
  count_launch();
  return cudaGetLastError();
}

template cudaError_t launch_fk_cc<double>(const RobotDev<double>&, SceneView<double>,
                                          const double*, int64_t, int, uint32_t*, uint32_t*,
                                          double, cudaStream_t);
template cudaError_t launch_fk_cc<float>(const RobotDev<float>&, SceneView<float>, const float*,
                                         int64_t, int, uint32_t*, uint32_t*, float, cudaStream_t);

}  // namespace acro
