// planner.cu - the sampling-based planner's inner loop as a batched pipeline.
//
// Reference: PathPt.to_joint_solutions (src/acrobotics/path/path_pt_base.py:87-111):
// for every path point, the IK solutions of all its sampled poses (_calc_ik :76-85, in
// sample order, each pose's solutions in Kuka.ik's order) filtered by
// `not robot.is_in_collision(q, scene)` (:107-109) - one numpy array of joint vectors per
// point (planning/sampling_based.py:94-106 collects them into the graph's layers).
//
// Here: T holds n_points * samples_per_point poses; the IK kernel and the filtered K2 kernel
// produce the (N,8,6) solutions and the per-pose "collision free" branch bits
// (acro_ik_kuka_cc_f64); this file compacts the surviving joint vectors into CSR form:
// offsets[p] .. offsets[p+1] are the rows of q_free that belong to path point p, in exactly
// the reference's order.
#include <cub/device/device_scan.cuh>

#include "acro_launch.h"

namespace acro {

// one warp per path point: number of surviving solutions
__global__ void __launch_bounds__(128)
    csr_count_kernel(const uint8_t* __restrict__ free_mask, int64_t n_points, int samples,
                     int64_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (p >= n_points) return;
  int c = 0;
  for (int s = lane; s < samples; s += 32) c += __popc((unsigned)free_mask[p * samples + s]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) counts[p] = c;
}

// one warp per path point: copy the surviving rows behind offsets[p]
__global__ void __launch_bounds__(128)
    csr_scatter_kernel(const double* __restrict__ q_all, const uint8_t* __restrict__ free_mask,
                       int64_t n_points, int samples, const int64_t* __restrict__ offsets,
                       double* __restrict__ q_free, int64_t capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (p >= n_points) return;
  int64_t base = offsets[p];
  for (int s0 = 0; s0 < samples; s0 += 32) {
    const int s = s0 + lane;
    const unsigned bits = s < samples ? (unsigned)free_mask[p * samples + s] : 0u;
    const int c = __popc(bits);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int64_t row = base + incl - c;
    unsigned b = bits;
    while (b) {
      const int sl = __ffs((int)b) - 1;
      b &= b - 1;
      if (row < capacity) {
        const double2* src = reinterpret_cast<const double2*>(q_all + ((p * samples + s) * 8 + sl) * 6);
        double2* dst = reinterpret_cast<double2*>(q_free + row * 6);
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
      }
      ++row;
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

__global__ void csr_total_kernel(const int64_t* counts, int64_t* offsets, int64_t n_points) {
  // offsets[0..n) hold the exclusive scan; close the CSR with the grand total
  if (threadIdx.x == 0 && blockIdx.x == 0)
    offsets[n_points] = n_points > 0 ? offsets[n_points - 1] + counts[n_points - 1] : 0;
}

size_t csr_temp_bytes(int64_t n_points) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int64_t*)nullptr, (int64_t*)nullptr,
                                (int)n_points);
  return bytes;
}

// scratch: n_points int64 counts followed by the scan's temporary storage
cudaError_t launch_csr_compact(const double* q_all, const uint8_t* free_mask, int64_t n_points,
                               int samples, int64_t* offsets, double* q_free, int64_t capacity,
                               void* scratch, size_t temp_bytes, cudaStream_t st) {
  if (n_points <= 0) {
    return cudaMemsetAsync(offsets, 0, sizeof(int64_t), st);
  }
  if (n_points > 0x7fffffffLL) return cudaErrorInvalidValue;
  int64_t* counts = reinterpret_cast<int64_t*>(scratch);
  void* temp = reinterpret_cast<unsigned char*>(scratch) + sizeof(int64_t) * n_points;
  const unsigned blocks = (unsigned)((n_points + 3) / 4);
  csr_count_kernel<<<blocks, 128, 0, st>>>(free_mask, n_points, samples, counts);
  count_launch();
  cudaError_t e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, offsets, (int)n_points, st);
  if (e != cudaSuccess) return e;
  csr_total_kernel<<<1, 32, 0, st>>>(counts, offsets, n_points);
  csr_scatter_kernel<<<blocks, 128, 0, st>>>(q_all, free_mask, n_points, samples, offsets, q_free,
                                             capacity);
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
