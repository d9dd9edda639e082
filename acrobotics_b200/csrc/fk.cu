// fk.cu - K1: batched forward kinematics of serial DH chains.
//
// Replaces reference Robot.fk (src/acrobotics/robot.py:59-69), fk_all_links
// (robot.py:82-91) and fk_rpy (robot.py:71-80) built on
// LinkKinematics.get_link_relative_transform (src/acrobotics/link.py:35-58).
//
// One thread per configuration, transform chained in registers.  HBM bound
// (48 B in, 128 B out per configuration for a 6-dof fk): both the joint rows
// and the 4x4 results cross global memory as flat coalesced blocks staged
// through shared memory (a 16-double row per thread would otherwise be a
// 128-byte-strided store).
#include "acro_launch.h"

namespace acro {

constexpr int kOutPad = 17;  // 16 reals + 1 pad: conflict-free staging rows

template <typename real>
__device__ __forceinline__ void stage_tf(real* row, const Tf<real>& T) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    row[4 * r + 0] = T.r[3 * r + 0];
    row[4 * r + 1] = T.r[3 * r + 1];
    row[4 * r + 2] = T.r[3 * r + 2];
    row[4 * r + 3] = T.p[r];
  }
  row[12] = real(0);
  row[13] = real(0);
  row[14] = real(0);
  row[15] = real(1);
}

// copy the staged 16-real rows of this CTA out to global memory; consecutive
// threads write consecutive reals, `row_stride` reals between configurations
template <typename real>
__device__ __forceinline__ void flush_tile(real* __restrict__ out, const real* s_out,
                                           int64_t row0, int64_t n, int64_t row_stride,
                                           int64_t col0) {
  const int64_t left = n - row0;
  const int rows = left < kTile ? (int)left : kTile;
  for (int k = threadIdx.x; k < rows * 16; k += blockDim.x) {
    const int r = k >> 4, c = k & 15;
    out[(row0 + r) * row_stride + col0 + c] = s_out[r * kOutPad + c];
  }
}

template <typename real, int NDOF, bool ALL>
__global__ void __launch_bounds__(kTile)
    fk_kernel(const __grid_constant__ RobotDev<real> rb, const real* __restrict__ q, int64_t n,
              int layout, real* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* s_q = reinterpret_cast<real*>(smem_raw);
  real* s_out = s_q + kTile * NDOF;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;

  real qv[NDOF];
  load_q<real, NDOF>(qv, s_q, q, row0, n, layout);

  Tf<real> T;
  tf_load<real>(T, rb.base);
  real* my_row = s_out + threadIdx.x * kOutPad;
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    chain_step<real>(T, rb, i, qv[i]);
    if (ALL) {
      stage_tf<real>(my_row, T);
      __syncthreads();
      flush_tile<real>(out, s_out, row0, n, (int64_t)NDOF * 16, (int64_t)i * 16);
      __syncthreads();
    }
  }
  if (!ALL) {
    if (rb.has_tool_tip) {
      Tf<real> E;
      tf_mul<real>(E, T, rb.tip);
      T = E;
    }
    stage_tf<real>(my_row, T);
    __syncthreads();
    flush_tile<real>(out, s_out, row0, n, 16, 0);
  }
}

template <int NDOF>
__global__ void __launch_bounds__(kTile)
    fk_rpy_kernel(const __grid_constant__ RobotDev<double> rb, const double* __restrict__ q,
                  int64_t n, int layout, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_q = reinterpret_cast<double*>(smem_raw);
  double* s_out = s_q + kTile * NDOF;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  double qv[NDOF];
  load_q<double, NDOF>(qv, s_q, q, row0, n, layout);
  Tf<double> T;
  tf_load<double>(T, rb.base);
#pragma unroll
  for (int i = 0; i < NDOF; ++i) chain_step<double>(T, rb, i, qv[i]);
  if (rb.has_tool_tip) {
    Tf<double> E;
    tf_mul<double>(E, T, rb.tip);
    T = E;
  }
  // robot.py:73-77 (r_z uses T[0,1] and T[2,2]: the reference's own formula)
  const double s = sqrt(T.r[0] * T.r[0] + T.r[1] * T.r[1]);
  double* row = s_out + threadIdx.x * 7;
  row[0] = T.p[0];
  row[1] = T.p[1];
  row[2] = T.p[2];
  row[3] = atan2(-T.r[5], T.r[8]);
  row[4] = atan2(T.r[2], s);
  row[5] = atan2(-T.r[1], T.r[8]);
  __syncthreads();
  const int64_t left = n - row0;
  const int rows = left < kTile ? (int)left : kTile;
  for (int k = threadIdx.x; k < rows * 6; k += blockDim.x) {
    const int r = k / 6, c = k - 6 * r;
    out[(row0 + r) * 6 + c] = s_out[r * 7 + c];
  }
}

template <typename real>
cudaError_t launch_fk(const RobotDev<real>& rb, const real* q, int64_t n, int layout, real* out,
                      bool all_links, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  ACRO_DISPATCH_NDOF(rb.ndof, {
    const size_t smem = sizeof(real) * (kTile * NDOF + kTile * kOutPad);
    if (all_links)
      fk_kernel<real, NDOF, true><<<(unsigned)tiles, kTile, smem, st>>>(rb, q, n, layout, out);
    else
      fk_kernel<real, NDOF, false><<<(unsigned)tiles, kTile, smem, st>>>(rb, q, n, layout, out);
  });
  count_launch();
  return cudaGetLastError();
}

template cudaError_t launch_fk<double>(const RobotDev<double>&, const double*, int64_t, int,
                                       double*, bool, cudaStream_t);
template cudaError_t launch_fk<float>(const RobotDev<float>&, const float*, int64_t, int, float*,
                                      bool, cudaStream_t);

cudaError_t launch_fk_rpy(const RobotDev<double>& rb, const double* q, int64_t n, int layout,
                          double* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  ACRO_DISPATCH_NDOF(rb.ndof, {
    const size_t smem = sizeof(double) * (kTile * NDOF + kTile * 7);
    fk_rpy_kernel<NDOF><<<(unsigned)tiles, kTile, smem, st>>>(rb, q, n, layout, out);
  });
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
