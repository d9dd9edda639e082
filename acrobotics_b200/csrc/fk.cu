// fk.cu - K1: batched forward kinematics of serial DH chains.
//
// Replaces reference Robot.fk (src/acrobotics/robot.py:59-69), fk_all_links
// (robot.py:82-91) and fk_rpy (robot.py:71-80) built on
// LinkKinematics.get_link_relative_transform (src/acrobotics/link.py:35-58).
//
// One thread per configuration, transform chained in registers.  HBM bound
// (48 B in, 128 B out per configuration for a 6-dof fk): joint rows are read
// with 16-byte loads, the 4x4 results leave through a bank-conflict-free
// shared-memory tile as whole 128-byte lines (a 16-double row per thread
// written directly would be a 128-byte-strided store of 16-byte pieces).
#include "acro_launch.h"

namespace acro {

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

// Every thread reads its joint row with 16-byte loads (a warp covers one contiguous span);
// results leave through a shared-memory tile as whole 128-byte lines (stage_store_rows).
template <typename real, int NDOF, bool ALL>
__global__ void __launch_bounds__(kTile)
    fk_kernel(const __grid_constant__ RobotDev<real> rb, const real* __restrict__ q, int64_t n,
              int layout, real* __restrict__ out, bool vec_in, bool vec_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t idx = row0 + threadIdx.x;
  real qv[NDOF];
  load_q_direct<real, NDOF>(qv, q, idx, n, layout, vec_in);

  Tf<real> T;
  tf_load<real>(T, rb.base);
  alignas(16) real row[16];
#pragma unroll
  for (int i = 0; i < NDOF; ++i) {
    chain_step<real>(T, rb, i, qv[i]);
    if (ALL) {
      stage_tf<real>(row, T);
      if (i > 0) __syncthreads();  // previous link's tile has left shared memory
      stage_store_rows<real, 16>(smem_raw, row, out, row0, n, (int64_t)NDOF * 16, (int64_t)i * 16,
                                 vec_out);
    }
  }
  if (!ALL) {
    if (rb.has_tool_tip) {
      Tf<real> E;
      tf_mul<real>(E, T, rb.tip);
      T = E;
    }
    stage_tf<real>(row, T);
    stage_store_rows<real, 16>(smem_raw, row, out, row0, n, 16, 0, vec_out);
  }
}

template <int NDOF>
__global__ void __launch_bounds__(kTile)
    fk_rpy_kernel(const __grid_constant__ RobotDev<double> rb, const double* __restrict__ q,
                  int64_t n, int layout, double* __restrict__ out, bool vec_in, bool vec_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t idx = row0 + threadIdx.x;
  double qv[NDOF];
  load_q_direct<double, NDOF>(qv, q, idx, n, layout, vec_in);
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
  const double s = sqrt_f64(T.r[0] * T.r[0] + T.r[1] * T.r[1]);
  alignas(16) double row[6];
  row[0] = T.p[0];
  row[1] = T.p[1];
  row[2] = T.p[2];
  row[3] = atan2_f64(-T.r[5], T.r[8]);
  row[4] = atan2_f64(T.r[2], s);
  row[5] = atan2_f64(-T.r[1], T.r[8]);
  stage_store_rows<double, 6>(smem_raw, row, out, row0, n, 6, 0, vec_out);
}

template <typename real>
cudaError_t launch_fk(const RobotDev<real>& rb, const real* q, int64_t n, int layout, real* out,
                      bool all_links, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t tiles = (n + kTile - 1) / kTile;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  ACRO_DISPATCH_NDOF(rb.ndof, {
    const bool vi = aligned16(q), vo = aligned16(out);
    const size_t smem = row_stage_bytes<real, 16>();
    if (all_links)
      fk_kernel<real, NDOF, true><<<(unsigned)tiles, kTile, smem, st>>>(rb, q, n, layout, out, vi, vo);
    else
      fk_kernel<real, NDOF, false><<<(unsigned)tiles, kTile, smem, st>>>(rb, q, n, layout, out, vi, vo);
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
    fk_rpy_kernel<NDOF><<<(unsigned)tiles, kTile, row_stage_bytes<double, 6>(), st>>>(rb, q, n, layout, out, aligned16(q),
                                                            aligned16(out));
  });
  count_launch();
  return cudaGetLastError();
}

}  // namespace acro
