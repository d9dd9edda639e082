// fk_cc_wave.cu - K2, all-FP64 variant ("wave"; the production kernel is fk_cc_filt.cu):
// fused FK -> collision organised as link-by-link wavefronts over a shared-memory work list.
//
// Same contract and bit-identical verdicts as fk_cc.cu (reference
// Robot.is_in_collision, src/acrobotics/robot.py:244-273 + :206-221,
// geometry.py:51-66, FCL boxBox2 via shapes.py:50-58); the difference is how the
// work is laid out on the machine.  Workload statistics:
// ~70 % of random configurations collide, half of them at the second link; only
// ~3 of the 96 link-box x scene-box pairs survive a face-axis cull, and 46 % of
// those really collide.  "One thread = one configuration from start to end"
// therefore idles most lanes most of the time (measured: 18 of 32).  Here:
//
//  * A CTA owns a tile of TILE configurations and a list of the ones still
//    undecided.  Wave s advances every listed configuration by one link: DH
//    step (running 3x4 transform kept in shared memory, SoA), the link's boxes
//    against the scene.  Hit configurations drop out; survivors are appended to
//    the next wave's list with warp-aggregated shared-memory atomics, so every
//    wave runs with full warps and a warp-uniform link index (robot constants
//    and the cull table come straight from the constant bank).
//  * Culling runs in FP32 on the FMA pipe: the scene box's three face axes
//    against the link box's bounding sphere, thresholds padded by a rigorous
//    bound on the FP32 error - a culled pair is one that FCL's own FP64 face-axis
//    test separates.  Everything that decides a verdict stays FP64.
//  * Pairs that survive the cull are queued (one pending pair per
//    configuration and round) and the exact 15-axis test runs over the dense
//    queue, again with full warps.
//  * Self collision is evaluated last, only for configurations that survived
//    the scene (the verdict is an OR, the order is free).
#include <cstdlib>
#include <string>

#include "acro_collide.cuh"
#include "acro_launch.h"

namespace acro {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// world pose of a robot box carried by frame F (geometry.py:56-57)
__device__ __forceinline__ void box_pose(const BoxDev<double>& bx, const Tf<double>& F, double* Rw,
                                         double* cw) {
  if (bx.rot_identity) {
#pragma unroll
    for (int k = 0; k < 9; ++k) Rw[k] = F.r[k];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      cw[r] = fma(F.r[3 * r + 2], bx.off[11],
                  fma(F.r[3 * r + 1], bx.off[7], fma(F.r[3 * r], bx.off[3], F.p[r])));
  } else {
    Tf<double> W;
    tf_mul<double>(W, F, bx.off);
#pragma unroll
    for (int k = 0; k < 9; ++k) Rw[k] = W.r[k];
#pragma unroll
    for (int r = 0; r < 3; ++r) cw[r] = W.p[r];
  }
}

// append `val` to a shared-memory list: one atomic per warp
__device__ __forceinline__ void warp_push(bool pred, uint16_t val, uint16_t* list, int* counter,
                                          unsigned lane) {
  const unsigned m = __ballot_sync(kFull, pred);
  if (m) {
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(kFull, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = val;
  }
}

// exact test with a cheap first exit: box 1's face axes against box 2's
// bounding sphere (implies FCL's u_k face test separates)
__device__ __forceinline__ bool pair_collides(const double A0, const double A1, const double A2,
                                              const double* R1, const double* c1,
                                              const double* b2) {
  const double p0 = b2[kOB_T + 0] - c1[0];
  const double p1 = b2[kOB_T + 1] - c1[1];
  const double p2 = b2[kOB_T + 2] - c1[2];
  const double rb = b2[kOB_RAD];
  const double z0 = fma(R1[6], p2, fma(R1[3], p1, R1[0] * p0));
  const double z1 = fma(R1[7], p2, fma(R1[4], p1, R1[1] * p0));
  const double z2 = fma(R1[8], p2, fma(R1[5], p1, R1[2] * p0));
  if ((fabs(z0) > A0 + rb) | (fabs(z1) > A1 + rb) | (fabs(z2) > A2 + rb)) return false;
  double g;
  return !sat_separated<double>(A0, A1, A2, R1, c1, b2, 0.0, g);
}

}  // namespace

template <int TILE, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
    fk_cc_wave_kernel(const __grid_constant__ RobotDev<double> rb,
                      const __grid_constant__ CullTable ct, const SceneView<double> sc,
                      const double* __restrict__ q, int64_t n, int layout,
                      uint32_t* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_T = reinterpret_cast<double*>(smem_raw);                         // [12][TILE]
  unsigned long long* s_cand = reinterpret_cast<unsigned long long*>(s_T + 12 * TILE);  // [TILE]
  double* s_scene = reinterpret_cast<double*>(s_cand + TILE);                // [n_scene][20]
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_scene + sc.n * kSceneStride);  // [2][TILE]
  uint16_t* s_items = s_list + 2 * TILE;                                     // [3][TILE]
  uint32_t* s_hit = reinterpret_cast<uint32_t*>(s_items + 3 * TILE);         // [TILE/32]
  int* s_cnt = reinterpret_cast<int*>(s_hit + TILE / 32);  // [0..1] list sizes, [2..4] item counts

  const int t = threadIdx.x;
  const unsigned lane = t & 31;
  const int warp0 = t & ~31;
  const int64_t tile0 = (int64_t)blockIdx.x * TILE;
  const int tile_n = (int)((n - tile0) < TILE ? (n - tile0) : TILE);
  const int n_scene = sc.n;
  const int ndof = rb.ndof;

  for (int k = t; k < n_scene * kSceneStride; k += THREADS) s_scene[k] = sc.boxes[k];
  for (int k = t; k < TILE / 32; k += THREADS) s_hit[k] = 0u;
  for (int k = t; k < tile_n; k += THREADS) s_list[k] = (uint16_t)k;
  if (t == 0) {
    s_cnt[0] = tile_n;
    s_cnt[1] = 0;
    s_cnt[2] = s_cnt[3] = s_cnt[4] = 0;
  }
  __syncthreads();

  auto load_T = [&](Tf<double>& T, int e) {
#pragma unroll
    for (int k = 0; k < 9; ++k) T.r[k] = s_T[k * TILE + e];
#pragma unroll
    for (int k = 0; k < 3; ++k) T.p[k] = s_T[(9 + k) * TILE + e];
  };
  auto store_T = [&](const Tf<double>& T, int e) {
#pragma unroll
    for (int k = 0; k < 9; ++k) s_T[k * TILE + e] = T.r[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) s_T[(9 + k) * TILE + e] = T.p[k];
  };
  auto load_q = [&](int e, int i) -> double {
    const int64_t row = tile0 + e;
    return layout == ACRO_Q_AOS ? q[row * ndof + i] : q[(int64_t)i * n + row];
  };

  // slots: 0 = base boxes at tf_base, 1..ndof = DH step of link slot-1 then its boxes,
  // ndof+1 = tool boxes on the last link frame.  Empty base / tool slots are skipped.
  const int first_slot = rb.slot_begin[1] > rb.slot_begin[0] ? 0 : 1;
  const int last_slot = rb.slot_begin[ndof + 2] > rb.slot_begin[ndof + 1] ? ndof + 1 : ndof;
  int cur = 0;

  for (int s = first_slot; s <= last_slot && n_scene > 0; ++s) {
    const bool has_chain = s >= 1 && s <= ndof;
    const int b_begin = rb.slot_begin[s];
    const int nb = rb.slot_begin[s + 1] - b_begin;
    const int n_iter = nb > 0 ? nb : 1;
    const int nxt = cur ^ 1;
    uint16_t* list_cur = s_list + cur * TILE;
    uint16_t* list_nxt = s_list + nxt * TILE;
    for (int bi = 0; bi < n_iter; ++bi) {
      const bool do_chain = has_chain && bi == 0;
      const bool has_box = bi < nb;
      const bool last_box = bi == n_iter - 1;
      const bool from_base = s == 0 || (s == 1 && bi == 0);
      const BoxDev<double>& bx = rb.boxes[has_box ? b_begin + bi : 0];
      const int n_cur = s_cnt[cur];

      // ---- pass A: advance every listed configuration by this link / box, cull in FP32
      for (int base = warp0; base < n_cur; base += THREADS) {
        const int idx = base + (int)lane;
        const bool valid = idx < n_cur;
        const int e = valid ? list_cur[idx] : 0;
        const bool alive = valid && !((s_hit[e >> 5] >> (e & 31)) & 1u);
        unsigned long long c = 0;
        if (alive) {
          Tf<double> T;
          if (from_base)
            tf_load<double>(T, rb.base);
          else
            load_T(T, e);
          if (do_chain) {
            chain_step<double>(T, rb, s - 1, load_q(e, s - 1));
            store_T(T, e);
          }
          if (has_box) {
            double Rw[9], cw[3];
            box_pose(bx, T, Rw, cw);
            const float cx = (float)cw[0], cy = (float)cw[1], cz = (float)cw[2];
            // padded bounding radius: the FP32 rounding of y_k stays below
            // 4 * 2^-24 * (|c|_1 + |t'|_1); the table pads the |t'| share
            const float ra = __double2float_ru(bx.radius * (1.0 + 1e-6)) +
                             1e-6f * (fabsf(cx) + fabsf(cy) + fabsf(cz));
#pragma unroll 4
            for (int j = 0; j < n_scene; ++j) {
              const float* ce = ct.e[j];
              const float y0 = fmaf(ce[0], cx, fmaf(ce[1], cy, fmaf(ce[2], cz, ce[3])));
              const float y1 = fmaf(ce[4], cx, fmaf(ce[5], cy, fmaf(ce[6], cz, ce[7])));
              const float y2 = fmaf(ce[8], cx, fmaf(ce[9], cy, fmaf(ce[10], cz, ce[11])));
              const float m =
                  fmaxf(fmaxf(fabsf(y0) - ce[12], fabsf(y1) - ce[13]), fabsf(y2) - ce[14]);
              if (!(m > ra)) c |= (1ull << j);
            }
          }
          if (c) s_cand[e] = c;
        }
        warp_push(alive && c != 0, (uint16_t)e, s_items, &s_cnt[2], lane);
        warp_push(alive && c == 0 && last_box, (uint16_t)e, list_nxt, &s_cnt[nxt], lane);
      }
      __syncthreads();

      // ---- rounds: one pending pair per configuration and round, dense over the queue
      int r = 0;
      for (;;) {
        const int ni = s_cnt[2 + r % 3];
        if (ni == 0) break;
        if (t == 0) s_cnt[2 + (r + 2) % 3] = 0;  // consumed in round r-1, refilled in round r+1
        const uint16_t* items = s_items + (r % 3) * TILE;
        uint16_t* items_next = s_items + ((r + 1) % 3) * TILE;
        for (int base = warp0; base < ni; base += THREADS) {
          const int k = base + (int)lane;
          const bool valid = k < ni;
          const int e = valid ? items[k] : 0;
          unsigned long long c = valid ? s_cand[e] : 0ull;
          bool hit = false;
          if (valid) {
            const int j = __ffsll((long long)c) - 1;
            c &= c - 1;
            Tf<double> T;
            if (s == 0)
              tf_load<double>(T, rb.base);
            else
              load_T(T, e);
            double Rw[9], cw[3];
            box_pose(bx, T, Rw, cw);
            hit = pair_collides(bx.half[0], bx.half[1], bx.half[2], Rw, cw,
                                s_scene + j * kSceneStride);
            if (hit)
              atomicOr(&s_hit[e >> 5], 1u << (e & 31));
            else if (c)
              s_cand[e] = c;
          }
          warp_push(valid && !hit && c != 0, (uint16_t)e, items_next, &s_cnt[2 + (r + 1) % 3], lane);
          warp_push(valid && !hit && c == 0 && last_box, (uint16_t)e, list_nxt, &s_cnt[nxt], lane);
        }
        __syncthreads();
        ++r;
      }
      if (t == 0) {
        s_cnt[2] = s_cnt[3] = s_cnt[4] = 0;
        if (last_box) s_cnt[cur] = 0;  // this list is consumed; it becomes the next "next"
      }
      __syncthreads();
    }
    cur = nxt;
  }

  // ---- self collision (robot.py:206-221) for the configurations that survived the scene.
  // FK is recomputed; the world poses of the boxes that are the earlier member of a pair
  // are parked in the (now free) transform area, one column per thread.
  if (rb.check_self && rb.pair_begin[rb.n_boxes] > 0) {
    double* scratch = s_T;
    const int n_cur = s_cnt[cur];
    const uint16_t* list_cur = s_list + cur * TILE;
    for (int base = warp0; base < n_cur; base += THREADS) {
      const int idx = base + (int)lane;
      const bool valid = idx < n_cur;
      const int e = valid ? list_cur[idx] : 0;
      bool hit = !valid;  // lanes without an entry just follow along
      Tf<double> T;
      tf_load<double>(T, rb.base);
      for (int s = 0; s <= ndof + 1; ++s) {
        if (s >= 1 && s <= ndof) chain_step<double>(T, rb, s - 1, valid ? load_q(e, s - 1) : 0.0);
        for (int b = rb.slot_begin[s]; b < rb.slot_begin[s + 1]; ++b) {
          const BoxDev<double>& bx = rb.boxes[b];
          if (rb.pair_begin[b + 1] == rb.pair_begin[b] && bx.save_slot < 0) continue;
          double Rw[9], cw[3];
          box_pose(bx, T, Rw, cw);
          for (int k = rb.pair_begin[b]; k < rb.pair_begin[b + 1]; ++k) {
            const BoxDev<double>& other = rb.boxes[rb.partner[k]];
            double ob[16];
#pragma unroll
            for (int m = 0; m < 12; ++m) ob[m] = scratch[(other.save_slot * 12 + m) * THREADS + t];
            ob[kOB_H + 0] = other.half[0];
            ob[kOB_H + 1] = other.half[1];
            ob[kOB_H + 2] = other.half[2];
            ob[kOB_RAD] = other.radius;
            if (!hit && !cull_separated<double>(bx.half[0], bx.half[1], bx.half[2], bx.radius, Rw,
                                                cw, ob, 0.0)) {
              double g;
              hit = !sat_separated<double>(bx.half[0], bx.half[1], bx.half[2], Rw, cw, ob, 0.0, g);
              if (hit) atomicOr(&s_hit[e >> 5], 1u << (e & 31));
            }
          }
          if (bx.save_slot >= 0) {
#pragma unroll
            for (int m = 0; m < 9; ++m) scratch[(bx.save_slot * 12 + m) * THREADS + t] = Rw[m];
#pragma unroll
            for (int m = 0; m < 3; ++m) scratch[(bx.save_slot * 12 + 9 + m) * THREADS + t] = cw[m];
          }
        }
        if (__all_sync(kFull, hit)) break;
      }
    }
  }

  __syncthreads();
  const int words = (tile_n + 31) / 32;
  for (int k = t; k < words; k += THREADS) mask[(tile0 >> 5) + k] = s_hit[k];
}

// capacity of the self-collision scratch area (saved boxes per configuration)
template <int TILE, int THREADS>
constexpr int wave_max_saved() {
  return TILE / THREADS;
}

template <int TILE, int THREADS, int MIN_CTAS>
static cudaError_t launch_wave(const RobotDev<double>& rb, const CullTable& ct,
                               SceneView<double> sc, const double* q, int64_t n, int layout,
                               uint32_t* mask, cudaStream_t st) {
  const int64_t tiles = (n + TILE - 1) / TILE;
  if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
  const size_t smem = sizeof(double) * 12 * TILE + sizeof(unsigned long long) * TILE +
                      sizeof(double) * sc.n * kSceneStride + sizeof(uint16_t) * 5 * TILE +
                      sizeof(uint32_t) * (TILE / 32) + sizeof(int) * 8;
  auto kern = fk_cc_wave_kernel<TILE, THREADS, MIN_CTAS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)tiles, THREADS, smem, st>>>(rb, ct, sc, q, n, layout, mask);
  count_launch();
  return cudaGetLastError();
}

bool wave_supports(const RobotDev<double>& rb) {
  int n_saved = 0;
  for (int b = 0; b < rb.n_boxes; ++b)
    if (rb.boxes[b].save_slot + 1 > n_saved) n_saved = rb.boxes[b].save_slot + 1;
  return !rb.check_self || n_saved <= 4;
}

cudaError_t launch_fk_cc_wave(const RobotDev<double>& rb, const CullTable& ct,
                              SceneView<double> sc, const double* q, int64_t n, int layout,
                              uint32_t* mask, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  // ACRO_K2_TILE selects the tile geometry for tuning runs: "512x128" (default),
  // "1024x256", "256x64"
  static const int variant = [] {
    const char* v = std::getenv("ACRO_K2_TILE");
    if (!v) return 0;
    const std::string s(v);
    if (s == "1024x256") return 1;
    if (s == "256x64") return 2;
    if (s == "512x128x3") return 3;
    return 0;
  }();
  switch (variant) {
    case 1: return launch_wave<1024, 256, 2>(rb, ct, sc, q, n, layout, mask, st);
    case 2: return launch_wave<256, 64, 8>(rb, ct, sc, q, n, layout, mask, st);
    case 3: return launch_wave<512, 128, 3>(rb, ct, sc, q, n, layout, mask, st);
    default: return launch_wave<512, 128, 4>(rb, ct, sc, q, n, layout, mask, st);
  }
}

}  // namespace acro
