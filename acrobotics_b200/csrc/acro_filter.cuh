// acro_filter.cuh - the FP32 filter of the fused FK -> collision evaluation, shared by the
// filtered K2 kernel (fk_cc_filt.cu) and the path sweep (path.cu).  See fk_cc_filt.cu for the
// contract: kHit / kSep verdicts are certain for the FP64 evaluation, everything else is
// reported as ambiguous and must be recomputed in FP64 by the caller.
#pragma once
#include "acro_collide.cuh"

namespace acro {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kFThreads = 128;
constexpr int kItemCap = 64;      // candidate pairs per warp and round
#ifndef ACRO_FILTER_MIN_CTAS
#define ACRO_FILTER_MIN_CTAS 6
#endif
#ifndef ACRO_FILTER_DYNAMIC
#define ACRO_FILTER_DYNAMIC 1
#endif
#ifndef ACRO_FILTER_PIN_IDS
#define ACRO_FILTER_PIN_IDS 1
#endif
#ifndef ACRO_SAT_PRETEST
#define ACRO_SAT_PRETEST 0
#endif
#ifndef ACRO_FILTER_DIRECT_MIN
#define ACRO_FILTER_DIRECT_MIN 28
#endif
constexpr int kSceneStrideF = 16; // floats per scene box in shared memory, see kScenePerm
// Shared-memory order of a scene box: the rotation is stored as the lane pairs the packed SAT
// consumes - (R00,R01) (R10,R11) (R20,R21) then R02 R12 R22 - so that 16-byte loads drop every
// pair into an aligned register pair; then T2 [9..11], half extents [12..14], radius [15].
__device__ __constant__ const int kScenePerm[16] = {0, 1, 3, 4, 6, 7, 2, 5, 8, 9, 10, 11, 12, 13, 14, 15};

enum : int { kSep = 0, kHit = 1, kAmb = 2 };

// sin/cos of a joint angle for the FP32 chain.  The argument is reduced in FP64
// (q - k pi/2, exact to ~1e-16 for |q| < 1e8), the polynomials run in FP32 on
// [-pi/4, pi/4] (Cephes sinf/cosf minimax coefficients, abs error < 1.2e-7).
__device__ __forceinline__ void sincos_filter(double q, float& s, float& c) {
  // k = rint(q * 2/pi) by the magic-number trick (|q| < 2^26: the sum is exact to the unit,
  // its low word is k in two's complement): two DADDs instead of F2I + I2F.  The second term
  // of pi/2 (6.1e-17) would change r by |k| 6.1e-17 < 2.7e-9: part of the bound's e_t.
  const double tm = fma(q, 0.63661977236758134308, 6755399441055744.0);  // 1.5 * 2^52
  const int k = __double2loint(tm);
  const double kd = tm - 6755399441055744.0;
  const double r = fma(-kd, 1.57079632679489655800e+00, q);
  const float x = (float)r;
  const float z = x * x;
  float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
  sp = fmaf(sp * z, x, x);
  float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
  cp = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
  const float a = (k & 1) ? cp : sp;
  const float b = (k & 1) ? sp : cp;
  s = (k & 2) ? -a : a;
  c = ((k + 1) & 2) ? -b : b;
}

// FCL boxBox2 in FP32 with tolerance: kSep when some axis separates by more than
// delta, kHit when every axis overlaps by more than delta, kAmb otherwise.
__device__ __forceinline__ int sat_filter(const float A0, const float A1, const float A2,
                                          const float* R1, const float* c1, const float* R2,
                                          const float* T2, const float B0, const float B1,
                                          const float B2, const float delta, const float rA,
                                          const float rB) {
  const float p0 = T2[0] - c1[0], p1 = T2[1] - c1[1], p2 = T2[2] - c1[2];
  const float A[3] = {A0, A1, A2};
  const float B[3] = {B0, B1, B2};
  float R[9], Q[9];
  float g = -1e30f;
#if ACRO_SAT_PRETEST
  // Cheap rejections before R = R1^T R2 is formed: a face axis of one box against the bounding
  // sphere of the other.  sum_i |R_ij| A_i <= |A| <= rA (rounded up on the host), so
  // |t_j| - (B_j + rA) is a lower bound of FCL's face-axis quantity; likewise for box 1's axes.
  {
    float ts[3], ps[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      ts[j] = fmaf(R2[6 + j], p2, fmaf(R2[3 + j], p1, R2[j] * p0));
      ps[j] = fmaf(R1[6 + j], p2, fmaf(R1[3 + j], p1, R1[j] * p0));
    }
    const float gs = fmaxf(fmaxf(fmaxf(fabsf(ts[0]) - B0, fabsf(ts[1]) - B1), fabsf(ts[2]) - B2) - rA,
                           fmaxf(fmaxf(fabsf(ps[0]) - A0, fabsf(ps[1]) - A1), fabsf(ps[2]) - A2) - rB);
    if (gs > delta) return kSep;
  }
#endif
#if ACRO_F32X2
  // columns 0 and 1 of R = R1^T R2 as packed lane pairs (FFMA2 with a scalar-broadcast operand),
  // column 2 scalar; same operation order as the scalar form below, hence the same bits
  const float2 r2p0 = make_float2(R2[0], R2[1]), r2p1 = make_float2(R2[3], R2[4]),
               r2p2 = make_float2(R2[6], R2[7]);
  float2 Rp[3], Qp[3];
  float Rs[3], Qs[3];
  (void)R; (void)Q;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Rp[i] = fma2(R1[6 + i], r2p2, fma2(R1[3 + i], r2p1, mul2(R1[i], r2p0)));
    Rs[i] = fmaf(R1[6 + i], R2[8], fmaf(R1[3 + i], R2[5], R1[i] * R2[2]));
    Qp[i] = make_float2(fabsf(Rp[i].x), fabsf(Rp[i].y));
    Qs[i] = fabsf(Rs[i]);
  }
  {
    const float2 tp = fma2(p2, r2p2, fma2(p1, r2p1, mul2(p0, r2p0)));
    const float ts = fmaf(R2[8], p2, fmaf(R2[5], p1, R2[2] * p0));
    const float2 radp = __fadd2_rn(fma2(A2, Qp[2], fma2(A1, Qp[1], mul2(A0, Qp[0]))),
                                   make_float2(B0, B1));
    const float rads = fmaf(Qs[2], A2, fmaf(Qs[1], A1, Qs[0] * A0)) + B2;
    g = fmaxf(g, fmaxf(fmaxf(fabsf(tp.x) - radp.x, fabsf(tp.y) - radp.y), fabsf(ts) - rads));
  }
  if (g > delta) return kSep;
  float pp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pp[i] = fmaf(R1[6 + i], p2, fmaf(R1[3 + i], p1, R1[i] * p0));
#pragma unroll
  for (int i = 0; i < 3; ++i)
    g = fmaxf(g, fabsf(pp[i]) - (fmaf(Qs[i], B2, fmaf(Qp[i].y, B1, Qp[i].x * B0)) + A[i]));
  // (no exit here: 5 % of the separated pairs would take it, and a warp waits for its slowest
  // lane anyway)
  const float2 fudge = make_float2(1.0e-6f, 1.0e-6f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Qp[i] = __fadd2_rn(Qp[i], fudge);
    Qs[i] += 1.0e-6f;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    const float2 tp = fma2(pp[i2], Rp[i1], mul2(-pp[i1], Rp[i2]));
    const float ts = fmaf(pp[i2], Rs[i1], -(pp[i1] * Rs[i2]));
    // box 2's share of the radius: B[j1] Q[i][j2] + B[j2] Q[i][j1] for j = 0, 1, 2
    const float in0 = fmaf(B1, Qs[i], B2 * Qp[i].y);
    const float in1 = fmaf(B2, Qp[i].x, B0 * Qs[i]);
    const float in2 = fmaf(B0, Qp[i].y, B1 * Qp[i].x);
    const float2 radp = fma2(A[i1], Qp[i2], fma2(A[i2], Qp[i1], make_float2(in0, in1)));
    const float rads = fmaf(A[i1], Qs[i2], fmaf(A[i2], Qs[i1], in2));
    g = fmaxf(g, fmaxf(fmaxf(fabsf(tp.x) - radp.x, fabsf(tp.y) - radp.y), fabsf(ts) - rads));
  }
#else
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      R[3 * i + j] = fmaf(R1[6 + i], R2[6 + j], fmaf(R1[3 + i], R2[3 + j], R1[i] * R2[j]));
      Q[3 * i + j] = fabsf(R[3 * i + j]);
    }
    const float t = fmaf(R2[6 + j], p2, fmaf(R2[3 + j], p1, R2[j] * p0));
    const float rad = fmaf(Q[6 + j], A2, fmaf(Q[3 + j], A1, Q[j] * A0)) + B[j];
    g = fmaxf(g, fabsf(t) - rad);
  }
  if (g > delta) return kSep;
  float pp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) pp[i] = fmaf(R1[6 + i], p2, fmaf(R1[3 + i], p1, R1[i] * p0));
#pragma unroll
  for (int i = 0; i < 3; ++i)
    g = fmaxf(g, fabsf(pp[i]) -
                     (fmaf(Q[3 * i + 2], B2, fmaf(Q[3 * i + 1], B1, Q[3 * i] * B0)) + A[i]));
  if (g > delta) return kSep;
#pragma unroll
  for (int k = 0; k < 9; ++k) Q[k] += 1.0e-6f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const float t = fmaf(pp[i2], R[3 * i1 + j], -(pp[i1] * R[3 * i2 + j]));
      const float rad = fmaf(A[i1], Q[3 * i2 + j],
                             fmaf(A[i2], Q[3 * i1 + j],
                                  fmaf(B[j1], Q[3 * i + j2], B[j2] * Q[3 * i + j1])));
      g = fmaxf(g, fabsf(t) - rad);
    }
  }
#endif
  if (g > delta) return kSep;
  return g < -delta ? kHit : kAmb;
}

template <typename MaskT>
__device__ __forceinline__ int popc_t(MaskT m);
template <>
__device__ __forceinline__ int popc_t<uint32_t>(uint32_t m) {
  return __popc(m);
}
template <>
__device__ __forceinline__ int popc_t<uint64_t>(uint64_t m) {
  return __popcll(m);
}
template <typename MaskT>
__device__ __forceinline__ int ffs_t(MaskT m);
template <>
__device__ __forceinline__ int ffs_t<uint32_t>(uint32_t m) {
  return __ffs((int)m) - 1;
}
template <>
__device__ __forceinline__ int ffs_t<uint64_t>(uint64_t m) {
  return __ffsll((long long)m) - 1;
}


// per-warp working set of the filter (shared-memory pointers already offset to this warp)
struct FilterCtx {
  const float* s_scene;  // n_scene * 16 floats (CTA wide)
  float* s_saved;        // n_saved * 12 * 32 floats (this warp)
  const float* s_half;   // kMaxSaved * 4 floats (CTA wide)
  uint16_t* list;        // kItemCap entries (this warp)
  unsigned char* extra;  // the caller's own per-warp bytes (start of this warp's block)
  int n_scene;
};

// Shared-memory layout: [saved-slot half extents: kMaxSaved * 4 floats][scene: n_scene * 16
// floats][one block per warp: `extra` caller bytes, the candidate list, the saved poses].
// Everything a warp touches besides the CTA-wide tables sits at a fixed offset from ONE base
// address, which keeps the address arithmetic (and its rematerialisation under register
// pressure) out of the inner loops.
__host__ __device__ inline size_t filter_warp_block_bytes(int n_saved, size_t extra) {
  return extra + sizeof(uint16_t) * kItemCap + sizeof(float) * n_saved * 12 * 32;
}

// shared memory the filter needs per CTA of `warps` warps (`extra`: caller bytes per warp,
// a multiple of 16)
inline size_t filter_ctx_bytes(int warps, int n_scene, int n_saved, size_t extra = 0) {
  return sizeof(float) * kMaxSaved * 4 + sizeof(float) * n_scene * kSceneStrideF +
         warps * filter_warp_block_bytes(n_saved, extra);
}

// carve the per-warp context out of `base` (16-byte aligned), CTA of `warps` warps
__device__ __forceinline__ FilterCtx filter_ctx_carve(unsigned char* base, int warps, int warp,
                                                      int n_scene, int n_saved, size_t extra = 0) {
  FilterCtx c;
  float* s_half = reinterpret_cast<float*>(base);
  float* s_scene = s_half + kMaxSaved * 4;
  unsigned char* wb = reinterpret_cast<unsigned char*>(s_scene + n_scene * kSceneStrideF) +
                      warp * filter_warp_block_bytes(n_saved, extra);
  c.s_half = s_half;
  c.s_scene = s_scene;
  c.extra = wb;
  c.list = reinterpret_cast<uint16_t*>(wb + extra);
  c.s_saved = reinterpret_cast<float*>(wb + extra + sizeof(uint16_t) * kItemCap);
  c.n_scene = n_scene;
  (void)warps;
  return c;
}

// CTA-wide staging of the scene and the saved-slot half extents (call before __syncthreads)
__device__ __forceinline__ void filter_ctx_stage(unsigned char* base, int warps,
                                                 const RobotDev<float>& rb,
                                                 const SceneView<float>& sc, int n_saved) {
  float* s_half = reinterpret_cast<float*>(base);
  float* s_scene = s_half + kMaxSaved * 4;
  (void)warps;
  (void)n_saved;
  for (int k = threadIdx.x; k < sc.n * kSceneStrideF; k += blockDim.x) {
    const int j = k >> 4, e = k & 15;
    s_scene[k] = sc.boxes[j * kSceneStride + kScenePerm[e]];
  }
  for (int b = threadIdx.x; b < rb.n_boxes; b += blockDim.x) {
    const int sl = rb.boxes[b].save_slot;
    if (sl >= 0) {
      s_half[4 * sl + 0] = rb.boxes[b].half[0];
      s_half[4 * sl + 1] = rb.boxes[b].half[1];
      s_half[4 * sl + 2] = rb.boxes[b].half[2];
      s_half[4 * sl + 3] = rb.boxes[b].radius;
    }
  }
}

// One warp, 32 configurations (one per lane; `active` false = no configuration in this lane).
// qfn(i) returns this lane's joint value i (double).  All 32 lanes must call it together.
// Slot window: the chain runs over slots 0 .. slot_end; boxes of slots whose bit is clear in
// check_mask get their pose computed (and parked for later self pairs) but are not tested - a
// caller that splits the work into passes gives every slot to exactly one of them.
// slot_end < 0 means "to the robot's last slot".  `retired` = the lane needs no further pass
// (surely colliding, or handed to the FP64 pass / skipped).
// PRISM = false compiles the prismatic-joint handling out (rb.prismatic_mask must be 0).
// ROTID = true: every box offset of the robot is a pure translation (the common case: the
// reference's example robots); the box rotation then IS the link frame's and no copy is made.
template <int NDOF, typename MaskT, bool PRISM, bool ROTID = false, typename QFn>
__device__ __forceinline__ void filter_chunk(const RobotDev<float>& rb, const GridView& grid,
                                             const FilterCtx& ctx, const unsigned lane,
                                             const bool active, QFn qfn, bool& hit, bool& amb,
                                             bool& retired, const unsigned check_mask = ~0u,
                                             const int slot_end = -1) {
  const float* s_scene = ctx.s_scene;
  float* s_saved = ctx.s_saved;
  const float* s_half = ctx.s_half;
  uint16_t* list = ctx.list;
  const int n_scene = ctx.n_scene;
  const float delta = grid.delta;
  const bool do_self = rb.check_self != 0;
  const int ndof = NDOF;
  const int robot_last = rb.slot_begin[ndof + 2] > rb.slot_begin[ndof + 1] ? ndof + 1 : ndof;
  const int last_slot = (slot_end < 0 || slot_end > robot_last) ? robot_last : slot_end;
bool done = !active;
hit = false;
amb = false;

  Tf<float> T;
  tf_load<float>(T, rb.base);

#pragma unroll 1
  for (int slot = 0; slot <= last_slot; ++slot) {
    if (slot >= 1 && slot <= ndof) {
      if (__all_sync(kFull, done)) break;
      const int i = slot - 1;
      double qi = qfn(i);  // lanes without a configuration read an arbitrary (unused) value
      const bool prismatic = PRISM && ((rb.prismatic_mask >> i) & 1);
      // joint values the FP32 chain cannot follow within its error bound: revolute |q| >= 2^26
      // (6.7e7; one integer compare on the exponent, which also catches NaN and Inf),
      // prismatic travel beyond the range the bound covers
      bool out_of_range;
      if (prismatic)
        out_of_range = !(fabs(qi) <= (double)grid.q_lin_max);
      else
        out_of_range = ((unsigned)__double2hiint(qi) & 0x7fffffffu) >= 0x41900000u;
      if (out_of_range) {
        if (!done) {
          amb = !(grid.skip_nan && qi != qi);
          done = true;  // decided by the FP64 pass (or an absent IK branch)
        }
        qi = 0.0;
      }
      float s, c;
      const float4 dh = rb.dh4[i];  // (cos alpha, sin alpha, a, d)
      if (PRISM) {
        sincos_filter(prismatic ? (double)rb.theta[i] : qi, s, c);
        tf_mul_dh<float>(T, c, s, dh.x, dh.y, dh.z, prismatic ? (float)qi : dh.w);
      } else {
        sincos_filter(qi, s, c);
        tf_mul_dh<float>(T, c, s, dh.x, dh.y, dh.z, dh.w);
      }
    }
    for (int b = rb.slot_begin[slot]; b < rb.slot_begin[slot + 1]; ++b) {
      // everything the filter needs of the box: its 32-byte BoxHot record (two constant loads)
      const float4 h0 = *reinterpret_cast<const float4*>(&rb.hot[b].off[0]);
      const float4 h1 = *reinterpret_cast<const float4*>(&rb.hot[b].half[0]);
      const unsigned meta = __float_as_uint(h1.w);
      const float box_radius = h0.w;
      float Rw_copy[9], cw[3];
      const float* Rw = ROTID ? T.r : Rw_copy;
      if (ROTID || ((meta >> 14) & 1u)) {
        if (!ROTID) {
#pragma unroll
          for (int k = 0; k < 9; ++k) Rw_copy[k] = T.r[k];
        }
#if ACRO_F32X2
        {  // rows 0 and 1 as one packed lane pair (the pairing of tf_mul_dh), row 2 scalar
          const float2 c01 = fma2(h0.z, make_float2(T.r[2], T.r[5]),
                                  fma2(h0.y, make_float2(T.r[1], T.r[4]),
                                       fma2(h0.x, make_float2(T.r[0], T.r[3]),
                                            make_float2(T.p[0], T.p[1]))));
          cw[0] = c01.x;
          cw[1] = c01.y;
          cw[2] = fmaf(T.r[8], h0.z, fmaf(T.r[7], h0.y, fmaf(T.r[6], h0.x, T.p[2])));
        }
#else
#pragma unroll
        for (int r = 0; r < 3; ++r)
          cw[r] = fmaf(T.r[3 * r + 2], h0.z,
                       fmaf(T.r[3 * r + 1], h0.y, fmaf(T.r[3 * r], h0.x, T.p[r])));
#endif
      } else {
        Tf<float> W;
        tf_mul<float>(W, T, rb.boxes[b].off);
#pragma unroll
        for (int k = 0; k < 9; ++k) Rw_copy[k] = W.r[k];
#pragma unroll
        for (int r = 0; r < 3; ++r) cw[r] = W.p[r];
      }
      const float A0 = h1.x, A1 = h1.y, A2 = h1.z;

      // ---- candidates: scene boxes from the grid, self partners from a padded cull
      MaskT ms = 0;
      unsigned mself = 0;
      if (!done && ((check_mask >> slot) & 1u)) {
        if (n_scene > 0) {
          const float fx = (cw[0] - grid.ox) * grid.inv_cell;
          const float fy = (cw[1] - grid.oy) * grid.inv_cell;
          const float fz = (cw[2] - grid.oz) * grid.inv_cell;
          const int ix = __float2int_rd(fx), iy = __float2int_rd(fy), iz = __float2int_rd(fz);
          const unsigned dim = (unsigned)grid.dim;
          if ((unsigned)ix < dim && (unsigned)iy < dim && (unsigned)iz < dim) {
            // < 2^32 cells by construction (24 classes x 160^3)
            const unsigned cell = ((((meta >> 4) & 31u) * dim + iz) * dim + iy) * dim + ix;
            ms = __ldg(reinterpret_cast<const MaskT*>(grid.cells) + cell);
          } else {
            ms = n_scene >= (int)(8 * sizeof(MaskT)) ? ~MaskT(0) : ((MaskT(1) << n_scene) - 1);
          }
        }
        if (do_self) {
          const int pair0 = (meta >> 16) & 255u, pair1 = pair0 + (int)(meta >> 24);
          for (int k = pair0; k < pair1; ++k) {
            const int sl = rb.partner_slot[k];
            // saved pose of the partner: 12 floats per lane as three 16-byte pieces (a lane
            // stride of 48 bytes is conflict free for 16-byte accesses)
            const float4* sv = reinterpret_cast<const float4*>(s_saved) + (sl * 32 + (int)lane) * 3;
            const float4 s0 = sv[0], s1 = sv[1], s2 = sv[2];  // R0 R3 R1 R4 | R2 R5 c0 c1 | R6 R7 R8 c2
            // partner's face axes against this box's bounding sphere, padded
            const float d0 = cw[0] - s1.z, d1 = cw[1] - s1.w, d2 = cw[2] - s2.w;
            const float ra = box_radius + delta;
            const float y0 = fmaf(s2.x, d2, fmaf(s0.y, d1, s0.x * d0));
            const float y1 = fmaf(s2.y, d2, fmaf(s0.w, d1, s0.z * d0));
            const float y2 = fmaf(s2.z, d2, fmaf(s1.y, d1, s1.x * d0));
            const bool sep = (fabsf(y0) > s_half[4 * sl + 0] + ra) |
                             (fabsf(y1) > s_half[4 * sl + 1] + ra) |
                             (fabsf(y2) > s_half[4 * sl + 2] + ra);
            if (!sep) mself |= 1u << sl;
          }
        }
      }

      // ---- rounds over the warp's candidate pairs.  A round is "direct" (every lane
      // tests its own next candidate with its own pose: no data movement) when most lanes
      // have one, or no lane has more than one; otherwise the remaining pairs of the whole
      // warp are redistributed over its lanes (items in a shared-memory list, the owner's
      // pose fetched with shuffles) so that the 15-axis test still runs on dense warps.
      auto load_other = [&](int other, int owner, float* R2, float* T2, float& B0, float& B1,
                            float& B2, float& rB) {
        if (other < 64) {
          const float* ob = s_scene + other * kSceneStrideF;
          R2[0] = ob[0]; R2[1] = ob[1]; R2[3] = ob[2]; R2[4] = ob[3];
          R2[6] = ob[4]; R2[7] = ob[5]; R2[2] = ob[6]; R2[5] = ob[7];
          R2[8] = ob[8];
          T2[0] = ob[9]; T2[1] = ob[10]; T2[2] = ob[11];
          B0 = ob[12]; B1 = ob[13]; B2 = ob[14];
          rB = ob[15];
        } else {
          const int sl = other - 64;
          const float4* sv = reinterpret_cast<const float4*>(s_saved) + (sl * 32 + owner) * 3;
          const float4 s0 = sv[0], s1 = sv[1], s2 = sv[2];
          R2[0] = s0.x; R2[3] = s0.y; R2[1] = s0.z; R2[4] = s0.w;
          R2[2] = s1.x; R2[5] = s1.y; T2[0] = s1.z; T2[1] = s1.w;
          R2[6] = s2.x; R2[7] = s2.y; R2[8] = s2.z; T2[2] = s2.w;
          B0 = s_half[4 * sl]; B1 = s_half[4 * sl + 1]; B2 = s_half[4 * sl + 2];
          rB = s_half[4 * sl + 3];
        }
      };
      for (;;) {
        const int c = done ? 0 : popc_t<MaskT>(ms) + __popc(mself);
        const unsigned b1 = __ballot_sync(kFull, c > 0);
        if (b1 == 0u) break;
        const unsigned b2 = __ballot_sync(kFull, c > 1);
        if (b2 == 0u || __popc(b1) >= ACRO_FILTER_DIRECT_MIN) {
          if (c > 0) {
            int other;
            if (ms != 0) {
              other = ffs_t<MaskT>(ms);
              ms &= ms - 1;
            } else {
              other = 64 + __ffs((int)mself) - 1;
              mself &= mself - 1;
            }
            float R2[9], T2[3], B0, B1, B2, rB;
            load_other(other, (int)lane, R2, T2, B0, B1, B2, rB);
            const int res = sat_filter(A0, A1, A2, Rw, cw, R2, T2, B0, B1, B2, delta, box_radius, rB);
            if (res == kHit) {
              hit = true;
              done = true;
            } else if (res == kAmb) {
              amb = true;
            }
          }
          __syncwarp();
          if (b2 == 0u) break;  // nobody had a second candidate: this box is finished
          continue;
        }
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(kFull, incl, d);
          if ((int)lane >= d) incl += v;
        }
        const int total = __shfl_sync(kFull, incl, 31);
        int pos = incl - c;
        if (!done) {
          while (pos < kItemCap && ms != 0) {
            const int j = ffs_t<MaskT>(ms);
            ms &= ms - 1;
            list[pos++] = (uint16_t)((lane << 8) | j);
          }
          while (pos < kItemCap && mself != 0) {
            const int j = __ffs((int)mself) - 1;
            mself &= mself - 1;
            list[pos++] = (uint16_t)((lane << 8) | (64 + j));
          }
        }
        __syncwarp();
        const int cnt = total < kItemCap ? total : kItemCap;
        for (int base = 0; base < cnt; base += 32) {
          const int k = base + (int)lane;
          const bool valid = k < cnt;
          const unsigned item = valid ? list[k] : (lane << 8);
          const int owner = item >> 8, other = item & 255;
          float R1[9], c1[3];
#pragma unroll
          for (int e = 0; e < 9; ++e) R1[e] = __shfl_sync(kFull, Rw[e], owner);
#pragma unroll
          for (int e = 0; e < 3; ++e) c1[e] = __shfl_sync(kFull, cw[e], owner);
          const bool owner_done = __shfl_sync(kFull, (int)done, owner) != 0;
          int res = kSep;
          if (valid && !owner_done) {
            float R2[9], T2[3], B0, B1, B2, rB;
            load_other(other, owner, R2, T2, B0, B1, B2, rB);
            res = sat_filter(A0, A1, A2, R1, c1, R2, T2, B0, B1, B2, delta, box_radius, rB);
          }
          const unsigned hb = __reduce_or_sync(kFull, res == kHit ? (1u << owner) : 0u);
          const unsigned ab = __reduce_or_sync(kFull, res == kAmb ? (1u << owner) : 0u);
          if ((hb >> lane) & 1u) {
            hit = true;
            done = true;
          }
          if ((ab >> lane) & 1u) amb = true;
        }
        __syncwarp();
      }

      const int save_slot = (int)((meta >> 9) & 31u) - 1;
      if (do_self && save_slot >= 0) {
        // order: (R0,R3) (R1,R4) (R2,R5) (c0,c1) (R6,R7) (R8,c2) - the register pairs the packed
        // DH step and pose leave behind, stored as they are
        float2* sv = reinterpret_cast<float2*>(s_saved) + (save_slot * 32 + (int)lane) * 6;
        sv[0] = make_float2(Rw[0], Rw[3]);
        sv[1] = make_float2(Rw[1], Rw[4]);
        sv[2] = make_float2(Rw[2], Rw[5]);
        sv[3] = make_float2(cw[0], cw[1]);
        sv[4] = make_float2(Rw[6], Rw[7]);
        sv[5] = make_float2(Rw[8], cw[2]);
      }
    }
  }
  retired = done;
}

}  // namespace acro
