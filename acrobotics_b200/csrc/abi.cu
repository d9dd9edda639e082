// abi.cu - the extern "C" boundary of libacro_b200.so (include/acro_b200.h):
// handle management, argument validation, error reporting, the host-buffer
// pipeline and the FMA-peak micro-benchmark.  No kernels of the hot path live
// here; see fk.cu, fk_cc.cu, jac_ik.cu, path.cu.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "acro_launch.h"

using namespace acro;

struct AcroRobot {
  AcroRobotDesc desc;
  RobotDev<double> d64;
  RobotDev<float> d32;
  // filtered K2: distinct bounding radii of the boxes, and the cube the boxes can reach
  int n_classes;
  double class_radius[ACRO_MAX_ROBOT_BOXES];
  double centre[3];   // base position
  double reach;       // bound on |box centre - base position| + box radius
  double q_lin_max;   // prismatic travel covered by `reach`
};

// Candidate grid + filter tolerance of one (robot geometry, scene) pair; see GridView.
struct CullPlan {
  int n_classes;
  double class_radius[ACRO_MAX_ROBOT_BOXES];
  double centre[3], reach;
  int dim;
  int mask_bits;
  void* cells;
  GridView view;
  cudaEvent_t ready;   // recorded after the grid build; consumers on other streams wait on it
  int device;          // the grid lives in this device's memory
  int pins;            // host threads between get_plan and the end of their launches
  uint64_t last_use;   // for eviction (a scene keeps at most kMaxPlans plans)
};

constexpr size_t kMaxPlans = 8;

struct AcroScene {
  int32_t n;
  int device;
  double* dev64;
  float* dev32;
  CullTable cull;  // FP32 face-axis cull table (kernel parameter of the wavefront K2)
  double bound;    // max over boxes of |centre| + bounding radius
  double max_centre;    // max over boxes of |centre|
  double max_half_sum;  // max over boxes of half[0] + half[1] + half[2]
  // plans are created on first use (or by acro_scene_prepare), never freed before the scene
  mutable std::mutex mu;
  mutable std::vector<CullPlan*> plans;
};

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* msg) {
  g_err = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  return (int)e;
}

bool is_identity_rot(const double* m) {
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      if (m[4 * r + c] != I[3 * r + c]) return false;
  return true;
}

bool is_identity_tf(const double* m) {
  return is_identity_rot(m) && m[3] == 0.0 && m[7] == 0.0 && m[11] == 0.0;
}

template <typename real>
int build_dev(const AcroRobotDesc& d, RobotDev<real>& o) {
  std::memset(&o, 0, sizeof(o));
  const int n = d.ndof;
  o.ndof = n;
  o.has_tool_tip = d.has_tool_tip;
  o.check_self = d.check_self;
  o.n_boxes = d.n_boxes;
  o.base_identity = is_identity_tf(d.tf_base);
  for (int i = 0; i < n; ++i) {
    if (d.joint_type[i]) o.prismatic_mask |= 1 << i;
    o.a[i] = (real)d.a[i];
    o.d[i] = (real)d.d[i];
    o.theta[i] = (real)d.theta[i];
    o.ca[i] = (real)d.cos_alpha[i];
    o.sa[i] = (real)d.sin_alpha[i];
  }
  for (int k = 0; k < 12; ++k) {
    o.base[k] = (real)d.tf_base[k];
    o.tip[k] = (real)d.tf_tool_tip[k];
  }
  // processing order: base boxes, link 0..n-1 boxes, tool boxes
  auto slot_of = [&](int carrier) {
    if (carrier == ACRO_CARRIER_BASE) return 0;
    if (carrier == ACRO_CARRIER_TOOL) return n + 1;
    return carrier + 1;
  };
  std::vector<int> order;
  for (int s = 0; s <= n + 1; ++s) {
    o.slot_begin[s] = (int)order.size();
    for (int b = 0; b < d.n_boxes; ++b)
      if (slot_of(d.boxes[b].carrier) == s) order.push_back(b);
  }
  o.slot_begin[n + 2] = (int)order.size();
  std::vector<int> new_index(d.n_boxes, -1);
  for (int k = 0; k < (int)order.size(); ++k) new_index[order[k]] = k;
  for (int k = 0; k < (int)order.size(); ++k) {
    const AcroRobotBox& src = d.boxes[order[k]];
    BoxDev<real>& bx = o.boxes[k];
    double r2 = 0;
    for (int e = 0; e < 3; ++e) {
      bx.half[e] = (real)src.half[e];
      r2 += src.half[e] * src.half[e];
    }
    // round the bounding radius up so that the sphere really contains the box
    bx.radius = (real)(std::sqrt(r2) * (1.0 + 1e-6));
    for (int e = 0; e < 12; ++e) bx.off[e] = (real)src.tf[e];
    bx.rot_identity = is_identity_rot(src.tf);
    bx.save_slot = -1;
    bx.cull_class = 0;
    bx.reserved = 0;
  }
  // radius classes (one candidate grid per distinct bounding radius)
  {
    std::vector<double> radii;
    for (int k = 0; k < (int)order.size(); ++k) {
      const AcroRobotBox& src = d.boxes[order[k]];
      const double rad = std::sqrt(src.half[0] * src.half[0] + src.half[1] * src.half[1] +
                                   src.half[2] * src.half[2]);
      int cls = -1;
      for (int c = 0; c < (int)radii.size(); ++c)
        if (radii[c] == rad) cls = c;
      if (cls < 0) {
        cls = (int)radii.size();
        radii.push_back(rad);
      }
      o.boxes[k].cull_class = cls;
    }
  }
  // self pairs grouped by their later box (in processing order)
  std::vector<std::vector<int>> partners(d.n_boxes);
  int n_saved = 0;
  for (int p = 0; p < d.n_self_pairs; ++p) {
    const int b1 = new_index[d.self_pairs[p][0]], b2 = new_index[d.self_pairs[p][1]];
    const int lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1;
    partners[hi].push_back(lo);
    if (o.boxes[lo].save_slot < 0) {
      if (n_saved >= kMaxSaved) return -1;
      o.boxes[lo].save_slot = n_saved++;
    }
  }
  int cursor = 0;
  for (int b = 0; b < d.n_boxes; ++b) {
    o.pair_begin[b] = (uint8_t)cursor;
    for (int lo : partners[b]) o.partner[cursor++] = (uint8_t)lo;
  }
  for (int b = d.n_boxes; b <= kMaxBoxes; ++b) o.pair_begin[b] = (uint8_t)cursor;
  for (int i = 0; i < n; ++i)
    o.dh4[i] = make_float4((float)o.ca[i], (float)o.sa[i], (float)o.a[i], (float)o.d[i]);
  // compact per-box records of the FP32 filter
  for (int k = 0; k < cursor; ++k) o.partner_slot[k] = (uint8_t)o.boxes[o.partner[k]].save_slot;
  for (int s = 0; s <= n + 1; ++s) {
    for (int b = o.slot_begin[s]; b < o.slot_begin[s + 1]; ++b) {
      const BoxDev<real>& bx = o.boxes[b];
      BoxHot& h = o.hot[b];
      h.off[0] = (float)bx.off[3];
      h.off[1] = (float)bx.off[7];
      h.off[2] = (float)bx.off[11];
      h.radius = (float)bx.radius;
      for (int e = 0; e < 3; ++e) h.half[e] = (float)bx.half[e];
      h.meta = (uint32_t)s | ((uint32_t)bx.cull_class << 4) | ((uint32_t)(bx.save_slot + 1) << 9) |
               ((uint32_t)(bx.rot_identity != 0) << 14) | ((uint32_t)o.pair_begin[b] << 16) |
               ((uint32_t)(o.pair_begin[b + 1] - o.pair_begin[b]) << 24);
    }
  }
  return 0;
}

cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename real>
SceneView<real> view_of(const AcroScene* sc);
template <>
SceneView<double> view_of<double>(const AcroScene* sc) {
  return sc ? SceneView<double>{sc->dev64, sc->n} : SceneView<double>{nullptr, 0};
}
template <>
SceneView<float> view_of<float>(const AcroScene* sc) {
  return sc ? SceneView<float>{sc->dev32, sc->n} : SceneView<float>{nullptr, 0};
}

int check_batch(const void* robot, const void* in, const void* out, int64_t n) {
  if (!robot) return fail(ACRO_E_INVALID, "robot handle is NULL");
  if (n < 0) return fail(ACRO_E_INVALID, "negative batch size");
  if (n > 0 && (!in || !out)) return fail(ACRO_E_INVALID, "NULL buffer for a non-empty batch");
  return 0;
}

// a scene's boxes live in the memory of the device that was current at acro_scene_create
int check_scene_device(const AcroScene* scene) {
  if (!scene) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(ACRO_E_NOGPU, "no usable CUDA device");
  if (dev != scene->device)
    return fail(ACRO_E_INVALID, "scene handle was created on another CUDA device");
  return 0;
}

int finish(cudaError_t e, const char* where) { return e == cudaSuccess ? 0 : cuda_fail(e, where); }

float round_up_to_float(double x);

// ACRO_K2_VARIANT selects the kernel behind the plain FP64 K2 path (A/B measurements,
// cross-checks): "filter" (default, fk_cc_filt.cu), "wave" (fk_cc_wave.cu), "simple" (fk_cc.cu)
std::atomic<int> g_k2_variant{-1};
std::atomic<int> g_grid_dim{-1};

int parse_variant(const std::string& s) {
  if (s == "simple") return 2;
  if (s == "wave") return 0;
  if (s == "filter") return 3;
  return -1;
}

int k2_variant() {
  int v = g_k2_variant.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("ACRO_K2_VARIANT");
    v = e ? parse_variant(e) : 3;
    if (v < 0) v = 3;
    g_k2_variant.store(v);
  }
  return v;
}

std::atomic<int> g_split{-1};  // -1 auto, 0 off, k > 0: split after slot k

// Slot after which the filtered kernel hands undecided configurations to its second pass.
// Auto: after the second link when there is a scene and the chain is long enough for the tail
// to matter; a scene-less (self-collision only) call decides nothing early, so no split.
int k2_split_slot(const AcroRobot* robot, const AcroScene* scene) {
  int v = g_split.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("ACRO_K2_SPLIT");
    if (e) {
      v = std::atoi(e);
      g_split.store(v < 0 ? -1 : v);
    }
  }
  if (v < 0) v = (scene && scene->n > 0 && robot->d32.ndof >= 5) ? 2 : 0;
  if (v >= robot->d32.ndof) v = 0;
  return v;
}

int grid_dim_setting() {
  int d = g_grid_dim.load(std::memory_order_relaxed);
  if (d < 0) {
    const char* e = std::getenv("ACRO_K2_GRID");
    d = e ? std::atoi(e) : 128;
    d = d < 8 ? 8 : (d > 160 ? 160 : d);
    g_grid_dim.store(d);
  }
  return d;
}

// Find or build the candidate grid + filter tolerance for (robot geometry, scene).
// The returned plan is pinned: call release_plan once the kernels that use it are enqueued.
// (Eviction frees a plan's grid with cudaFree, which waits for kernels still running.)
void release_plan(const AcroScene* scene, const CullPlan* plan) {
  if (!plan) return;
  std::lock_guard<std::mutex> lock(scene->mu);
  --const_cast<CullPlan*>(plan)->pins;
}

struct PlanPin {  // RAII holder
  const AcroScene* scene = nullptr;
  const CullPlan* plan = nullptr;
  ~PlanPin() { release_plan(scene, plan); }
};

// ---------------------------------------------------------------------------
// FP32 filter tolerance: a running error bound, not a heuristic.
//
// delta must bound |s_fp32 - s_fp64| for every quantity s that FCL's test compares with zero,
// for every configuration the filter accepts (revolute |q| < 2^26, prismatic |q| <= q_lin_max),
// every robot box and every partner box (scene boxes; other robot boxes for self pairs).  The
// FP64 side is treated as exact (its own rounding, ~1e-15, is covered by the final slack).
// u = 2^-24 (FP32 unit roundoff; an FMA rounds once).  Norms: Frobenius for the 3x3 rotation
// error (it bounds every entry and every column), Euclidean for position errors.
//
//  sin/cos   |s32 - sin q| <= e_t = 1.4e-7: sincos_filter reduces q in FP64 (error <= 1.2e-8
//            at |q| = 2^26), rounds the reduced angle to FP32 and evaluates the Cephes
//            polynomials; measured over 8e7 arguments: 6.4e-8 (sin), 9.5e-8 (cos)
//            (tests/test_filter_tolerance_cpu.py repeats the measurement).
//  joint k   R_k = R_{k-1} D_k.  ||dD||_F <= 2 e_t + 2u (the four trigonometric entries of D
//            times cos/sin(alpha), cos^2 + sin^2 = 1; + the FP32 constants).  The factored
//            evaluation tf_mul_dh rounds each output column to within 2u, 4u, 4u (column
//            norms <= 1): 6u in Frobenius norm.  Rotations are isometries, so
//            rho_k <= rho_{k-1} (1 + ||dD||) + ||dD|| + 6u.
//            p_k = p_{k-1} + a u_k + d z_{k-1}:
//            pi_k <= pi_{k-1} + |a| (rho_k + u) + |d| (rho_{k-1} + u) + 2u |p_k|
//            (+ u |q| for a prismatic joint, whose travel is rounded to FP32).
//  box       offset = pure translation t: R_box = R_k, c = p + R t:
//            pi_b <= pi_k + |t| (rho_k + u) + 3u (|p| + |t|); general offset: one more product,
//            rho_b <= rho_k (1 + 3u) + 3u + 9u.
//  pair      p = T2 - c1, D >= |p|.  eps_p = pi_1 + pi_2 + u D.  Entries of R = R1^T R2:
//            eps_R = rho_1 + rho_2 + 3u.  Projections R^T p: eps_pp = max(rho) D + eps_p + 3u D.
//            face axis  |t| - (sum |R| A + B):  eps_pp + eps_R S + 5u (S + D)
//            edge axis  |pp R - pp R| - (A Q + A Q + B Q + B Q), |R| <= 1:
//                       2 (D eps_R + eps_pp) + 2u D + eps_R S + 6u S
//            with S = the largest sum of half extents entering one radius.
// A scene box has rho_2 = 3u, pi_2 = sqrt(3) u |T2| (its FP32 copy).  The bound is evaluated
// with D = the largest possible centre distance, i.e. it holds for every pair, near or far.
// ---------------------------------------------------------------------------
double derive_filter_tolerance(const AcroRobot* robot, const AcroScene* scene) {
  const double u = std::ldexp(1.0, -24), e_t = 1.4e-7;
  const AcroRobotDesc& d = robot->desc;
  const int n = d.ndof;
  auto norm3 = [](double x, double y, double z) { return std::sqrt(x * x + y * y + z * z); };
  double rho = 3.0 * u;  // tf_base rounded to FP32
  double pnorm = norm3(d.tf_base[3], d.tf_base[7], d.tf_base[11]);
  double pi_ = std::sqrt(3.0) * u * pnorm;
  double rho_box = 0.0, pi_box = 0.0, reach_box = 0.0, half_sum = 0.0;
  auto visit_boxes = [&](int carrier) {
    for (int b = 0; b < d.n_boxes; ++b) {
      if (d.boxes[b].carrier != carrier) continue;
      const double* t = d.boxes[b].tf;
      const double tn = norm3(t[3], t[7], t[11]);
      const bool ident = is_identity_rot(t);
      const double rb = ident ? rho : rho * (1.0 + 3.0 * u) + 12.0 * u;
      const double pb = pi_ + tn * (rho + u) + 3.0 * u * (pnorm + tn);
      rho_box = std::max(rho_box, rb);
      pi_box = std::max(pi_box, pb);
      reach_box = std::max(reach_box, pnorm + tn);
      half_sum = std::max(half_sum, d.boxes[b].half[0] + d.boxes[b].half[1] + d.boxes[b].half[2]);
    }
  };
  visit_boxes(ACRO_CARRIER_BASE);
  const double dD = 2.0 * e_t + 2.0 * u;
  for (int k = 0; k < n; ++k) {
    const double rho_prev = rho;
    rho = rho_prev * (1.0 + dD) + dD + 6.0 * u;
    const double a = std::fabs(d.a[k]);
    const double len = d.joint_type[k] ? robot->q_lin_max : std::fabs(d.d[k]);
    pnorm += a + len;
    pi_ += a * (rho + u) + len * (rho_prev + u) + 2.0 * u * pnorm + (d.joint_type[k] ? u * len : 0.0);
    visit_boxes(k);
    if (k == n - 1) visit_boxes(ACRO_CARRIER_TOOL);
  }
  // partner: scene box or (self pairs) another robot box
  const double sc_centre = scene ? scene->max_centre : 0.0;
  const double sc_half = scene ? scene->max_half_sum : 0.0;
  const double rho2 = std::max(3.0 * u, robot->desc.n_self_pairs > 0 ? rho_box : 0.0);
  const double pi2 = std::max(std::sqrt(3.0) * u * sc_centre, robot->desc.n_self_pairs > 0 ? pi_box : 0.0);
  const double D = reach_box + std::max(sc_centre, reach_box);
  const double S = half_sum + std::max(sc_half, half_sum);
  const double eps_p = pi_box + pi2 + u * D;
  const double eps_R = rho_box + rho2 + 3.0 * u;
  const double eps_pp = std::max(rho_box, rho2) * D + eps_p + 3.0 * u * D;
  const double face = eps_pp + eps_R * S + 5.0 * u * (S + D);
  const double edge = 2.0 * (D * eps_R + eps_pp) + 2.0 * u * D + eps_R * S + 6.0 * u * S;
  return 1.05 * std::max(face, edge) + 1.0e-9;
}

// create = false: only look an existing plan up (*out stays nullptr when there is none)
// A new plan is built on stream `st` (no device-wide synchronisation); every consumer makes its
// stream wait on the plan's `ready` event, so a plan found by another thread / stream while its
// grid is still being filled is safe to use.
cudaError_t get_plan(const AcroRobot* robot, const AcroScene* scene, const CullPlan** out,
                     bool create = true, cudaStream_t st = nullptr) {
  std::lock_guard<std::mutex> lock(scene->mu);
  static std::atomic<uint64_t> clock{0};
  int dev = 0;
  cudaError_t de = cudaGetDevice(&dev);
  if (de != cudaSuccess) return de;
  for (CullPlan* p : scene->plans) {
    if (p->device != dev) continue;
    if (p->n_classes != robot->n_classes || p->reach != robot->reach ||
        p->dim != grid_dim_setting())
      continue;
    bool same = p->centre[0] == robot->centre[0] && p->centre[1] == robot->centre[1] &&
                p->centre[2] == robot->centre[2];
    for (int c = 0; same && c < p->n_classes; ++c)
      same = p->class_radius[c] == robot->class_radius[c];
    if (same) {
      ++p->pins;
      p->last_use = ++clock;
      *out = p;
      return p->ready ? cudaStreamWaitEvent(st, p->ready, 0) : cudaSuccess;
    }
  }
  if (!create) {
    *out = nullptr;
    return cudaSuccess;
  }
  // a robot whose base keeps moving through one scene would otherwise pile up grids
  while (scene->plans.size() >= kMaxPlans) {
    size_t victim = scene->plans.size();
    for (size_t k = 0; k < scene->plans.size(); ++k)
      if (scene->plans[k]->pins == 0 &&
          (victim == scene->plans.size() || scene->plans[k]->last_use < scene->plans[victim]->last_use))
        victim = k;
    if (victim == scene->plans.size()) break;  // everything is in use: grow instead
    cudaFree(scene->plans[victim]->cells);
    if (scene->plans[victim]->ready) cudaEventDestroy(scene->plans[victim]->ready);
    delete scene->plans[victim];
    scene->plans.erase(scene->plans.begin() + victim);
  }
  CullPlan* p = new CullPlan();
  p->pins = 1;
  p->device = dev;
  p->last_use = ++clock;
  p->n_classes = robot->n_classes;
  for (int c = 0; c < p->n_classes; ++c) p->class_radius[c] = robot->class_radius[c];
  for (int k = 0; k < 3; ++k) p->centre[k] = robot->centre[k];
  p->reach = robot->reach;
  p->dim = grid_dim_setting();
  p->mask_bits = scene->n <= 32 ? 32 : 64;
  p->cells = nullptr;
  p->ready = nullptr;
  // FP32 filter tolerance: the running error bound of derive_filter_tolerance (measured over
  // 4 M configurations x 96 pairs x 15 axes with acro_measure_filter_error: 1.5e-6 for the
  // Kuka and the 16-box scene, for which the bound is ~1.1e-4)
  const double delta = derive_filter_tolerance(robot, scene);
  const double half = std::max(robot->reach, 1e-3);
  const double cell = 2.0 * half / p->dim;
  const double origin[3] = {p->centre[0] - half, p->centre[1] - half, p->centre[2] - half};
  p->view.cells = nullptr;
  p->view.ox = (float)origin[0];
  p->view.oy = (float)origin[1];
  p->view.oz = (float)origin[2];
  p->view.inv_cell = (float)(1.0 / cell);
  p->view.dim = p->dim;
  p->view.delta = round_up_to_float(delta);
  p->view.q_lin_max = (float)robot->q_lin_max;
  p->view.skip_nan = 0;
  cudaError_t e = cudaSuccess;
  if (scene->n > 0 && p->n_classes > 0) {
    const size_t cells = (size_t)p->dim * p->dim * p->dim * p->n_classes;
    // one allocation per (robot geometry, scene) pair, on first use or in acro_scene_prepare;
    // the class radii travel as a kernel parameter, the build is ordered on `st`
    e = cudaMalloc(&p->cells, cells * (p->mask_bits / 8));
    // pad: half diagonal of a cell + FP32 position error + index rounding slack
    const double pad = 0.5 * cell * std::sqrt(3.0) * (1.0 + 1e-6) + delta + 1e-3 * cell;
    if (e == cudaSuccess)
      e = launch_build_grid(SceneView<double>{scene->dev64, scene->n}, p->cells, p->mask_bits,
                            p->dim, p->n_classes, origin, cell, p->class_radius, pad, st);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(p->ready, st);
    if (e != cudaSuccess) {
      cudaFree(p->cells);
      if (p->ready) cudaEventDestroy(p->ready);
      delete p;
      return e;
    }
    p->view.cells = p->cells;
  }
  scene->plans.push_back(p);
  *out = p;
  return cudaSuccess;
}

float round_up_to_float(double x) {
  float f = (float)x;
  if ((double)f < x) f = std::nextafterf(f, INFINITY);
  return f;
}

// see CullTable in acro_device.cuh
void build_cull_table(const double* packed, int n, CullTable& ct) {
  std::memset(&ct, 0, sizeof(ct));
  for (int j = 0; j < n; ++j) {
    const double* b = packed + (size_t)j * kSceneStride;
    double t[3];
    for (int k = 0; k < 3; ++k)
      t[k] = -(b[k] * b[9] + b[3 + k] * b[10] + b[6 + k] * b[11]);
    const double tn = std::fabs(t[0]) + std::fabs(t[1]) + std::fabs(t[2]);
    float* e = ct.e[j];
    for (int k = 0; k < 3; ++k) {
      e[4 * k + 0] = (float)b[k];
      e[4 * k + 1] = (float)b[3 + k];
      e[4 * k + 2] = (float)b[6 + k];
      e[4 * k + 3] = (float)t[k];
      e[12 + k] = round_up_to_float(b[12 + k] * (1.0 + 1e-6) + 1e-6 * tn);
    }
  }
}

constexpr int64_t kFilterMinBatch = 4096;  // below this the plain FP64 kernel is used
// building a candidate grid (~1 ms) only pays for batches of this size; smaller ones take the
// filtered path when the (robot, scene) pair already has a plan (acro_scene_prepare, or an
// earlier large batch)
constexpr int64_t kPlanMinBatch = 1 << 16;

// near_mask != nullptr: margin mode (verdicts + the configurations within +-margin of contact)
cudaError_t run_fk_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* q,
                          int64_t n, int layout, uint32_t* mask, cudaStream_t st,
                          bool skip_nan = false, uint32_t* near_mask = nullptr,
                          double margin = 0.0, void* workspace = nullptr,
                          size_t workspace_bytes = 0) {
  const int variant = k2_variant();
  bool filtered = (variant >= 3 || skip_nan) && n >= kFilterMinBatch;
  const CullPlan* plan = nullptr;
  static const AcroScene empty_scene{};
  PlanPin pin;
  if (filtered) {
    cudaError_t e = get_plan(robot, scene ? scene : &empty_scene, &plan, n >= kPlanMinBatch, st);
    if (e != cudaSuccess) return e;
    if (!plan) filtered = false;
    pin.scene = scene ? scene : &empty_scene;
    pin.plan = plan;
    // the grid's cull is only good for margins well inside the filter tolerance
    if (plan && near_mask && !(margin <= 0.5 * (double)plan->view.delta)) filtered = false;
  }
  if (filtered) {
    cudaError_t e = cudaSuccess;
    // stream-ordered scratch for the ambiguity mask (one bit per configuration)
    static std::once_flag pool_once;
    std::call_once(pool_once, [] {
      int dev = 0;
      cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess &&
          cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
    });
    // scratch (ambiguity mask + chunk counter): the caller's workspace when it is large enough
    // (acro_fk_cc_f64_ws: no allocation in the hot call), else stream-ordered pool memory
    uint32_t* amb = nullptr;
    const bool own_scratch = !(workspace && workspace_bytes >= filter_scratch_bytes(n) &&
                               (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0);
    if (own_scratch) {
      e = cudaMallocAsync(&amb, filter_scratch_bytes(n), st);
      if (e != cudaSuccess) return e;
    } else {
      amb = static_cast<uint32_t*>(workspace);
    }
    GridView view = plan->view;
    view.skip_nan = skip_nan ? 1 : 0;
    if (near_mask) {
      view.delta = round_up_to_float((double)view.delta + margin);
      e = cudaMemsetAsync(near_mask, 0, sizeof(uint32_t) * ((n + 31) / 32), st);
    }
    if (e == cudaSuccess)
      e = launch_fk_cc_filter(robot->d32, robot->d64, view, plan->mask_bits,
                              view_of<float>(scene), view_of<double>(scene), q, n, layout, mask,
                              amb, near_mask, margin, k2_split_slot(robot, scene), st);
    cudaError_t e2 = own_scratch ? cudaFreeAsync(amb, st) : cudaSuccess;
    return e != cudaSuccess ? e : e2;
  }
  if (near_mask || variant >= 2 || (variant == 0 && !wave_supports(robot->d64)))
    return launch_fk_cc<double>(robot->d64, view_of<double>(scene), q, n, layout, mask, near_mask,
                                margin, st);
  static const CullTable empty_table = {};
  return launch_fk_cc_wave(robot->d64, scene ? scene->cull : empty_table, view_of<double>(scene),
                           q, n, layout, mask, st);
}

}  // namespace

namespace acro {
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace acro

extern "C" {

int acro_abi_version(void) { return ACRO_ABI_VERSION; }

const char* acro_last_error(void) { return g_err.c_str(); }

int64_t acro_launch_count(void) { return g_launches.load(); }

int acro_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
    return -1;
  }
  return n;
}

// K7: batches of at least this many pairs go to the persistent kernel
static constexpr int64_t kSweptPersistentMin = 4096;
static std::atomic<int> g_swept_persistent{0};
static bool swept_persistent() { return g_swept_persistent.load() != 0; }

int acro_set_option(const char* name, const char* value) {
  if (!name || !value) return fail(ACRO_E_INVALID, "NULL option");
  const std::string n(name), v(value);
  if (n == "k2_variant") {
    const int k = parse_variant(v);
    if (k < 0) return fail(ACRO_E_INVALID, "k2_variant: filter | wave | simple");
    g_k2_variant.store(k);
    return 0;
  }
  if (n == "k2_split") {
    const int k = v == "auto" ? -1 : std::atoi(value);
    if (k < -1 || k > ACRO_MAX_DOF) return fail(ACRO_E_INVALID, "k2_split: auto | 0 .. ndof-1");
    g_split.store(k);
    return 0;
  }
  if (n == "k2_grid") {
    const int d = std::atoi(value);
    if (d < 8 || d > 160) return fail(ACRO_E_INVALID, "k2_grid: 8..160 cells per axis");
    g_grid_dim.store(d);
    return 0;
  }
  if (n == "swept_persistent") {
    g_swept_persistent.store(std::atoi(value) != 0 ? 1 : 0);
    return 0;
  }
  return fail(ACRO_E_INVALID, "unknown option");
}

int acro_robot_create(const AcroRobotDesc* desc, AcroRobot** out) {
  if (!desc || !out) return fail(ACRO_E_INVALID, "NULL argument");
  *out = nullptr;
  if (desc->ndof < 1 || desc->ndof > ACRO_MAX_DOF)
    return fail(ACRO_E_UNSUPPORTED, "ndof outside 1..ACRO_MAX_DOF");
  if (desc->n_boxes < 0 || desc->n_boxes > ACRO_MAX_ROBOT_BOXES)
    return fail(ACRO_E_UNSUPPORTED, "too many robot boxes");
  if (desc->n_self_pairs < 0 || desc->n_self_pairs > ACRO_MAX_SELF_PAIRS)
    return fail(ACRO_E_UNSUPPORTED, "too many self-collision pairs");
  for (int i = 0; i < desc->ndof; ++i)
    if (desc->joint_type[i] != 0 && desc->joint_type[i] != 1)
      return fail(ACRO_E_INVALID, "unknown joint type");  // link.py:26-29 ValueError
  for (int b = 0; b < desc->n_boxes; ++b) {
    const int c = desc->boxes[b].carrier;
    if (!(c == ACRO_CARRIER_BASE || c == ACRO_CARRIER_TOOL || (c >= 0 && c < desc->ndof)))
      return fail(ACRO_E_INVALID, "box carrier out of range");
  }
  for (int p = 0; p < desc->n_self_pairs; ++p) {
    const int b1 = desc->self_pairs[p][0], b2 = desc->self_pairs[p][1];
    if (b1 < 0 || b2 < 0 || b1 >= desc->n_boxes || b2 >= desc->n_boxes || b1 == b2)
      return fail(ACRO_E_INVALID, "self pair index out of range");
  }
  AcroRobot* r = new AcroRobot;
  r->desc = *desc;
  if (build_dev<double>(*desc, r->d64) || build_dev<float>(*desc, r->d32)) {
    delete r;
    return fail(ACRO_E_UNSUPPORTED, "too many boxes take part in self-collision pairs");
  }
  // geometry summary for the filtered K2 kernel
  r->n_classes = 0;
  double box_reach = 0.0;
  for (int b = 0; b < r->d64.n_boxes; ++b) {
    const BoxDev<double>& bx = r->d64.boxes[b];
    if (bx.cull_class + 1 > r->n_classes) r->n_classes = bx.cull_class + 1;
    r->class_radius[bx.cull_class] = bx.radius;  // already rounded up
    const double off = std::sqrt(bx.off[3] * bx.off[3] + bx.off[7] * bx.off[7] + bx.off[11] * bx.off[11]);
    box_reach = std::max(box_reach, off + bx.radius);
  }
  r->q_lin_max = 4.0;
  double chain = 0.0;
  for (int i = 0; i < desc->ndof; ++i)
    chain += std::fabs(desc->a[i]) + (desc->joint_type[i] ? r->q_lin_max : std::fabs(desc->d[i]));
  r->reach = chain + box_reach;
  r->centre[0] = desc->tf_base[3];
  r->centre[1] = desc->tf_base[7];
  r->centre[2] = desc->tf_base[11];
  *out = r;
  return 0;
}

void acro_robot_destroy(AcroRobot* robot) { delete robot; }

int acro_scene_create(const AcroBox* boxes, int32_t n, AcroScene** out) {
  if (!out || n < 0 || (n > 0 && !boxes)) return fail(ACRO_E_INVALID, "bad scene arguments");
  *out = nullptr;
  if (n > ACRO_MAX_SCENE_BOXES) return fail(ACRO_E_UNSUPPORTED, "too many scene boxes");
  std::vector<double> p64((size_t)(n > 0 ? n : 1) * kSceneStride, 0.0);
  double bound = 0.0, max_centre = 0.0, max_half_sum = 0.0;
  for (int j = 0; j < n; ++j) {
    double* o = &p64[(size_t)j * kSceneStride];
    double r2 = 0, hmax = 0;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) o[3 * r + c] = boxes[j].tf[4 * r + c];
      o[9 + r] = boxes[j].tf[4 * r + 3];
      o[12 + r] = boxes[j].half[r];
      r2 += boxes[j].half[r] * boxes[j].half[r];
      hmax = boxes[j].half[r] > hmax ? boxes[j].half[r] : hmax;
    }
    o[15] = std::sqrt(r2) * (1.0 + 1e-6);
    o[16] = hmax;
    const double cn = std::sqrt(o[9] * o[9] + o[10] * o[10] + o[11] * o[11]);
    bound = std::max(bound, cn + o[15]);
    max_centre = std::max(max_centre, cn);
    max_half_sum = std::max(max_half_sum, o[12] + o[13] + o[14]);
  }
  std::vector<float> p32(p64.size());
  for (size_t k = 0; k < p64.size(); ++k) p32[k] = (float)p64[k];
  AcroScene* s = new AcroScene;
  s->n = n;
  s->bound = bound;
  s->max_centre = max_centre;
  s->max_half_sum = max_half_sum;
  s->device = 0;
  s->dev64 = nullptr;
  s->dev32 = nullptr;
  build_cull_table(p64.data(), n, s->cull);
  cudaError_t e = cudaGetDevice(&s->device);
  if (e == cudaSuccess) e = cudaMalloc(&s->dev64, p64.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&s->dev32, p32.size() * sizeof(float));
  if (e == cudaSuccess)
    e = cudaMemcpy(s->dev64, p64.data(), p64.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(s->dev32, p32.data(), p32.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(s->dev64);
    cudaFree(s->dev32);
    delete s;
    cudaGetLastError();
    return cuda_fail(e, "acro_scene_create");
  }
  *out = s;
  return 0;
}

void acro_scene_destroy(AcroScene* scene) {
  if (!scene) return;
  for (CullPlan* p : scene->plans) {
    cudaFree(p->cells);
    if (p->ready) cudaEventDestroy(p->ready);
    delete p;
  }
  cudaFree(scene->dev64);
  cudaFree(scene->dev32);
  delete scene;
}

int acro_scene_prepare(const AcroScene* scene, const AcroRobot* robot) {
  if (!scene || !robot) return fail(ACRO_E_INVALID, "NULL handle");
  if (int rc = check_scene_device(scene)) return rc;
  const CullPlan* plan = nullptr;
  cudaError_t e = get_plan(robot, scene, &plan, true, nullptr);
  if (e == cudaSuccess) {
    if (plan->ready) e = cudaEventSynchronize(plan->ready);  // the explicit call may wait
    release_plan(scene, plan);
  }
  return finish(e, "acro_scene_prepare");
}

int acro_collide_boxes_f64(const AcroBox* a, const AcroBox* b, int64_t n, uint32_t* mask,
                           void* stream) {
  if (n < 0 || (n > 0 && (!a || !b || !mask))) return fail(ACRO_E_INVALID, "bad arguments");
  return finish(launch_collide_boxes(a, b, n, mask, as_stream(stream)),
                "acro_collide_boxes_f64");
}

int acro_fk_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                double* T_out, int32_t all_links, void* stream) {
  if (int rc = check_batch(robot, q, T_out, n)) return rc;
  return finish(launch_fk<double>(robot->d64, q, n, q_layout, T_out, all_links != 0,
                                  as_stream(stream)),
                "acro_fk_f64");
}

int acro_fk_f32(const AcroRobot* robot, const float* q, int64_t n, int32_t q_layout, float* T_out,
                int32_t all_links, void* stream) {
  if (int rc = check_batch(robot, q, T_out, n)) return rc;
  return finish(
      launch_fk<float>(robot->d32, q, n, q_layout, T_out, all_links != 0, as_stream(stream)),
      "acro_fk_f32");
}

int acro_fk_rpy_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                    double* out, void* stream) {
  if (int rc = check_batch(robot, q, out, n)) return rc;
  return finish(launch_fk_rpy(robot->d64, q, n, q_layout, out, as_stream(stream)),
                "acro_fk_rpy_f64");
}

int acro_fk_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* q, int64_t n,
                   int32_t q_layout, uint32_t* mask, uint32_t* near_mask, double margin,
                   void* stream) {
  if (int rc = check_batch(robot, q, mask, n)) return rc;
  if (int rc = check_scene_device(scene)) return rc;
  if (near_mask && !(margin >= 0.0)) return fail(ACRO_E_INVALID, "margin must be >= 0");
  return finish(run_fk_cc_f64(robot, scene, q, n, q_layout, mask, as_stream(stream), false,
                              near_mask, margin),
                "acro_fk_cc_f64");
}

int64_t acro_fk_cc_workspace_bytes(int64_t n) { return n > 0 ? (int64_t)filter_scratch_bytes(n) : 0; }

int acro_fk_cc_f64_ws(const AcroRobot* robot, const AcroScene* scene, const double* q, int64_t n,
                      int32_t q_layout, uint32_t* mask, uint32_t* near_mask, double margin,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = check_batch(robot, q, mask, n)) return rc;
  if (int rc = check_scene_device(scene)) return rc;
  if (near_mask && !(margin >= 0.0)) return fail(ACRO_E_INVALID, "margin must be >= 0");
  if (workspace && workspace_bytes < acro_fk_cc_workspace_bytes(n))
    return fail(ACRO_E_INVALID, "workspace smaller than acro_fk_cc_workspace_bytes(n)");
  if (workspace && (reinterpret_cast<uintptr_t>(workspace) & 7u))
    return fail(ACRO_E_INVALID, "workspace must be 8-byte aligned");
  return finish(run_fk_cc_f64(robot, scene, q, n, q_layout, mask, as_stream(stream), false,
                              near_mask, margin, workspace, (size_t)(workspace ? workspace_bytes : 0)),
                "acro_fk_cc_f64_ws");
}

int acro_fk_cc_f32(const AcroRobot* robot, const AcroScene* scene, const float* q, int64_t n,
                   int32_t q_layout, uint32_t* mask, void* stream) {
  if (int rc = check_batch(robot, q, mask, n)) return rc;
  if (int rc = check_scene_device(scene)) return rc;
  cudaStream_t st = as_stream(stream);
  if (k2_variant() < 3 || n < kFilterMinBatch || q_layout != ACRO_Q_AOS)
    return finish(launch_fk_cc<float>(robot->d32, view_of<float>(scene), q, n, q_layout, mask,
                                      nullptr, 0.f, st),
                  "acro_fk_cc_f32");
  // large batches: the joint values are widened (exactly) chunk by chunk and go through the
  // filtered FP64-exact path - faster than the all-FP32 kernel and without its rounding
  const int ndof = robot->d64.ndof;
  const int64_t chunk = 1 << 24;  // multiple of 32: chunk boundaries fall on mask words
  double* tmp = nullptr;
  cudaError_t e = cudaMallocAsync(&tmp, sizeof(double) * std::min<int64_t>(n, chunk) * ndof, st);
  for (int64_t done = 0; done < n && e == cudaSuccess; done += chunk) {
    const int64_t m = std::min<int64_t>(chunk, n - done);
    e = launch_widen(q + done * ndof, tmp, m * ndof, st);
    if (e == cudaSuccess)
      e = run_fk_cc_f64(robot, scene, tmp, m, ACRO_Q_AOS, mask + done / 32, st);
  }
  if (tmp) {
    cudaError_t e2 = cudaFreeAsync(tmp, st);
    if (e == cudaSuccess) e = e2;
  }
  return finish(e, "acro_fk_cc_f32");
}

// ---------------------------------------------------------------------------
// host-buffer pipeline: chunked H2D -> K2 -> D2H over kSlots rotating slots
// ---------------------------------------------------------------------------
namespace {
constexpr int kSlots = 3;
struct HostPipe {
  std::mutex mu;
  int device = -1;
  int64_t chunk = 0;
  int ndof = 0;
  double* d_q[kSlots] = {};
  float* d_q32[kSlots] = {};  // staging of FP32 joint rows (acro_fk_cc_f32_host), on first use
  uint32_t* d_mask[kSlots] = {};
  cudaStream_t st[kSlots] = {};
  void release() {
    for (int s = 0; s < kSlots; ++s) {
      if (d_q[s]) cudaFree(d_q[s]);
      if (d_q32[s]) cudaFree(d_q32[s]);
      if (d_mask[s]) cudaFree(d_mask[s]);
      if (st[s]) cudaStreamDestroy(st[s]);
      d_q[s] = nullptr;
      d_q32[s] = nullptr;
      d_mask[s] = nullptr;
      st[s] = nullptr;
    }
    chunk = 0;
  }
};
// One pipeline per CUDA device (streams and staging buffers belong to a device): calls on
// different devices run concurrently from different host threads; calls on the same device
// share its copy engines anyway and take turns.
constexpr int kMaxPipeDevices = 64;
HostPipe g_pipes[kMaxPipeDevices];
}  // namespace

// q_host: doubles (f32 == false) or floats (f32 == true; widened exactly on the device, so the
// verdicts are those of the FP64 evaluation at the float-valued joint angles)
static int host_pipeline(const AcroRobot* robot, const AcroScene* scene, const void* q_host,
                         bool f32, int64_t n, uint32_t* mask_host, int64_t chunk, const char* what) {
  if (int rc = check_batch(robot, q_host, mask_host, n)) return rc;
  if (int rc = check_scene_device(scene)) return rc;
  if (n == 0) return 0;
  if (chunk <= 0) chunk = 1 << 22;
  chunk = (chunk + 31) / 32 * 32;  // chunk boundaries fall on mask words
  if (chunk > n) chunk = (n + 31) / 32 * 32;
  const int ndof = robot->d64.ndof;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= kMaxPipeDevices) return fail(ACRO_E_UNSUPPORTED, "device ordinal too large");
  HostPipe& g_pipe = g_pipes[dev];
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  if (g_pipe.device != dev || g_pipe.chunk < chunk || g_pipe.ndof < ndof) {
    g_pipe.release();
    for (int s = 0; s < kSlots && e == cudaSuccess; ++s) {
      e = cudaMalloc(&g_pipe.d_q[s], sizeof(double) * chunk * ndof);
      if (e == cudaSuccess) e = cudaMalloc(&g_pipe.d_mask[s], sizeof(uint32_t) * (chunk / 32));
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_pipe.st[s], cudaStreamNonBlocking);
    }
    if (e != cudaSuccess) {
      g_pipe.release();
      cudaGetLastError();
      return cuda_fail(e, "host pipeline workspace");
    }
    g_pipe.device = dev;
    g_pipe.chunk = chunk;
    g_pipe.ndof = ndof;
  }
  for (int s = 0; s < kSlots && f32 && e == cudaSuccess; ++s)
    if (!g_pipe.d_q32[s]) e = cudaMalloc(&g_pipe.d_q32[s], sizeof(float) * g_pipe.chunk * g_pipe.ndof);
  if (e != cudaSuccess) return cuda_fail(e, "host pipeline workspace");
  int64_t done = 0;
  int slot = 0;
  while (done < n && e == cudaSuccess) {
    const int64_t m = (n - done) < chunk ? (n - done) : chunk;
    cudaStream_t st = g_pipe.st[slot];
    // work queued on a stream is ordered, so reusing the slot's buffers is safe
    if (f32) {
      e = cudaMemcpyAsync(g_pipe.d_q32[slot], static_cast<const float*>(q_host) + done * ndof,
                          sizeof(float) * m * ndof, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = launch_widen(g_pipe.d_q32[slot], g_pipe.d_q[slot], m * ndof, st);
    } else {
      e = cudaMemcpyAsync(g_pipe.d_q[slot], static_cast<const double*>(q_host) + done * ndof,
                          sizeof(double) * m * ndof, cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess)
      e = run_fk_cc_f64(robot, scene, g_pipe.d_q[slot], m, ACRO_Q_AOS, g_pipe.d_mask[slot], st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(mask_host + done / 32, g_pipe.d_mask[slot],
                          sizeof(uint32_t) * ((m + 31) / 32), cudaMemcpyDeviceToHost, st);
    done += m;
    slot = (slot + 1) % kSlots;
  }
  for (int s = 0; s < kSlots; ++s) {
    cudaError_t e2 = cudaStreamSynchronize(g_pipe.st[s]);
    if (e == cudaSuccess) e = e2;
  }
  return finish(e, what);
}

int acro_fk_cc_f64_host(const AcroRobot* robot, const AcroScene* scene, const double* q_host,
                        int64_t n, uint32_t* mask_host, int64_t chunk) {
  return host_pipeline(robot, scene, q_host, false, n, mask_host, chunk, "acro_fk_cc_f64_host");
}

int acro_fk_cc_f32_host(const AcroRobot* robot, const AcroScene* scene, const float* q_host,
                        int64_t n, uint32_t* mask_host, int64_t chunk) {
  return host_pipeline(robot, scene, q_host, true, n, mask_host, chunk, "acro_fk_cc_f32_host");
}

int acro_jac_pos_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                     double* J_out, void* stream) {
  if (int rc = check_batch(robot, q, J_out, n)) return rc;
  return finish(launch_jac(robot->d64, q, n, q_layout, J_out, false, as_stream(stream)),
                "acro_jac_pos_f64");
}

int acro_jac_rpy_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                     double* J_out, void* stream) {
  if (int rc = check_batch(robot, q, J_out, n)) return rc;
  return finish(launch_jac(robot->d64, q, n, q_layout, J_out, true, as_stream(stream)),
                "acro_jac_rpy_f64");
}

static int check_kuka(const AcroRobot* robot) {
  if (!robot) return fail(ACRO_E_INVALID, "robot handle is NULL");
  if (robot->d64.ndof != 6 || robot->d64.prismatic_mask != 0)
    return fail(ACRO_E_UNSUPPORTED, "Kuka IK needs a 6-dof all-revolute chain");
  // the closed form (robot_examples.py:188-218) is that of the Kuka's DH pattern
  // (robot_examples.py:161-168): alpha = (pi/2, 0, pi/2, -pi/2, pi/2, 0), a = (a1, a2, 0, 0, 0, 0),
  // d = (0, 0, 0, d4, 0, d6), no theta offsets.  Any other 6R chain must not get these formulas.
  const AcroRobotDesc& d = robot->desc;
  const double sa[6] = {1, 0, 1, -1, 1, 0}, ca[6] = {0, 1, 0, 0, 0, 1};
  bool ok = true;
  for (int i = 0; i < 6; ++i) {
    ok = ok && std::fabs(d.sin_alpha[i] - sa[i]) < 1e-12 && std::fabs(d.cos_alpha[i] - ca[i]) < 1e-12;
    ok = ok && d.theta[i] == 0.0;
    if (i >= 2) ok = ok && d.a[i] == 0.0;
    if (i != 3 && i != 5) ok = ok && d.d[i] == 0.0;
  }
  if (!ok)
    return fail(ACRO_E_UNSUPPORTED,
                "closed-form IK is only defined for the Kuka DH pattern (robot_examples.py:161-168)");
  return 0;
}

int acro_ik_kuka_f64(const AcroRobot* robot, const double* T, int64_t n, double* q_out,
                     uint8_t* valid, void* stream) {
  if (int rc = check_kuka(robot)) return rc;
  if (int rc = check_batch(robot, T, q_out, n)) return rc;
  if (n > 0 && !valid) return fail(ACRO_E_INVALID, "valid is NULL");
  return finish(launch_ik_kuka(robot->d64, T, n, q_out, valid, as_stream(stream)),
                "acro_ik_kuka_f64");
}

int acro_ik_kuka_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* T, int64_t n,
                        double* q_out, uint8_t* valid, uint8_t* free_mask, void* stream) {
  if (int rc = check_kuka(robot)) return rc;
  if (int rc = check_batch(robot, T, q_out, n)) return rc;
  if (int rc = check_scene_device(scene)) return rc;
  if (n > 0 && (!valid || !free_mask)) return fail(ACRO_E_INVALID, "valid/free_mask is NULL");
  cudaStream_t st = as_stream(stream);
  if (8 * n < kFilterMinBatch)
    return finish(launch_ik_kuka_cc(robot->d64, view_of<double>(scene), T, n, q_out, valid,
                                    free_mask, st),
                  "acro_ik_kuka_cc_f64");
  // large batches: IK kernel, then the filtered K2 kernel over the (8n, 6) solution rows
  // (absent branches are NaN rows and are skipped), then free = valid & ~hit.  Configuration
  // 8*i + s is bit s of byte i of the K2 mask, so the mask bytes line up with `valid`.
  cudaError_t e = launch_ik_kuka(robot->d64, T, n, q_out, valid, st);
  uint32_t* hit = nullptr;
  if (e == cudaSuccess) e = cudaMallocAsync(&hit, sizeof(uint32_t) * ((8 * n + 31) / 32), st);
  if (e == cudaSuccess)
    e = run_fk_cc_f64(robot, scene, q_out, 8 * n, ACRO_Q_AOS, hit, st, /*skip_nan=*/true);
  if (e == cudaSuccess)
    e = launch_free_mask(valid, reinterpret_cast<const uint8_t*>(hit), n, free_mask, st);
  if (hit) {
    cudaError_t e2 = cudaFreeAsync(hit, st);
    if (e == cudaSuccess) e = e2;
  }
  return finish(e, "acro_ik_kuka_cc_f64");
}

int acro_joint_solutions_csr_f64(const AcroRobot* robot, const AcroScene* scene, const double* T,
                                 int64_t n_points, int32_t samples_per_point, double* q_all,
                                 uint8_t* valid, uint8_t* free_mask, double* q_free,
                                 int64_t capacity, int64_t* offsets, void* stream) {
  if (n_points < 0 || samples_per_point < 1) return fail(ACRO_E_INVALID, "bad point / sample count");
  if (!offsets || capacity < 0 || (capacity > 0 && !q_free))
    return fail(ACRO_E_INVALID, "offsets / q_free is NULL");
  // the compaction moves rows of 6 doubles as three 16-byte pieces
  if ((reinterpret_cast<uintptr_t>(q_all) | reinterpret_cast<uintptr_t>(q_free)) & 15u)
    return fail(ACRO_E_INVALID, "q_all and q_free must be 16-byte aligned");
  const int64_t n = n_points * samples_per_point;
  if (int rc = acro_ik_kuka_cc_f64(robot, scene, T, n, q_all, valid, free_mask, stream)) return rc;
  cudaStream_t st = as_stream(stream);
  const size_t temp = n_points > 0 ? csr_temp_bytes(n_points) : 0;
  void* scratch = nullptr;
  cudaError_t e = cudaSuccess;
  if (n_points > 0) e = cudaMallocAsync(&scratch, sizeof(int64_t) * n_points + temp + 16, st);
  if (e == cudaSuccess)
    e = launch_csr_compact(q_all, free_mask, n_points, samples_per_point, offsets, q_free, capacity,
                           scratch, temp, st);
  if (scratch) {
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    if (e == cudaSuccess) e = e2;
  }
  return finish(e, "acro_joint_solutions_csr_f64");
}

int acro_shortest_path_f64(const double* q, int32_t ndof, const int64_t* offsets_host,
                           int64_t n_layers, int32_t cost_type, const double* weights_host,
                           const double* state_cost, double* value_work, int32_t* pred_work,
                           int64_t* offsets_work, int64_t* path_rows, double* length, void* stream) {
  if (!offsets_host || n_layers < 1) return fail(ACRO_E_INVALID, "need at least one layer");
  if (ndof < 1 || ndof > ACRO_MAX_DOF) return fail(ACRO_E_UNSUPPORTED, "ndof outside 1..ACRO_MAX_DOF");
  if (cost_type < 0 || cost_type > 3) return fail(ACRO_E_INVALID, "unknown cost function");
  if (cost_type == 3 && !weights_host) return fail(ACRO_E_INVALID, "weights are required");
  if (!q || !value_work || !pred_work || !offsets_work || !path_rows || !length)
    return fail(ACRO_E_INVALID, "NULL buffer");
  for (int64_t l = 0; l < n_layers; ++l)
    if (offsets_host[l + 1] <= offsets_host[l])
      return fail(ACRO_E_INVALID, "empty layer: no path exists");  // success = False upstream
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemcpyAsync(offsets_work, offsets_host, sizeof(int64_t) * (n_layers + 1),
                                  cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = launch_shortest_path(q, ndof, offsets_host, offsets_work, n_layers, cost_type, weights_host,
                             state_cost, value_work, pred_work, path_rows, length, st);
  return finish(e, "acro_shortest_path_f64");
}

int acro_ik_arm2_f64(double a1, double a2, double a3, const double* p, int64_t n, double* q_out,
                     uint8_t* valid, void* stream) {
  if (n < 0 || (n > 0 && (!p || !q_out || !valid))) return fail(ACRO_E_INVALID, "bad arguments");
  return finish(launch_ik_arm2(a1, a2, a3, p, n, q_out, valid, as_stream(stream)),
                "acro_ik_arm2_f64");
}

int acro_ik_anthropomorphic_f64(double l2, double l3, const double* p, int64_t n, double* q_out,
                                uint8_t* valid, void* stream) {
  if (n < 0 || (n > 0 && (!p || !q_out || !valid))) return fail(ACRO_E_INVALID, "bad arguments");
  return finish(launch_ik_anthro(l2, l3, p, n, q_out, valid, as_stream(stream)),
                "acro_ik_anthropomorphic_f64");
}

int acro_ik_wrist_f64(const double* R, const double* R_base_host, int64_t n, double* q_out,
                      void* stream) {
  if (n < 0 || !R_base_host || (n > 0 && (!R || !q_out)))
    return fail(ACRO_E_INVALID, "bad arguments");
  return finish(launch_ik_wrist(R, R_base_host, n, q_out, as_stream(stream)), "acro_ik_wrist_f64");
}

int acro_swept_boxes_f64(const AcroBox* a_start, const AcroBox* a_goal, const AcroBox* b, int64_t n,
                         int32_t n_samples, uint32_t* mask, void* stream) {
  if (n < 0 || (n > 0 && (!a_start || !a_goal || !b || !mask)))
    return fail(ACRO_E_INVALID, "bad arguments");
  if (n_samples == 0) n_samples = 10;
  if (n_samples < 2) return fail(ACRO_E_INVALID, "n_samples must be 0 (FCL's 10) or >= 2");
  return finish(launch_swept_boxes(a_start, a_goal, b, n, n_samples, mask, as_stream(stream)),
                "acro_swept_boxes_f64");
}

int acro_swept_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* q_start,
                      const double* q_goal, int64_t n_pairs, int32_t n_samples, uint32_t* mask,
                      uint32_t* near_mask, double margin, void* stream) {
  if (int rc = check_batch(robot, q_start, mask, n_pairs)) return rc;
  if (int rc = check_scene_device(scene)) return rc;
  if (n_pairs > 0 && !q_goal) return fail(ACRO_E_INVALID, "q_goal is NULL");
  if (n_samples == 0) n_samples = 10;  // FCL: min(num_max_iterations = 10, ceil(1 / toc_err))
  if (n_samples < 2) return fail(ACRO_E_INVALID, "n_samples must be 0 (FCL's 10) or >= 2");
  if (near_mask && !(margin >= 0.0)) return fail(ACRO_E_INVALID, "margin must be >= 0");
  cudaStream_t st = as_stream(stream);
  // "swept_persistent" = "1": large batches go to the persistent kernel (lanes take pairs from
  // a counter; the counter word comes from the stream-ordered pool).  Off by default: measured
  // slower than one thread per pair (profiles/r2_swept_ab.jsonl).
  unsigned long long* counter = nullptr;
  cudaError_t e = cudaSuccess;
  if (n_pairs >= kSweptPersistentMin && swept_persistent()) {
    e = cudaMallocAsync(reinterpret_cast<void**>(&counter), sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
  }
  if (e == cudaSuccess)
    e = launch_swept_cc(robot->d64, view_of<double>(scene), q_start, q_goal, n_pairs, n_samples,
                        mask, near_mask, margin, counter, st);
  if (counter) {
    cudaError_t e2 = cudaFreeAsync(counter, st);
    if (e == cudaSuccess) e = e2;
  }
  return finish(e, "acro_swept_cc_f64");
}

int acro_path_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* q_start,
                     const double* q_goal, int64_t n_segments, double max_q_step, uint32_t* mask,
                     void* stream) {
  if (int rc = check_batch(robot, q_start, mask, n_segments)) return rc;
  if (int rc = check_scene_device(scene)) return rc;
  if (n_segments > 0 && !q_goal) return fail(ACRO_E_INVALID, "q_goal is NULL");
  if (!(max_q_step > 0.0)) return fail(ACRO_E_INVALID, "max_q_step must be > 0");
  cudaStream_t st = as_stream(stream);
  if (k2_variant() < 3 || n_segments < 1024)
    return finish(launch_path_cc(robot->d64, view_of<double>(scene), q_start, q_goal, n_segments,
                                 max_q_step, mask, st),
                  "acro_path_cc_f64");
  static const AcroScene empty_scene{};
  const CullPlan* plan = nullptr;
  cudaError_t e = get_plan(robot, scene ? scene : &empty_scene, &plan, true, st);
  if (e != cudaSuccess) return finish(e, "acro_path_cc_f64");
  PlanPin pin;
  pin.scene = scene ? scene : &empty_scene;
  pin.plan = plan;
  const int64_t n_words = (n_segments + 31) / 32;
  void* scratch = nullptr;
  e = cudaMallocAsync(&scratch, ((sizeof(uint32_t) * n_words + 7) / 8) * 8 + 8, st);
  if (e == cudaSuccess)
    e = launch_path_cc_filter(robot->d32, robot->d64, plan->view, plan->mask_bits,
                              view_of<float>(scene), view_of<double>(scene), q_start, q_goal,
                              n_segments, max_q_step, mask, scratch, st);
  if (scratch) {
    cudaError_t e2 = cudaFreeAsync(scratch, st);
    if (e == cudaSuccess) e = e2;
  }
  return finish(e, "acro_path_cc_f64");
}

int acro_filter_tolerance(const AcroRobot* robot, const AcroScene* scene, double* delta) {
  if (!robot || !delta) return fail(ACRO_E_INVALID, "NULL argument");
  static const AcroScene empty_scene{};
  const CullPlan* plan = nullptr;
  cudaError_t e = get_plan(robot, scene ? scene : &empty_scene, &plan);
  if (e != cudaSuccess) return finish(e, "acro_filter_tolerance");
  *delta = (double)plan->view.delta;
  release_plan(scene ? scene : &empty_scene, plan);
  return 0;
}

int acro_measure_filter_error(const AcroRobot* robot, const AcroScene* scene, const double* q,
                              int64_t n, double* out2, void* stream) {
  if (int rc = check_batch(robot, q, out2, n)) return rc;
  if (!scene) return fail(ACRO_E_INVALID, "scene handle is NULL");
  return finish(launch_filter_error(robot->d32, robot->d64, view_of<float>(scene),
                                    view_of<double>(scene), q, n, out2, as_stream(stream)),
                "acro_measure_filter_error");
}

int acro_math_probe_f64(int32_t fn, const double* a, const double* b, double* out0, double* out1,
                        int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!a || !out0))) return fail(ACRO_E_INVALID, "bad arguments");
  if (fn < 0 || fn > 4) return fail(ACRO_E_INVALID, "fn must be 0..4");
  if (n > 0 && ((fn == 0 && !out1) || ((fn == 1 || fn == 3) && !b)))
    return fail(ACRO_E_INVALID, "missing operand / output array");
  return finish(launch_math_probe(fn, a, b, out0, out1, n, as_stream(stream)), "acro_math_probe_f64");
}

int acro_measure_fma_peak(int32_t fp64, int32_t iters, double* tflops, double* ms) {
  if (!tflops || !ms || iters <= 0) return fail(ACRO_E_INVALID, "bad arguments");
  return finish(measure_fma_peak(fp64 != 0, iters, tflops, ms), "acro_measure_fma_peak");
}

}  // extern "C"

// ---------------------------------------------------------------------------
// FMA-chain micro-benchmark (compute-roofline denominator)
// ---------------------------------------------------------------------------
namespace acro {

// element-wise evaluation of the lean FP64 functions of acro_math.cuh (test hook)
__global__ void __launch_bounds__(256)
    math_probe_kernel(int fn, const double* __restrict__ a, const double* __restrict__ b,
                      double* __restrict__ out0, double* __restrict__ out1, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = a[i];
  if (fn == 0) {
    double s, c;
    sincos_lean(x, &s, &c);
    out0[i] = s;
    out1[i] = c;
  } else if (fn == 1) {
    out0[i] = atan2_lean(x, b[i]);
  } else if (fn == 2) {
    out0[i] = sqrt_lean(x);
  } else if (fn == 3) {
    out0[i] = div_lean(x, b[i]);
  } else {
    out0[i] = rcp_lean(x);
  }
}

cudaError_t launch_math_probe(int fn, const double* a, const double* b, double* out0, double* out1,
                              int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fn, a, b, out0, out1, n);
  count_launch();
  return cudaGetLastError();
}

template <typename real>
__global__ void __launch_bounds__(256) fma_chain_kernel(real* out, int iters, real a, real b) {
  real x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = real(threadIdx.x + k) * real(1e-3);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fma_t<real>(x[k], a, b);
  }
  real s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == real(123.456)) out[0] = s;  // keep the chain alive
}

cudaError_t measure_fma_peak(bool fp64, int iters, double* tflops, double* ms_out) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  void* out = nullptr;
  e = cudaMalloc(&out, 64);
  if (e != cudaSuccess) return e;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  const int blocks = sms * 8, threads = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(t0, 0);
    if (fp64)
      fma_chain_kernel<double><<<blocks, threads>>>((double*)out, iters, 0.999999, 1e-7);
    else
      fma_chain_kernel<float><<<blocks, threads>>>((float*)out, iters, 0.999999f, 1e-7f);
    cudaEventRecord(t1, 0);
    e = cudaEventSynchronize(t1);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(out);
  if (e != cudaSuccess) return e;
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  *ms_out = best;
  *tflops = flops / (best * 1e-3) / 1e12;
  return cudaGetLastError();
}

}  // namespace acro
