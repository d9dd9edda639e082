/*
 * acro_b200.h - C ABI of libacro_b200.so: the B200 (sm_100a) implementation of
 * the acrobotics forward-kinematics + collision-checking hot path.
 *
 * The reference (JeroenDM/acrobotics, pure Python) has no FFI layer of its own
 * for this path: the seam is a set of Python methods plus one native call,
 *   fcl.collide(o1, o2, request, result)          src/acrobotics/shapes.py:52-58
 * Each entry point below names the reference interface it replaces.  A
 * maintainer binds them with ctypes.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every array is a raw pointer + sizes;
 *   - 4x4 homogeneous transforms travel as 12 doubles, row-major 3x4 (the
 *     bottom row [0 0 0 1] is implicit);
 *   - "device" entry points take DEVICE pointers and are asynchronous on the
 *     given CUDA stream (a cudaStream_t passed as void*, NULL = default stream);
 *     "_host" entry points take HOST pointers, run a chunked, double-buffered
 *     H2D -> kernel -> D2H pipeline on library-owned streams and return after
 *     the results are in the host buffer;
 *   - the caller owns every buffer; the library owns only the opaque handles;
 *     handles are immutable after creation, so calls are thread-safe given
 *     distinct output buffers;
 *   - return value: 0 = OK, negative = ACRO_E_* library error, positive =
 *     cudaError_t.  Nothing is thrown across the boundary; acro_last_error()
 *     returns a thread-local message.
 *   - collision masks are bit sets: word i/32, bit i%32 = configuration i,
 *     ceil(N/32) uint32 words, tail bits zero.
 */
#ifndef ACRO_B200_H
#define ACRO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACRO_ABI_VERSION 1

#define ACRO_MAX_DOF 8
#define ACRO_MAX_ROBOT_BOXES 24
#define ACRO_MAX_SELF_PAIRS 160
#define ACRO_MAX_SCENE_BOXES 64

#define ACRO_E_INVALID (-1)     /* bad argument                                  */
#define ACRO_E_UNSUPPORTED (-2) /* e.g. ndof or box count above the compiled caps */
#define ACRO_E_NOGPU (-3)       /* no usable CUDA device                          */

/* carrier codes for AcroRobotBox.carrier */
#define ACRO_CARRIER_BASE (-1) /* rides on tf_base          (robot.py:257-260)  */
#define ACRO_CARRIER_TOOL (-2) /* rides on last link frame  (robot.py:251-254)  */

/* q layouts */
#define ACRO_Q_AOS 0 /* (N, ndof) row-major: the reference's natural layout      */
#define ACRO_Q_SOA 1 /* (ndof, N): joint-major, fully coalesced                  */

/* One oriented box: reference Box(dx,dy,dz) (shapes.py:105-114) + its 4x4 tf.   */
typedef struct AcroBox {
  double half[3]; /* dx/2, dy/2, dz/2 */
  double tf[12];  /* row-major 3x4    */
} AcroBox;

/* A collision box of the robot: link.geometry / geometry_tool / geometry_base
 * entries (link.py:80-83, robot.py:178-179, :197-204).                          */
typedef struct AcroRobotBox {
  int32_t carrier; /* link index 0..ndof-1, ACRO_CARRIER_BASE or ACRO_CARRIER_TOOL */
  int32_t reserved;
  double half[3];
  double tf[12]; /* offset of the box in its carrier frame */
} AcroRobotBox;

/* Flattened Robot (robot.py:36-57, :174-195) + DH table (link.py:8, :35-58).    */
typedef struct AcroRobotDesc {
  int32_t ndof;
  int32_t joint_type[ACRO_MAX_DOF]; /* 0 revolute ('r'), 1 prismatic ('p') (link.py:11-13) */
  double a[ACRO_MAX_DOF];
  double d[ACRO_MAX_DOF];
  double theta[ACRO_MAX_DOF];
  /* cos/sin of the DH alpha, computed on the host exactly as the reference does
   * (np.cos(alpha), link.py:48-49: cos(pi/2) is 6.1e-17, not 0).                */
  double cos_alpha[ACRO_MAX_DOF];
  double sin_alpha[ACRO_MAX_DOF];
  double tf_base[12];     /* robot.py:55 */
  int32_t has_tool_tip;   /* robot.py:57 (None -> 0) */
  int32_t check_self;     /* robot.py:182 do_check_self_collision */
  double tf_tool_tip[12]; /* relative to the last link */
  int32_t n_boxes;
  int32_t n_self_pairs;
  AcroRobotBox boxes[ACRO_MAX_ROBOT_BOXES];
  /* unordered pairs of indices into boxes[] that the self-collision test visits
   * (collision_matrix band robot.py:185-186, :206-211; tool vs links :214-220)  */
  int32_t self_pairs[ACRO_MAX_SELF_PAIRS][2];
} AcroRobotDesc;

typedef struct AcroRobot AcroRobot; /* opaque */
typedef struct AcroScene AcroScene; /* opaque */

int acro_abi_version(void);
const char* acro_last_error(void);
/* number of CUDA devices visible, <0 on error; never initialises a context */
int acro_device_count(void);

/* Tuning knobs (process wide; defaults are the production settings):
 *   "k2_variant" = "filter" (default) | "wave" | "simple": kernel behind acro_fk_cc_f64
 *                  without a near mask; all variants return identical verdicts;
 *   "k2_split"   = "auto" (default) | 0 | k: slot after which the filtered kernel regroups the
 *                  still undecided configurations into full warps (0 = single pass);
 *   "k2_grid"    = cells per axis of the candidate grid of the filtered kernel (default 128);
 *   "swept_persistent" = "0" (default) | "1": acro_swept_cc_f64 runs one thread per pair, or
 *                  (>= 4096 pairs) the persistent kernel whose lanes take pairs from a counter;
 *                  identical verdicts, the persistent kernel measured slower.
 * The environment variables ACRO_K2_VARIANT / ACRO_K2_SPLIT / ACRO_K2_GRID set the initial values.    */
int acro_set_option(const char* name, const char* value);

/* Robot(links, joint_limits) + geometry  ->  immutable device descriptor.
 * Replaces the per-call Python object walk in robot.py:244-273.                  */
int acro_robot_create(const AcroRobotDesc* desc, AcroRobot** out);
void acro_robot_destroy(AcroRobot* robot);

/* Scene(shapes, tf_shapes) (geometry.py:12-14), boxes only.  n may be 0.         */
int acro_scene_create(const AcroBox* boxes, int32_t n, AcroScene** out);
void acro_scene_destroy(AcroScene* scene);
/* Optional: build now what acro_fk_cc_f64 would otherwise build on its first LARGE call
 * (>= 65536 configurations) with this (robot, scene) pair - the candidate grid of the
 * filtered kernel (about 50 MB of device memory owned by the scene handle for six radius classes, ~1 ms; a scene
 * keeps the eight most recently used grids).  This call waits for the build; a first-use
 * build inside a hot call is ordered on the caller's stream instead (one cudaMalloc, no
 * device-wide synchronisation; other streams wait on an event).  With a grid in place,
 * batches of 4096 configurations and more take the filtered kernel and the hot call
 * allocates nothing but stream-ordered scratch (cudaMallocAsync) for one bit per
 * configuration - or nothing at all through acro_fk_cc_f64_ws; without a grid, smaller
 * batches run the all-FP64 kernel.                                                    */
int acro_scene_prepare(const AcroScene* scene, const AcroRobot* robot);

/* ---- the fcl.collide seam -------------------------------------------------------
 * n independent Box-Box tests: a[i] against b[i] (each an AcroBox with its WORLD
 * transform), FCL boxBox2 semantics (touching = colliding).  Replaces
 * Shape.is_in_collision -> fcl.collide (shapes.py:50-58), the one native call on
 * the reference's path.  mask bit i = pair i collides.  Device pointers.         */
int acro_collide_boxes_f64(const AcroBox* a, const AcroBox* b, int64_t n, uint32_t* mask,
                           void* stream);

/* ---- K1: forward kinematics ------------------------------------------------
 * Robot.fk (robot.py:59-69) when all_links == 0: T_out is (N,4,4) row-major,
 * tool tip applied.  Robot.fk_all_links (robot.py:82-91) when all_links != 0:
 * T_out is (N,ndof,4,4), base applied, tool tip not.                           */
int acro_fk_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                double* T_out, int32_t all_links, void* stream);
int acro_fk_f32(const AcroRobot* robot, const float* q, int64_t n, int32_t q_layout,
                float* T_out, int32_t all_links, void* stream);
/* Robot.fk_rpy (robot.py:71-80): out (N,6) = [x,y,z,r_x,r_y,r_z].               */
int acro_fk_rpy_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                    double* out, void* stream);

/* ---- K2: fused FK -> collision ----------------------------------------------
 * Robot.is_in_collision(q, scene) (robot.py:244-273) for N configurations:
 * tool x scene, base x scene, link x scene, and (check_self) the self pairs,
 * each box pair decided by FCL's boxBox2 separating-axis test (shapes.py:50-58).
 * scene may be NULL (self collision only: Robot.is_in_self_collision,
 * robot.py:239-242).  near_mask (nullable): bit set when the configuration's
 * decisive SAT quantity is within +-margin of contact (reported separately).
 * n < 2^32 per call (larger batches: several calls over row ranges).           */
int acro_fk_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* q, int64_t n,
                   int32_t q_layout, uint32_t* mask, uint32_t* near_mask, double margin,
                   void* stream);
/* Same with a caller-owned workspace of acro_fk_cc_workspace_bytes(n) bytes of DEVICE memory
 * (8-byte aligned; contents undefined on entry and exit): with it, and with the (robot, scene)
 * pair prepared by acro_scene_prepare, the hot call allocates nothing and synchronises nothing.
 * workspace may be NULL (then this is acro_fk_cc_f64).                                      */
int64_t acro_fk_cc_workspace_bytes(int64_t n);
int acro_fk_cc_f64_ws(const AcroRobot* robot, const AcroScene* scene, const double* q, int64_t n,
                      int32_t q_layout, uint32_t* mask, uint32_t* near_mask, double margin,
                      void* workspace, int64_t workspace_bytes, void* stream);
/* FP32 joint values.  Small batches run an all-FP32 kernel (poses good to ~1e-5); large AoS
 * batches are widened exactly and take the FP64-exact filtered path, i.e. the verdicts are
 * those of acro_fk_cc_f64 at the float-valued joint angles.                              */
int acro_fk_cc_f32(const AcroRobot* robot, const AcroScene* scene, const float* q, int64_t n,
                   int32_t q_layout, uint32_t* mask, void* stream);
/* Same, HOST buffers: pipelined H2D/kernel/D2H inside the call.  The pipeline (three staging
 * slots, three streams) is per CUDA device: calls on different devices run concurrently from
 * different host threads, calls on one device take turns.                       */
int acro_fk_cc_f64_host(const AcroRobot* robot, const AcroScene* scene, const double* q_host,
                        int64_t n, uint32_t* mask_host, int64_t chunk);
/* The same pipeline for FP32 joint values on the host (24 instead of 48 bytes per 6-dof
 * configuration over PCIe - the link is what bounds the host path): the rows are widened
 * exactly on the device, the verdicts are those of the FP64 evaluation at the float-valued
 * joint angles.                                                                  */
int acro_fk_cc_f32_host(const AcroRobot* robot, const AcroScene* scene, const float* q_host,
                        int64_t n, uint32_t* mask_host, int64_t chunk);

/* ---- K3/K4: Jacobians (robot.py:157-171; casadi AD in the reference) ----------
 * J_out (N,3,ndof) / (N,6,ndof) row-major.  The descriptor's base / tool tip are
 * used as given (the Python layer passes the as-constructed ones).              */
int acro_jac_pos_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                     double* J_out, void* stream);
int acro_jac_rpy_f64(const AcroRobot* robot, const double* q, int64_t n, int32_t q_layout,
                     double* J_out, void* stream);

/* ---- K5: analytical IK ---------------------------------------------------------
 * Kuka.ik (robot_examples.py:188-218 = arm_2.py:5-104 + spherical_wrist.py:5-30).
 * T (N,4,4) row-major -> q_out (N,8,6), slot = 2*arm_branch + wrist_branch in the
 * reference's order; valid[i] bit s = slot s exists (invalid slots are NaN).
 * The robot must have the Kuka's DH pattern (robot_examples.py:161-168: alpha =
 * (pi/2, 0, pi/2, -pi/2, pi/2, 0), a = (a1, a2, 0, 0, 0, 0), d = (0, 0, 0, d4, 0, d6), no
 * theta offsets); any other chain is refused with ACRO_E_UNSUPPORTED.                */
int acro_ik_kuka_f64(const AcroRobot* robot, const double* T, int64_t n, double* q_out,
                     uint8_t* valid, void* stream);
/* arm_2.ik: p (N,3) -> q_out (N,4,3), valid bits [pos_a,pos_b,neg_a,neg_b].      */
int acro_ik_arm2_f64(double a1, double a2, double a3, const double* p, int64_t n, double* q_out,
                     uint8_t* valid, void* stream);
/* anthropomorphic_arm.ik: p (N,3) -> q_out (N,4,3), valid[i] = 0 or 0xF.         */
int acro_ik_anthropomorphic_f64(double l2, double l3, const double* p, int64_t n, double* q_out,
                                uint8_t* valid, void* stream);
/* spherical_wrist.ik: R (N,3,3), R_base (3,3) host -> q_out (N,2,3).             */
int acro_ik_wrist_f64(const double* R, const double* R_base_host, int64_t n, double* q_out,
                      void* stream);
/* Fused IK -> collision filter (path/path_pt_base.py:76-85, :107-109):
 * free_mask[i] bit s = slot s exists AND is collision free.                     */
int acro_ik_kuka_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* T,
                        int64_t n, double* q_out, uint8_t* valid, uint8_t* free_mask,
                        void* stream);

/* ---- planner inner loop ------------------------------------------------------------
 * PathPt.to_joint_solutions (path/path_pt_base.py:87-111) for n_points path points with
 * samples_per_point sampled poses each (T is (n_points * samples_per_point, 4, 4), point
 * major): IK of every pose, collision filter, and compaction of the surviving joint vectors
 * into CSR form.  q_all (N,8,6), valid (N), free_mask (N) are the outputs of
 * acro_ik_kuka_cc_f64 (kept: the caller may want them).  q_free receives up to `capacity`
 * rows of 6 doubles; rows offsets[p] .. offsets[p+1]-1 belong to point p, in the reference's
 * order (samples in order, each pose's solutions in Kuka.ik's order); offsets has
 * n_points + 1 entries and offsets[n_points] is the total number of surviving solutions
 * (rows beyond `capacity` are not written - compare the total with the capacity).
 * q_all and q_free must be 16-byte aligned (rows are moved as 16-byte pieces).           */
int acro_joint_solutions_csr_f64(const AcroRobot* robot, const AcroScene* scene, const double* T,
                                 int64_t n_points, int32_t samples_per_point, double* q_all,
                                 uint8_t* valid, uint8_t* free_mask, double* q_free,
                                 int64_t capacity, int64_t* offsets, void* stream);

/* Shortest path through the layered graph of joint solutions (find_shortest_joint_path,
 * planning/sampling_based.py:109-142 -> acrolib.dynamic_programming.shortest_path[_with_state_cost]).
 * q (M, ndof) device rows in CSR order, offsets_host (n_layers + 1, HOST: one launch per layer
 * is sized from it).  cost_type: 0 norm_l1, 1 norm_l2, 2 sum_squared, 3 weighted_sum_squared
 * (weights_host: ndof doubles).  state_cost: M device doubles or NULL.  Work buffers (device):
 * value_work M doubles, pred_work M int32, offsets_work n_layers + 1 int64.  Outputs (device):
 * path_rows[l] = row of q chosen in layer l, *length = total cost.                        */
int acro_shortest_path_f64(const double* q, int32_t ndof, const int64_t* offsets_host,
                           int64_t n_layers, int32_t cost_type, const double* weights_host,
                           const double* state_cost, double* value_work, int32_t* pred_work,
                           int64_t* offsets_work, int64_t* path_rows, double* length, void* stream);

/* ---- K6: discretised path sweep -----------------------------------------------
 * Robot.is_path_in_collision_discrete (robot.py:229-237, :275-283) for E
 * segments: q_start/q_goal (E,ndof); mask bit e = some interpolant collides.    */
int acro_path_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* q_start,
                     const double* q_goal, int64_t n_segments, double max_q_step, uint32_t* mask,
                     void* stream);

/* ---- K7: swept (continuous) collision --------------------------------------------
 * Robot.is_path_in_collision (robot.py:285-317) -> ShapeSoup.is_path_in_collision
 * (geometry.py:68-81) -> fcl.continuousCollide(..., CCDM_LINEAR) (shapes.py:60-75) for E
 * (q_start, q_goal) pairs: every tool / link box moves between its two world poses under
 * FCL's InterpMotion (chord translation, constant-axis rotation), sampled n_samples times
 * (0 = FCL's default request: 10) with the Box-Box test at every sample; base boxes fixed;
 * no self collision.  mask bit e = the pair collides.  near_mask (or NULL) / margin as in
 * acro_fk_cc_f64.  Stateless: the reference's python-fcl result object is sticky per
 * Shape.                                                          */
/* The fcl.continuousCollide seam itself: Shape.is_path_in_collision (shapes.py:60-75) for n
 * independent pairs - box a moves from a_start[i] to a_goal[i] (same half extents; FCL's
 * InterpMotion, n_samples samples, 0 = 10), box b[i] is fixed.  mask bit i = they collide at
 * some sample.                                                                            */
int acro_swept_boxes_f64(const AcroBox* a_start, const AcroBox* a_goal, const AcroBox* b, int64_t n,
                         int32_t n_samples, uint32_t* mask, void* stream);
int acro_swept_cc_f64(const AcroRobot* robot, const AcroScene* scene, const double* q_start,
                      const double* q_goal, int64_t n_pairs, int32_t n_samples, uint32_t* mask,
                      uint32_t* near_mask, double margin, void* stream);

/* ---- measurement helpers ----------------------------------------------------------
 * DFMA / FFMA chain micro-benchmark: the compute-roofline denominator.  Returns
 * achieved TFLOP/s (FMA = 2 flop) in *tflops; runs on the current device.        */
int acro_measure_fma_peak(int32_t fp64, int32_t iters, double* tflops, double* ms);
/* Test hook for the FP64 elementary functions the kernels use in place of the CUDA library's
 * (csrc/acro_math.cuh): element-wise over device arrays.  fn = 0 sincos(a) -> out0 = sin,
 * out1 = cos; 1 atan2(a, b); 2 sqrt(a); 3 a / b; 4 1 / a (3, 4: normal b / a only).          */
int acro_math_probe_f64(int32_t fn, const double* a, const double* b, double* out0, double* out1,
                        int64_t n, void* stream);
/* The filtered K2 kernel decides in FP32 with a tolerance delta.  These two
 * calls expose the evidence: the tolerance used for a (robot, scene) pair, and the measured
 * distance between the FP32 and FP64 decision quantities over a batch of configurations
 * (device pointers; out2[0] = max |s_fp32 - s_fp64| over every configuration x robot box x
 * scene box x axis, out2[1] = max error of a box centre).  Revolute joints within |q| < 2^26 (6.7e7). */
int acro_filter_tolerance(const AcroRobot* robot, const AcroScene* scene, double* delta);
int acro_measure_filter_error(const AcroRobot* robot, const AcroScene* scene, const double* q,
                              int64_t n, double* out2, void* stream);
/* counters of kernels launched by this library since load (all threads)          */
int64_t acro_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ACRO_B200_H */
