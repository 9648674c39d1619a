/*
 * diffco_b200 — C ABI of the B200-native DiffCo collision-score hot path.
 *
 * This is the drop-in boundary.  The reference (ucsdarclab/diffco) is pure Python and has no
 * FFI layer; the functions below are what a Python binding for its hot path binds (INTEGRATION.md shows the
 * ctypes stubs), one entry point per reference call site:
 *
 *   dc_score_grad      <- DiffCo.score_original      diffco/kernel_perceptrons.py:362-370
 *                         DiffCo.poly_score          diffco/kernel_perceptrons.py:309-319
 *                         MultiDiffCo.score/rbf_score diffco/deprecated/MultiDiffCo.py:118-123,156-170
 *                         + the autograd backward the optimisers run through them (diffco/optim.py:86-103,209-218)
 *   dc_kernel_matrix   <- kernel_func(X_t[i], X_t)   diffco/kernel_perceptrons.py:118 (training row)
 *                         rbf_kernel(S, S)           diffco/kernel_perceptrons.py:280 (fit_poly)
 *                         kernel_func(S, novel)      diffco/kernel_perceptrons.py:246 (jump start)
 *   dc_fk_forward/vjp  <- *.fkine                    diffco/model.py:40-48,90-93,156-159,225-241,366-383,430-453,486-503
 *   dc_perceptron_train<- DiffCo.train_perceptron    diffco/kernel_perceptrons.py:98-158
 *   dc_traj_step       <- Weighted.step iteration   diffco/optim.py:706-752 (penalty, its gradient, Adam, wrap)
 *   dc_pack_supports_tc<- (none)  optional tensor-core operand image of the support set: with it, dc_score_grad runs
 *                         DiffCo.score (RQKernel p = 2, one class, F <= 14, fp32) on tcgen05 tensor cores
 *
 * Conventions: all pointers except descriptors are DEVICE pointers to row-major contiguous arrays of `dtype`
 * (DC_F32 / DC_F64); descriptors are plain-old-data structs in HOST memory, copied by value into the launch.
 * Every call enqueues work on `stream` and returns without synchronising; return value 0 = ok, <0 = dc_status.
 * No torch types, no exceptions; the only process-wide state is the launch counter and the dc_set_option knobs.
 */
#ifndef DIFFCO_B200_H
#define DIFFCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DC_ABI_VERSION 7

#define DC_MAX_DOF 16
#define DC_MAX_LINKS 16
#define DC_MAX_KEYPOINTS 16
#define DC_MAX_ARMS 2
#define DC_MAX_ARM_JOINTS 8
#define DC_MAX_TOOL_POINTS 2
#define DC_MAX_FEATURES 64
#define DC_MAX_TREE_NODES 24
#define DC_MAX_CLASSES 8

typedef struct CUstream_st* dc_stream_t; /* == cudaStream_t */

typedef enum dc_status {
  DC_OK = 0,
  DC_ERR_INVALID_ARG = -1,
  DC_ERR_UNSUPPORTED = -2,
  DC_ERR_CUDA = -3,
  DC_ERR_NO_DEVICE = -4
} dc_status;

typedef enum dc_dtype { DC_F32 = 0, DC_F64 = 1 } dc_dtype;

/* Forward-kinematics feature maps q[B,dof] -> X[B, n_points*point_dim] (diffco/model.py). */
typedef enum dc_fk_type {
  DC_FK_NONE = 0,               /* transform=None: the configuration itself is the feature vector     */
  DC_FK_PLANAR_CHAIN = 1,       /* RevolutePlanarRobot.fkine           model.py:40-48                   */
  DC_FK_SE2_BODY = 2,           /* RigidPlanarBody.fkine               model.py:90-93                   */
  DC_FK_SE3_BODY = 3,           /* RigidBody.fkine                     model.py:156-159                 */
  DC_FK_DH_ARMS = 4,            /* Baxter L/R/dual, Panda, DualPanda   model.py:225-241,366-383,430-503 */
  DC_FK_SE2_BASE_PLANAR_ARM = 5, /* SE(2) base carrying a planar chain  (BASELINE.json configs[3])       */
  DC_FK_JOINT_TREE = 6          /* URDF kinematic tree: URDFRobot.compute_forward_kinematics_all_links,
                                   collision_interfaces/urdf_interface.py:517-553 + rigid_body.py:86-141 */
} dc_fk_type;

/* One serial standard-DH arm (utils.DH2mat, diffco/utils.py:66-77). */
typedef struct dc_dh_arm {
  int32_t n_joints;
  int32_t n_tool;                          /* points rigidly attached to the last frame (Panda fingers)   */
  int32_t joint_index[DC_MAX_ARM_JOINTS];  /* column of q driving joint i                                 */
  int32_t out_slot[DC_MAX_ARM_JOINTS];     /* output point index of frame i's origin, or -1 (fk_mask)     */
  int32_t tool_slot[DC_MAX_TOOL_POINTS];
  double a[DC_MAX_ARM_JOINTS];
  double d[DC_MAX_ARM_JOINTS];
  double s_alpha[DC_MAX_ARM_JOINTS];
  double c_alpha[DC_MAX_ARM_JOINTS];
  double theta0[DC_MAX_ARM_JOINTS];
  double base[12];                         /* row-major 3x4 [R|t] pre-multiplied base (identity if unused) */
  double offset[3];                        /* translation added to the outputs (DualPandaFK bases)        */
  double tool[DC_MAX_TOOL_POINTS][3];
} dc_dh_arm;

/* One body of a URDF tree with the joint that attaches it to its parent (rigid_body.py:86-141).  Bodies are listed
 * parents first.  Frame of body i:  R_i = R_parent rot Rot(q'),  t_i = t_parent + R_parent trans   (revolute / continuous)
 *                                   R_i = R_parent rot,          t_i = t_parent + R_parent (trans + rot axis q')  (prismatic)
 * with q' = mimic_mul * q[q_index] + mimic_off (q_index < 0: fixed joint, q' = 0). */
typedef enum dc_joint_kind { DC_JOINT_FIXED = 0, DC_JOINT_REV_X = 1, DC_JOINT_REV_Y = 2, DC_JOINT_REV_Z = 3, DC_JOINT_PRISMATIC = 4 } dc_joint_kind;
typedef struct dc_tree_node {
  int32_t parent;   /* index of the parent body, -1 for the root (its rot / trans then hold the robot's base transform) */
  int32_t q_index;  /* column of q driving the joint, -1 = fixed */
  int32_t joint;    /* dc_joint_kind */
  int32_t out_slot; /* output point index of the body frame's origin, or -1 */
  double rot[9];    /* row-major fixed rotation of the joint origin (rpy) */
  double trans[3];  /* joint origin translation */
  double axis[3];   /* prismatic: the axis; revolute: axis[0] = sign applied to q' (rigid_body.py:103-108) */
  double mimic_mul; /* 1 unless the joint mimics another */
  double mimic_off;
} dc_tree_node;

typedef struct dc_fk_desc {
  int32_t type;       /* dc_fk_type */
  int32_t dof;        /* D */
  int32_t n_points;   /* M */
  int32_t point_dim;  /* d in {1,2,3}; features F = M*d (type NONE: F = dof) */
  int32_t n_arms;
  int32_t n_keypoints;
  int32_t n_links;
  int32_t n_repeat;   /* > 1: the map is applied to n_repeat consecutive blocks of (dof - time_last)/n_repeat columns and the
                         features are concatenated (LineFKKernel, kernel.py:145-173: a path segment's two end configurations
                         side by side); n_points * point_dim is then the TOTAL feature count and point_dim = 1 */
  int32_t time_last;  /* 1: the last column of q is a time stamp passed through as the last feature (TemporalFKKernel,
                         kernel.py:175-202) */
  int32_t n_nodes;    /* DC_FK_JOINT_TREE: bodies in tree[] */
  double link_length[DC_MAX_LINKS];
  double keypoints[3][DC_MAX_KEYPOINTS]; /* body-frame key points, row r = coordinate r */
  dc_dh_arm arms[DC_MAX_ARMS];
  dc_tree_node tree[DC_MAX_TREE_NODES];
} dc_fk_desc;

/* Radial kernels (diffco/kernel.py). */
typedef enum dc_kernel_kind {
  DC_K_RQ = 1,           /* RQKernel      kernel.py:12-29   k = (1 + gamma/p r^2)^-p   param=gamma order=p */
  DC_K_POLYHARMONIC = 2, /* Polyharmonic  kernel.py:59-79   k = r^k/eps | r^k log r/eps param=eps   order=k */
  DC_K_MULTIQUADRIC = 3, /* MultiQuadratic kernel.py:45-57  k = sqrt(r^2/eps^2 + 1)    param=eps          */
  DC_K_RQ_TEMPORAL = 4   /* TemporalFKKernel kernel.py:175-202: the LAST feature is time;
                            k = RQ(param, order)(|dx|^2 over the first F-1 features) * RQ(param2, order2)(dt^2) ^ alpha */
} dc_kernel_kind;

typedef struct dc_kernel_desc {
  int32_t kind;
  int32_t order;
  double param;
  double param2; /* DC_K_RQ_TEMPORAL: gamma of the time kernel */
  double alpha;  /* DC_K_RQ_TEMPORAL: exponent of the time kernel */
  int32_t order2; /* DC_K_RQ_TEMPORAL: p of the time kernel */
  int32_t reserved;
} dc_kernel_desc;

/* Support set packed for the fused kernels: row n = [-s_n[0..F) zero-padded to f_pad | w[n,0..C) zero-padded to 1 or a
 * multiple of 4 classes], rows padded to a multiple of 4 elements (16-byte aligned for float). */
typedef struct dc_supports {
  const void* table;   /* device, [n][row_stride] of dtype, 16-byte aligned rows */
  int64_t n;           /* N support vectors */
  int32_t n_features;  /* F */
  int32_t n_class;     /* C */
  int32_t f_pad;
  int32_t row_stride;  /* elements per row */
  int32_t dtype;
  int32_t reserved;
  const void* tc_blob; /* device, optional (NULL = none): image written by dc_pack_supports_tc for the same S_feat / W */
  double tc_s2max;     /* max_n |s_n|^2 if known (> 0): lets dc_score_grad skip the tensor-core path when the kernel
                          width makes too many pairs "near" (they are re-evaluated exactly on the FP32 pipe); 0 = unknown */
  double tc_gamma;     /* RQKernel gamma the image was packed for (the kernel width is folded into the operands) */
  const void* table_lo; /* device, optional (NULL = none), float32 only: [n][row_stride] low parts -(s - fl32(s)) of the
                          supports' features from dc_pack_supports_lo; the tensor-core kernel's exact near-pair path adds
                          them to its differences x - s */
} dc_supports;

typedef enum dc_grad_mode {
  DC_GRAD_NONE = 0, /* score only                                                                   */
  DC_GRAD_SUM = 1,  /* grad[B,D]   = sum_c grad_out[b,c] * d score[b,c]/dq   (grad_out NULL = ones) */
  DC_GRAD_JAC = 2   /* grad[B,C,D] = d score[b,c]/dq                                                */
} dc_grad_mode;

int dc_abi_version(void);
const char* dc_status_string(int status);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches evidence). */
int64_t dc_launch_count(void);

/* Layout of the packed support table for (F, C, dtype). */
int dc_supports_layout(int32_t n_features, int32_t n_class, int32_t dtype, int32_t* f_pad, int32_t* row_stride);
/* Pack S_feat[N,F] (= support_transformed.reshape(N,-1)) and W[N,C] (gains or rbf_nodes) into `table`. */
int dc_pack_supports(const void* s_feat, const void* w, int64_t n, int32_t n_features, int32_t n_class, int32_t dtype,
                     void* table, dc_stream_t stream);

/*
 * Tensor-core operand image of a support set (fp32, one class, n_features <= 30, RQKernel with p = 2; the layout
 * depends on whether n_features <= 14, so every call names n_features):
 * dc_supports_tc_bytes gives the buffer size (DC_ERR_UNSUPPORTED for shapes the tensor-core kernel does not cover),
 * dc_pack_supports_tc fills a 128-byte aligned device buffer from S_feat[N,F], W[N] and the kernel (gamma and the
 * weights are folded into the operands, so the image belongs to this (S, W, kernel) triple).  dc_supports_tc_info
 * (synchronises; call once after packing) returns max|s|^2 and whether the scaled operands fit fp16's range — put the
 * pointer into dc_supports.tc_blob, gamma into tc_gamma and max|s|^2 into tc_s2max only when `valid` is 1.
 */
int dc_supports_tc_bytes(int64_t n, int32_t n_features, int32_t n_class, int32_t dtype, int64_t* bytes);
int dc_pack_supports_tc(const void* s_feat, const void* w, int64_t n, int32_t n_features, const dc_kernel_desc* kernel,
                        void* blob, dc_stream_t stream);
int dc_supports_tc_info(const void* blob, int64_t n, int32_t n_features, double* s2max, int32_t* valid);

/* Process-wide tuning knobs (the only global state of the library besides the launch counter). */
typedef enum dc_option {
  DC_OPT_TC_ENABLE = 1,   /* 0 / 1 (default 1; environment DIFFCO_B200_TC=0 also disables): use the tensor-core kernel */
  DC_OPT_TC_ERR_COEF = 2, /* bound on the error of the tensor-core rho, relative to |x|^2 + max|s|^2 (default 5e-7)     */
  DC_OPT_TC_TOL_PAIR = 3, /* admissible error of one pair's kernel value (default 2e-7): sets the near-pair threshold   */
  DC_OPT_TC_MIN_BATCH = 4, /* smallest batch sent to the tensor-core kernel (default 4096)                              */
  DC_OPT_TC_STATS = 5,     /* set 1: (re)start counting the pairs re-evaluated exactly, 0: stop; get: the count (syncs)  */
  DC_OPT_PEER_TIMEOUT_S = 6 /* seconds dc_peer_barrier spins before it gives up and traps (default 120)                  */
} dc_option;
int dc_set_option(int32_t option, double value);
double dc_get_option(int32_t option);
/* Which kernel the last dc_score_grad call of this process launched: 0 lane-split, 1 thread-per-query, 2 tensor-core. */
int dc_last_score_kernel(void);

/*
 * The hot path.  score[B,C] = sum_n w[n,c] k(|FK(q_b) - s_n|^2) and, per grad_mode, its gradient w.r.t. q.
 * fk->type == DC_FK_NONE: q is the feature matrix X[B,F] itself (poly_score(transformed_point=...),
 * kernel_perceptrons.py:316-317) and gradients are w.r.t. X.
 * score_ld / grad_ld: elements between consecutive rows of score / grad; 0 = dense (C, and D or C*D).  Passing the
 * two halves of one [B, C+D] buffer (score_ld = grad_ld = C+D) makes the launch write the fused record that the
 * multi-GPU all-gather ships.
 */
int dc_score_grad(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                  int64_t batch, void* score, int64_t score_ld, void* grad, int64_t grad_ld, const void* grad_out,
                  int32_t grad_mode, dc_stream_t stream);

/*
 * Host buffers in, host buffers out (what a CPU-side caller such as the reference's optimisers or an evaluation grid
 * holds): q_host[B,D] and out_host[B, C (+D with DC_GRAD_SUM)] = [score | grad] records are HOST pointers, ideally
 * pinned.  The pipeline object owns a few streams and device staging buffers; the call cuts the batch into chunk_rows
 * pieces that flow H2D -> dc_score_grad -> D2H on those streams (forked from and joined back into `stream`), overlapping
 * the PCIe copies with the kernels, and returns without synchronising.  One pipeline per calling thread.
 */
typedef struct dc_host_pipeline dc_host_pipeline;
int dc_host_pipeline_create(dc_host_pipeline** out, int64_t chunk_rows, int32_t n_slots);
void dc_host_pipeline_destroy(dc_host_pipeline* pipeline);
int dc_score_grad_host(dc_host_pipeline* pipeline, const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv,
                       const void* q_host, int64_t batch, void* out_host, int32_t grad_mode, dc_stream_t stream);

/* K[Na,Nb] = k(|xa_i - xb_j|^2) on pre-transformed features (training rows, fit_poly, jump-start block). */
int dc_kernel_matrix(const dc_kernel_desc* kernel, const void* xa, int64_t na, const void* xb, int64_t nb,
                     int32_t n_features, int32_t dtype, void* k_out, dc_stream_t stream);

/* X[B,F] = FK(q[B,D]);  gq[B,D] = J_FK(q)^T gX[B,F].  float32 features are evaluated in float64 and rounded once;
 * dc_fk_forward_split (float32) also returns what the rounding dropped: FK(q) = x_hi + x_lo to ~1e-9, x_hi == dc_fk_forward's
 * result.  dc_pack_supports_lo lays S_lo[N,F] out like the packed support table (dc_supports.table_lo). */
int dc_fk_forward(const dc_fk_desc* fk, const void* q, int64_t batch, int32_t dtype, void* x_out, dc_stream_t stream);
int dc_fk_forward_split(const dc_fk_desc* fk, const void* q, int64_t batch, void* x_hi, void* x_lo, dc_stream_t stream);
int dc_pack_supports_lo(const void* s_lo, int64_t n, int32_t n_features, int32_t n_class, void* table_lo, dc_stream_t stream);
int dc_fk_vjp(const dc_fk_desc* fk, const void* q, int64_t batch, int32_t dtype, const void* g_x, void* g_q,
              dc_stream_t stream);
/* DC_FK_JOINT_TREE only: frames[B][n_nodes][12] = row-major [R | t] of every body (what the reference's
 * compute_forward_kinematics_all_links returns per link, urdf_interface.py:517-553). */
int dc_fk_tree_frames(const dc_fk_desc* fk, const void* q, int64_t batch, int32_t dtype, void* frames, dc_stream_t stream);

/*
 * Greedy kernel-perceptron training, the whole loop in one launch (DiffCo.train_perceptron,
 * kernel_perceptrons.py:98-133; legacy_multi != 0: MultiDiffCo.train_perceptron, deprecated/MultiDiffCo.py:50-83).
 * x_feat[N,F] are the transformed training points, y[N,C] the +-1 labels.  gains[N,C], hypothesis[N,C],
 * kernel_matrix[N,N] and diag[N] (the diagonal of kernel_matrix; 0 == "row not computed yet") are IN/OUT: zero them
 * for a fresh fit, or pre-load them for the jump-start update (kernel_perceptrons.py:222-269).
 * iterations_out[2] (device): index of the last iteration executed, number of kernel rows evaluated.
 */
int dc_perceptron_train(const dc_kernel_desc* kernel, const void* x_feat, const void* y, int64_t n, int32_t n_features,
                        int32_t n_class, int32_t dtype, double beta, int64_t max_iteration, void* gains, void* hypothesis,
                        void* kernel_matrix, void* diag, int32_t legacy_multi, int64_t* iterations_out, dc_stream_t stream);
/*
 * The same loop without the N x N kernel matrix (the reference zero-initialises one: 400 MB at N = 10 000,
 * kernel_perceptrons.py:90-96,204-220): only the rows the loop asks for are stored, in kernel_rows[row_capacity][N];
 * row_slot[N] (int32, IN/OUT, -1 = not computed; pre-loaded rows — the jump-start update — occupy slots 0, 1, ...) maps
 * a training point to its row.  When row_capacity is exhausted the loop stops with iterations_out[1] = -1 and nothing
 * else is meaningful: call again with a larger capacity (row_capacity = n can never run out).
 */
int dc_perceptron_train_rows(const dc_kernel_desc* kernel, const void* x_feat, const void* y, int64_t n, int32_t n_features,
                             int32_t n_class, int32_t dtype, double beta, int64_t max_iteration, void* gains, void* hypothesis,
                             void* kernel_rows, int32_t* row_slot, int64_t row_capacity, void* diag, int32_t legacy_multi,
                             int64_t* iterations_out, dc_stream_t stream);

/*
 * Multi-GPU (one process per GPU of one box).  dc_peer_alloc gives a zeroed device buffer plus a handle
 * other processes open with dc_peer_open (CUDA IPC over NVLink).  dc_score_grad_bcast is dc_score_grad with the fused
 * [score | grad] record of row b stored at row (row_offset + b) of EVERY buffer in `outs` (this rank's own and its peers'
 * mappings) — the all-gather happens inside the kernel's epilogue.  Returns DC_ERR_UNSUPPORTED when the call does not go
 * to the tensor-core kernel (use dc_score_grad + a collective then).  `q` may be a pinned host buffer (read zero-copy) and
 * `mirror` (optional, device or pinned host, [batch][record]) receives a second copy of THIS launch's records — the
 * host-buffers-in / host-buffers-out call of the sharded scorer is this one launch.  dc_peer_barrier (same stream, afterwards) publishes
 * the stores: flags->ptr[r] is rank r's flag array (>= world uint32, zero-initialised, e.g. the head of a dc_peer_alloc
 * buffer) as mapped in this process; `epoch` must increase by one per call.
 */
#define DC_MAX_PEERS 8
typedef struct dc_peer_handle { unsigned char bytes[64]; } dc_peer_handle;
typedef struct dc_peer_table { void* ptr[DC_MAX_PEERS]; } dc_peer_table;
int dc_peer_alloc(int64_t bytes, void** ptr, dc_peer_handle* handle);
int dc_peer_open(const dc_peer_handle* handle, void** ptr);
int dc_peer_close(void* ptr);
int dc_peer_free(void* ptr);
int dc_peer_barrier(const dc_peer_table* flags, int32_t rank, int32_t world, uint32_t epoch, dc_stream_t stream);
/* dc_score_grad_bcast + dc_peer_barrier in ONE launch: the last CTA of the grid to finish publishes `epoch` and waits
 * for every rank's — the step is complete (this rank's gathered buffer holds every rank's records) when the kernel ends.
 * Every rank must call it for every step with the same epoch sequence as dc_peer_barrier (the two may be mixed);
 * one such launch in flight per device. */
int dc_score_grad_bcast_sync(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                             int64_t batch, const dc_peer_table* outs, int32_t n_outs, int64_t row_offset, int32_t grad_mode,
                             void* mirror, const dc_peer_table* flags, int32_t rank, uint32_t epoch, dc_stream_t stream);
int dc_score_grad_bcast(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                        int64_t batch, const dc_peer_table* outs, int32_t n_outs, int64_t row_offset, int32_t grad_mode,
                        void* mirror, dc_stream_t stream);

/*
 * One iteration of the reference's penalty trajectory optimiser (Weighted.step, diffco/optim.py:706-752; the same terms
 * as adam_traj_optimize, optim.py:86-127) for W waypoints in one launch: control points, path length, max-move and
 * joint-limit hinges, the collision hinge on `score` (+ `score_grad` = d score/dp from dc_score_grad; both NULL = no
 * collision term), the analytic gradient of the weighted sum, the `mask` multipliers, torch.optim.Adam's update and
 * robot.wrap.  p[W,dof], exp_avg, exp_avg_sq are IN/OUT device arrays of `dtype`; `step` is the optimiser's step count
 * (device double, incremented); terms[6] receives {path length, collision, joint limit, max move, constraint loss,
 * squared norm of the masked gradient}.
 */
typedef struct dc_traj_params {
  double dif_weight, max_move_weight, collision_weight, joint_limit_weight; /* optim.py:19-22,669-676 */
  double safety_bias, max_speed;
  double lr, beta1, beta2, eps;          /* torch.optim.Adam */
  double limits[DC_MAX_DOF][2];          /* robot.limits */
  int32_t wrap[DC_MAX_DOF];              /* 1: wrap this coordinate to [-pi, pi) after the update (robot.wrap) */
} dc_traj_params;
int dc_traj_step(const dc_fk_desc* fk, const dc_traj_params* params, int64_t n_waypoints, int32_t dtype, void* p,
                 const void* score, const void* score_grad, const void* mask, void* exp_avg, void* exp_avg_sq, double* step,
                 void* terms, dc_stream_t stream);

/*
 * Dense collision checking and the device-side exit test of Weighted.step (diffco/optim.py:709-711,747-752).
 * dc_traj_dense_path is utils.dense_path (diffco/utils.py:87-102) on the device: per segment ceil(|dq| / max_step) points
 * q[i] + k max_step dq/|dq|, then the last waypoint, written to dense[max_points][dof] (rows beyond the point count are
 * filled with the last waypoint, so a scoring launch of the static size max_points is well defined); seg_offset[W]
 * receives the index of each segment's first point, *count the number of points (-1: more than max_points).
 * dc_traj_step_ex is dc_traj_step with (a) `dense` != NULL: the collision term is mean(hinge over the dense points) x W
 * and its gradient is chained through the interpolation to the waypoints (score / score_grad must then be NULL);
 * (b) `state` != NULL (device int32[2]): state[0] != 0 turns the launch (and dc_traj_dense_path) into a no-op, after
 * the update state[1] += 1 and state[0] = 1 once the constraint loss is <= exit_constraint — the reference's early
 * exit, so that a CUDA graph can be replayed several times between host read-backs without overshooting.
 */
typedef struct dc_traj_dense {
  const int32_t* count;
  const int32_t* seg_offset;
  const void* score;      /* [max_points]       dc_score_grad of the dense points */
  const void* score_grad; /* [max_points][dof] */
  int32_t max_points;
  int32_t reserved;
} dc_traj_dense;
int dc_traj_dense_path(const void* p, int64_t n_waypoints, int32_t dof, int32_t dtype, double max_step, int32_t max_points,
                       void* dense, int32_t* seg_offset, int32_t* count, const int32_t* state, dc_stream_t stream);
int dc_traj_step_ex(const dc_fk_desc* fk, const dc_traj_params* params, int64_t n_waypoints, int32_t dtype, void* p,
                    const void* score, const void* score_grad, const dc_traj_dense* dense, const void* mask, void* exp_avg,
                    void* exp_avg_sq, double* step, void* terms, int32_t* state, double exit_constraint, dc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFCO_B200_H */
