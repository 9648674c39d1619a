// extern "C" entry points of libdiffco_b200.so (see include/diffco_b200.h) and the small helper kernels
// (support-table packing, kernel matrix, stand-alone FK / FK-VJP).
#include <cstring>

#include "dc_common.cuh"
#include "dc_fk.cuh"
#include "dc_radial.cuh"
#include "dc_score_ls.cuh"  // LsArgs only; kernels are instantiated in dc_score_ls_f32.cu / _f64.cu
#include "dc_score_tq.cuh"

namespace dc {

long long g_launch_count = 0;
int g_last_score_kernel = -1;  // 0 lane-split, 1 thread-per-query, 2 tensor-core (dc_last_score_kernel)

// dc_score_tc.cu
bool takes_tensor_core_kernel(const dc_fk_desc& fk, const dc_kernel_desc& kernel, const dc_supports& sv, int64_t batch,
                              int grad_mode);
int tc_score_grad(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q, int64_t batch,
                  void* score, int64_t score_ld, void* grad, int64_t grad_ld, const void* grad_out, int32_t grad_mode,
                  int num_sms, cudaStream_t stream, const dc_peer_table* bcast = nullptr, int n_bcast = 0,
                  void* mirror = nullptr);

#define DC_TQ_DECL(name) int name(int n_feat, ScoreArgs<float>& a, int num_sms, cudaStream_t stream);
DC_TQ_DECL(tq_rq2_c1_score)
DC_TQ_DECL(tq_rq2_c1_grad)
DC_TQ_DECL(tq_rq2_c4_score)
DC_TQ_DECL(tq_rq2_c4_grad)
DC_TQ_DECL(tq_ph1_c1_score)
DC_TQ_DECL(tq_ph1_c1_grad)
DC_TQ_DECL(tq_ph1_c4_score)
DC_TQ_DECL(tq_ph1_c4_grad)
DC_TQ_DECL(tq_mq_c1_score)
DC_TQ_DECL(tq_mq_c1_grad)
DC_TQ_DECL(tq_mq_c4_score)
DC_TQ_DECL(tq_mq_c4_grad)
#undef DC_TQ_DECL

int ls_launch_f32(LsArgs<float>& a, int num_sms, cudaStream_t stream);
int ls_launch_f64(LsArgs<double>& a, int num_sms, cudaStream_t stream);
static int ls_launch(LsArgs<float>& a, int num_sms, cudaStream_t stream) { return ls_launch_f32(a, num_sms, stream); }
static int ls_launch(LsArgs<double>& a, int num_sms, cudaStream_t stream) { return ls_launch_f64(a, num_sms, stream); }

typedef int (*tq_fn)(int, ScoreArgs<float>&, int, cudaStream_t);
// [kind][cw4][mode]
static tq_fn const kTqTable[3][2][2] = {
    {{tq_rq2_c1_score, tq_rq2_c1_grad}, {tq_rq2_c4_score, tq_rq2_c4_grad}},
    {{tq_ph1_c1_score, tq_ph1_c1_grad}, {tq_ph1_c4_score, tq_ph1_c4_grad}},
    {{tq_mq_c1_score, tq_mq_c1_grad}, {tq_mq_c4_score, tq_mq_c4_grad}},
};

int device_sm_count(int* out) {
  static std::atomic<int> cached[64];  // per device ordinal; 0 = not queried yet (thread-safe: worst case two queries)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return DC_ERR_NO_DEVICE;
  }
  int sms = (dev >= 0 && dev < 64) ? cached[dev].load(std::memory_order_relaxed) : 0;
  if (sms == 0) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      (void)cudaGetLastError();
      return DC_ERR_NO_DEVICE;
    }
    if (dev >= 0 && dev < 64) cached[dev].store(sms, std::memory_order_relaxed);
  }
  *out = sms;
  return DC_OK;
}

static int fk_n_features(const dc_fk_desc& fk) { return fk.type == DC_FK_NONE ? fk.dof : fk.n_points * fk.point_dim; }

static bool fk_valid(const dc_fk_desc& fk) {
  if (fk.dof < 1) return false;
  const int F = fk_n_features(fk);
  if (F < 1 || F > DC_MAX_FEATURES) return false;
  // composite maps: the per-type rules below apply to ONE block of din columns / fin features
  const int rep = fk.n_repeat > 1 ? fk.n_repeat : 1, tl = fk.time_last ? 1 : 0;
  const bool composite = rep > 1 || tl;
  if (fk.type == DC_FK_NONE) return !composite;
  if (fk.dof > DC_MAX_DOF || (fk.dof - tl) % rep || (F - tl) % rep) return false;
  const int din = (fk.dof - tl) / rep, fin = (F - tl) / rep;
  if (din < 1 || fin < 1) return false;
  if (composite ? fk.point_dim != 1 : fk.point_dim < 1) return false;
  const auto points = [&](int dim, int m) { return fin == dim * m && (composite || (fk.point_dim == dim && fk.n_points == m)); };
  switch (fk.type) {
    case DC_FK_PLANAR_CHAIN:
      return fk.n_links == din && fk.n_links <= DC_MAX_LINKS && points(2, din);
    case DC_FK_SE2_BODY:
      return din == 3 && fk.n_keypoints <= DC_MAX_KEYPOINTS && points(2, fk.n_keypoints);
    case DC_FK_SE3_BODY:
      return din == 6 && fk.n_keypoints <= DC_MAX_KEYPOINTS && points(3, fk.n_keypoints);
    case DC_FK_SE2_BASE_PLANAR_ARM:
      return din == 3 + fk.n_links && fk.n_keypoints <= DC_MAX_KEYPOINTS && points(2, fk.n_keypoints + fk.n_links);
    case DC_FK_DH_ARMS: {
      if (fin % 3 || !points(3, fin / 3) || fk.n_arms < 1 || fk.n_arms > DC_MAX_ARMS) return false;
      const int m = fin / 3;
      for (int a = 0; a < fk.n_arms; ++a) {
        const dc_dh_arm& arm = fk.arms[a];
        if (arm.n_joints < 1 || arm.n_joints > DC_MAX_ARM_JOINTS || arm.n_tool < 0 || arm.n_tool > DC_MAX_TOOL_POINTS)
          return false;
        for (int i = 0; i < arm.n_joints; ++i) {
          if (arm.joint_index[i] < 0 || arm.joint_index[i] >= din) return false;
          if (arm.out_slot[i] >= m) return false;
        }
        for (int k = 0; k < arm.n_tool; ++k)
          if (arm.tool_slot[k] < 0 || arm.tool_slot[k] >= m) return false;
      }
      return true;
    }
    case DC_FK_JOINT_TREE: {
      if (fin % 3 || !points(3, fin / 3) || fk.n_nodes < 1 || fk.n_nodes > DC_MAX_TREE_NODES) return false;
      for (int i = 0; i < fk.n_nodes; ++i) {
        const dc_tree_node& nd = fk.tree[i];
        if (nd.parent >= i || nd.parent < -1 || nd.q_index >= din || nd.q_index < -1) return false;
        if (nd.joint < DC_JOINT_FIXED || nd.joint > DC_JOINT_PRISMATIC || nd.out_slot >= fin / 3 || nd.out_slot < -1) return false;
        if (nd.joint == DC_JOINT_FIXED && nd.q_index >= 0) return false;
      }
      return true;
    }
    default:
      return false;
  }
}

// ---------------------------------------------------------------------------------------------------
// helper kernels
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_supports_kernel(const T* __restrict__ s, const T* __restrict__ w, long long n, int F, int C, int f_pad,
                                     int row, T* __restrict__ table) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * row) return;
  const long long r = idx / row;
  const int e = (int)(idx - r * row);
  T v = (T)0;
  if (e < F)
    v = -s[r * F + e];
  else if (e >= f_pad && e < f_pad + C)
    v = w[r * C + (e - f_pad)];
  table[idx] = v;
}

// K[i,j] = k(|xa_i - xb_j|^2): one CTA per (row i, 128-column block); xa_i staged in shared memory.
template <typename T>
__global__ void __launch_bounds__(128) kernel_matrix_kernel(RadialConsts<T> rc, const T* __restrict__ xa, long long na,
                                                            const T* __restrict__ xb, long long nb, int F,
                                                            T* __restrict__ out) {
  __shared__ T xi[DC_MAX_FEATURES];
  const long long i = blockIdx.y;
  for (int f = threadIdx.x; f < F; f += blockDim.x) xi[f] = xa[i * F + f];
  __syncthreads();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nb) return;
  const T* xj = xb + j * F;
  const bool temporal = rc.kind == DC_K_RQ_TEMPORAL;
  const int Fx = temporal ? F - 1 : F;
  T rho = (T)0;
  for (int f = 0; f < Fx; ++f) {
    const T d = xi[f] - xj[f];
    rho = fma(d, d, rho);
  }
  T k, coef, coef_t;
  if (temporal)
    radial_eval_temporal<T>(rc, rho, xi[F - 1] - xj[F - 1], k, coef, coef_t);
  else
    radial_eval<KR_GENERIC, T>(rc, rho, k, coef);
  out[i * nb + j] = k * rc.score_scale;
}

template <typename T>
__global__ void __launch_bounds__(128) fk_forward_kernel(const __grid_constant__ dc_fk_desc fk, const T* __restrict__ q,
                                                         long long batch, T* __restrict__ x) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int F = fk.type == DC_FK_NONE ? fk.dof : fk.n_points * fk.point_dim;
  if (fk.type == DC_FK_NONE) {
    for (int f = 0; f < F; ++f) x[b * F + f] = q[b * F + f];
    return;
  }
  T qv[DC_MAX_DOF];
#pragma unroll
  for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < fk.dof) ? q[b * fk.dof + i] : (T)0;
  fk_features(fk, qv, x + b * F, 1);  // float32: evaluated in float64 and rounded once (dc_fk.cuh)
}

// [R | t] of every body of a joint tree (float32: evaluated in float64, rounded once, like the features)
template <typename T>
__global__ void __launch_bounds__(128) fk_tree_frames_kernel(const __grid_constant__ dc_fk_desc fk, const T* __restrict__ q,
                                                             long long batch, T* __restrict__ out) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double qv[DC_MAX_DOF];
#pragma unroll
  for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < fk.dof) ? (double)q[b * fk.dof + i] : 0.0;
  double fr[DC_MAX_TREE_NODES][12];
  fk_tree_frames<double>(fk, qv, fr);
  T* o = out + b * fk.n_nodes * 12;
  for (int i = 0; i < fk.n_nodes; ++i)
    for (int e = 0; e < 12; ++e) o[i * 12 + e] = (T)fr[i][e];
}

// float32 features as (hi, lo) pairs: x = hi + lo to ~1e-9 (hi is what fk_forward_kernel<float> returns)
__global__ void __launch_bounds__(128) fk_forward_split_kernel(const __grid_constant__ dc_fk_desc fk, const float* __restrict__ q,
                                                               long long batch, float* __restrict__ xh, float* __restrict__ xl) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int F = fk.type == DC_FK_NONE ? fk.dof : fk.n_points * fk.point_dim;
  if (fk.type == DC_FK_NONE) {
    for (int f = 0; f < F; ++f) {
      xh[b * F + f] = q[b * F + f];
      xl[b * F + f] = 0.f;
    }
    return;
  }
  float qv[DC_MAX_DOF];
#pragma unroll
  for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < fk.dof) ? q[b * fk.dof + i] : 0.f;
  fk_forward_f32x<true>(fk, qv, xh + b * F, 1, xl + b * F, 1);
}

// table_lo[n][row] = -s_lo[n][f] for f < F, zero elsewhere (same row layout as the packed support table)
__global__ void pack_supports_lo_kernel(const float* __restrict__ s_lo, long long n, int F, int row, float* __restrict__ table_lo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * row) return;
  const long long r = idx / row;
  const int e = (int)(idx - r * row);
  table_lo[idx] = (e < F) ? -s_lo[r * F + e] : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(128) fk_vjp_kernel(const __grid_constant__ dc_fk_desc fk, const T* __restrict__ q,
                                                     long long batch, const T* __restrict__ gx, T* __restrict__ gq) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int F = fk.type == DC_FK_NONE ? fk.dof : fk.n_points * fk.point_dim;
  if (fk.type == DC_FK_NONE) {
    for (int f = 0; f < F; ++f) gq[b * F + f] = gx[b * F + f];
    return;
  }
  T qv[DC_MAX_DOF], out[DC_MAX_DOF], xl[DC_MAX_FEATURES], gl[DC_MAX_FEATURES];
#pragma unroll
  for (int i = 0; i < DC_MAX_DOF; ++i) {
    qv[i] = (i < fk.dof) ? q[b * fk.dof + i] : (T)0;
    out[i] = (T)0;
  }
  for (int f = 0; f < F; ++f) gl[f] = gx[b * F + f];
  fk_features(fk, qv, xl, 1);
  fk_vjp<T>(fk, qv, xl, 1, gl, 1, out);
  for (int i = 0; i < fk.dof; ++i) gq[b * fk.dof + i] = out[i];
}

// ---------------------------------------------------------------------------------------------------
// typed implementations
// ---------------------------------------------------------------------------------------------------
template <typename T>
static int score_grad_generic(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                              int64_t batch, void* score, int64_t score_ld, void* grad, int64_t grad_ld,
                              const void* grad_out, int32_t grad_mode, int num_sms, cudaStream_t stream) {
  LsArgs<T> a;
  a.fk = *fk;
  if (!make_radial_consts<T>(*kernel, &a.rc)) return DC_ERR_INVALID_ARG;
  a.table = static_cast<const T*>(sv->table);
  a.q = static_cast<const T*>(q);
  a.score = static_cast<T*>(score);
  a.grad = static_cast<T*>(grad);
  a.grad_out = static_cast<const T*>(grad_out);
  a.batch = batch;
  a.score_ld = score_ld;
  a.grad_ld = grad_ld;
  a.n_sv = (int)sv->n;
  a.n_feat = sv->n_features;
  a.n_class = sv->n_class;
  a.n_in = fk->dof;
  a.f_pad = sv->f_pad;
  a.row_stride = sv->row_stride;
  a.grad_mode = grad_mode;
  a.wpq = 1;
  return ls_launch(a, num_sms, stream);
}

// Batches below this go to the lane-split kernel (too few 64-query tiles to occupy the SMs).
static constexpr int64_t kTqMinBatch = 2048;

static bool tq_feature_count(int f) {  // keep in sync with the switch in dc_score_tq_inst.cu
  switch (f) {
    case 2: case 3: case 4: case 6: case 7: case 8: case 10: case 12: case 14: case 15: case 16: case 21: case 24: case 27:
      return true;
    default:
      return false;
  }
}

// Will dc_score_grad take the thread-per-query kernel (tile I/O staged through shared memory, coalesced)?  Used by the
// host-buffer entry point to decide between zero-copy access to pinned memory and staged copies.
bool takes_thread_per_query_kernel(const dc_fk_desc& fk, const dc_kernel_desc& kernel, const dc_supports& sv, int64_t batch) {
  if (takes_tensor_core_kernel(fk, kernel, sv, batch, DC_GRAD_SUM)) return true;  // same staged, coalesced tile I/O
  return sv.dtype == DC_F32 && fast_radial_kind(kernel) != KR_GENERIC && sv.n_class <= 4 && batch >= kTqMinBatch &&
         tq_feature_count(fk_n_features(fk));
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_abi_version(void) { return DC_ABI_VERSION; }

const char* dc_status_string(int status) {
  switch (status) {
    case DC_OK: return "ok";
    case DC_ERR_INVALID_ARG: return "invalid argument";
    case DC_ERR_UNSUPPORTED: return "unsupported configuration";
    case DC_ERR_CUDA: return "CUDA error";
    case DC_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

int64_t dc_launch_count(void) { return g_launch_count; }

int dc_last_score_kernel(void) { return g_last_score_kernel; }

int dc_supports_layout(int32_t n_features, int32_t n_class, int32_t dtype, int32_t* f_pad, int32_t* row_stride) {
  if (n_features < 1 || n_features > DC_MAX_FEATURES || n_class < 1 || n_class > DC_MAX_CLASSES) return DC_ERR_INVALID_ARG;
  if (dtype != DC_F32 && dtype != DC_F64) return DC_ERR_INVALID_ARG;
  const int fp = 2 * ceil_div(n_features, 2);
  // classes are padded to the width the thread-per-query kernel is instantiated for (1 or 4) with zero weights
  const int cp = (n_class == 1) ? 1 : round_up(n_class, 4);
  if (f_pad) *f_pad = fp;
  if (row_stride) *row_stride = round_up(fp + cp, 4);
  return DC_OK;
}

int dc_pack_supports(const void* s_feat, const void* w, int64_t n, int32_t n_features, int32_t n_class, int32_t dtype,
                     void* table, dc_stream_t stream) {
  int32_t f_pad = 0, row = 0;
  const int st = dc_supports_layout(n_features, n_class, dtype, &f_pad, &row);
  if (st != DC_OK) return st;
  if (n < 1 || !s_feat || !w || !table) return DC_ERR_INVALID_ARG;
  const long long total = (long long)n * row;
  const int threads = 256;
  const long long blocks = ceil_div64(total, threads);
  if (dtype == DC_F32)
    pack_supports_kernel<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const float*)s_feat, (const float*)w, n, n_features, n_class, f_pad, row, (float*)table);
  else
    pack_supports_kernel<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const double*)s_feat, (const double*)w, n, n_features, n_class, f_pad, row, (double*)table);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_score_grad(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                  int64_t batch, void* score, int64_t score_ld, void* grad, int64_t grad_ld, const void* grad_out,
                  int32_t grad_mode, dc_stream_t stream) {
  if (!fk || !kernel || !sv || !fk_valid(*fk)) return DC_ERR_INVALID_ARG;
  if (batch == 0) return DC_OK;
  if (batch < 0 || !q || !score || !sv->table || sv->n < 1 || sv->n > 0x7fffffff) return DC_ERR_INVALID_ARG;
  if (grad_mode != DC_GRAD_NONE && grad_mode != DC_GRAD_SUM && grad_mode != DC_GRAD_JAC) return DC_ERR_INVALID_ARG;
  if (grad_mode != DC_GRAD_NONE && !grad) return DC_ERR_INVALID_ARG;
  const int F = fk_n_features(*fk);
  int32_t f_pad = 0, row = 0;
  if (dc_supports_layout(sv->n_features, sv->n_class, sv->dtype, &f_pad, &row) != DC_OK) return DC_ERR_INVALID_ARG;
  if (F != sv->n_features || f_pad != sv->f_pad || row != sv->row_stride) return DC_ERR_INVALID_ARG;
  const int64_t grad_cols = (grad_mode == DC_GRAD_JAC) ? (int64_t)sv->n_class * fk->dof : fk->dof;
  if (score_ld == 0) score_ld = sv->n_class;
  if (grad_ld == 0) grad_ld = grad_cols;
  if (score_ld < sv->n_class || (grad_mode != DC_GRAD_NONE && grad_ld < grad_cols)) return DC_ERR_INVALID_ARG;
  int num_sms = 0;
  const int st = device_sm_count(&num_sms);
  if (st != DC_OK) return st;
  cudaStream_t cs = (cudaStream_t)stream;
  if (grad_mode == DC_GRAD_JAC && sv->n_class == 1) {
    // one class: the Jacobian [B,1,D] IS the summed gradient with a unit upstream gradient — lets the autograd path
    // (functional._ScoreFunction asks for Jacobians) use the large-batch kernels, tensor cores included
    grad_mode = DC_GRAD_SUM;
    grad_out = nullptr;
  }

  if (sv->dtype == DC_F64) {
    g_last_score_kernel = 0;
    return score_grad_generic<double>(fk, kernel, sv, q, batch, score, score_ld, grad, grad_ld, grad_out, grad_mode, num_sms,
                                      cs);
  }

  // fp32, RQ(p = 2), one class, F <= 14, large batch, packed operand image present: tensor-core kernel
  if (takes_tensor_core_kernel(*fk, *kernel, *sv, batch, grad_mode)) {
    const int r = tc_score_grad(fk, kernel, sv, q, batch, score, score_ld, grad, grad_ld, grad_out, grad_mode, num_sms, cs);
    if (r != DC_ERR_UNSUPPORTED) {
      g_last_score_kernel = 2;
      return r;
    }
  }

  // fp32: thread-per-query kernel for large batches of the instantiated shapes, lane-split kernel otherwise
  const int kind = fast_radial_kind(*kernel);
  const int C = sv->n_class;
  if (kind != KR_GENERIC && C <= 4 && batch >= kTqMinBatch) {
    tq_fn fn = kTqTable[kind][C == 1 ? 0 : 1][grad_mode == DC_GRAD_NONE ? M_SCORE : M_GRAD];
    ScoreArgs<float> a;
    a.fk = *fk;
    if (!make_radial_consts<float>(*kernel, &a.rc)) return DC_ERR_INVALID_ARG;
    a.table = (const float*)sv->table;
    a.q = (const float*)q;
    a.score = (float*)score;
    a.grad = (float*)grad;
    a.grad_out = (grad_mode == DC_GRAD_SUM) ? (const float*)grad_out : nullptr;
    a.batch = batch;
    a.score_ld = score_ld;
    a.grad_ld = grad_ld;
    a.n_sv = (int)sv->n;
    a.n_feat = F;
    a.n_class = C;
    a.n_in = fk->dof;
    a.jac_class = -1;
    a.write_score = 1;
    a.n_tiles = 0;
    a.st_q = 0;
    a.chunk_rows = 0;
    int r;
    if (grad_mode == DC_GRAD_JAC && C > 1) {
      // Jacobian at large batch: one gradient pass per class (one-hot upstream gradient); the first pass writes the
      // scores.  The optimisers' Jacobian calls are small-batch and take the lane-split kernel instead.
      r = DC_OK;
      for (int c = 0; c < C && r == DC_OK; ++c) {
        a.jac_class = c;
        a.write_score = (c == 0);
        r = fn(F, a, num_sms, cs);
      }
    } else {
      r = fn(F, a, num_sms, cs);
    }
    if (r != DC_ERR_UNSUPPORTED) {
      g_last_score_kernel = 1;
      return r;
    }
  }
  g_last_score_kernel = 0;
  return score_grad_generic<float>(fk, kernel, sv, q, batch, score, score_ld, grad, grad_ld, grad_out, grad_mode, num_sms, cs);
}

int dc_kernel_matrix(const dc_kernel_desc* kernel, const void* xa, int64_t na, const void* xb, int64_t nb,
                     int32_t n_features, int32_t dtype, void* k_out, dc_stream_t stream) {
  if (!kernel || n_features < 1 || n_features > DC_MAX_FEATURES) return DC_ERR_INVALID_ARG;
  if (na == 0 || nb == 0) return DC_OK;
  if (na < 0 || nb < 0 || !xa || !xb || !k_out || na > 65535LL * 65535LL) return DC_ERR_INVALID_ARG;
  cudaStream_t cs = (cudaStream_t)stream;
  const long long bx = ceil_div64(nb, 128);
  // gridDim.y is limited to 65535: walk the rows in slabs
  for (long long i0 = 0; i0 < na; i0 += 65535) {
    const long long rows = (na - i0 < 65535) ? (na - i0) : 65535;
    dim3 grid((unsigned)bx, (unsigned)rows);
    if (dtype == DC_F32) {
      RadialConsts<float> rc;
      if (!make_radial_consts<float>(*kernel, &rc)) return DC_ERR_INVALID_ARG;
      kernel_matrix_kernel<float><<<grid, 128, 0, cs>>>(rc, (const float*)xa + i0 * n_features, rows, (const float*)xb, nb,
                                                       n_features, (float*)k_out + i0 * nb);
    } else if (dtype == DC_F64) {
      RadialConsts<double> rc;
      if (!make_radial_consts<double>(*kernel, &rc)) return DC_ERR_INVALID_ARG;
      kernel_matrix_kernel<double><<<grid, 128, 0, cs>>>(rc, (const double*)xa + i0 * n_features, rows, (const double*)xb,
                                                        nb, n_features, (double*)k_out + i0 * nb);
    } else {
      return DC_ERR_INVALID_ARG;
    }
    DC_LAUNCH_CHECK();
  }
  return DC_OK;
}

int dc_fk_forward(const dc_fk_desc* fk, const void* q, int64_t batch, int32_t dtype, void* x_out, dc_stream_t stream) {
  if (!fk || !fk_valid(*fk)) return DC_ERR_INVALID_ARG;
  if (batch == 0) return DC_OK;
  if (batch < 0 || !q || !x_out) return DC_ERR_INVALID_ARG;
  const long long blocks = ceil_div64(batch, 128);
  if (dtype == DC_F32)
    fk_forward_kernel<float><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*fk, (const float*)q, batch, (float*)x_out);
  else if (dtype == DC_F64)
    fk_forward_kernel<double><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*fk, (const double*)q, batch, (double*)x_out);
  else
    return DC_ERR_INVALID_ARG;
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_fk_forward_split(const dc_fk_desc* fk, const void* q, int64_t batch, void* x_hi, void* x_lo, dc_stream_t stream) {
  if (!fk || !fk_valid(*fk)) return DC_ERR_INVALID_ARG;
  if (batch == 0) return DC_OK;
  if (batch < 0 || !q || !x_hi || !x_lo) return DC_ERR_INVALID_ARG;
  const long long blocks = ceil_div64(batch, 128);
  fk_forward_split_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*fk, (const float*)q, batch, (float*)x_hi, (float*)x_lo);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_pack_supports_lo(const void* s_lo, int64_t n, int32_t n_features, int32_t n_class, void* table_lo, dc_stream_t stream) {
  int32_t f_pad = 0, row = 0;
  const int st = dc_supports_layout(n_features, n_class, DC_F32, &f_pad, &row);
  if (st != DC_OK) return st;
  if (n < 1 || !s_lo || !table_lo) return DC_ERR_INVALID_ARG;
  const long long total = (long long)n * row;
  pack_supports_lo_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)s_lo, n, n_features, row,
                                                                                        (float*)table_lo);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_fk_tree_frames(const dc_fk_desc* fk, const void* q, int64_t batch, int32_t dtype, void* frames, dc_stream_t stream) {
  if (!fk || !fk_valid(*fk) || fk->type != DC_FK_JOINT_TREE || fk->n_repeat > 1 || fk->time_last) return DC_ERR_INVALID_ARG;
  if (batch < 0 || (dtype != DC_F32 && dtype != DC_F64)) return DC_ERR_INVALID_ARG;
  if (batch == 0) return DC_OK;
  if (!q || !frames) return DC_ERR_INVALID_ARG;
  const unsigned grid = (unsigned)((batch + 127) / 128);
  if (dtype == DC_F32)
    fk_tree_frames_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(*fk, (const float*)q, batch, (float*)frames);
  else
    fk_tree_frames_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(*fk, (const double*)q, batch, (double*)frames);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_fk_vjp(const dc_fk_desc* fk, const void* q, int64_t batch, int32_t dtype, const void* g_x, void* g_q,
              dc_stream_t stream) {
  if (!fk || !fk_valid(*fk)) return DC_ERR_INVALID_ARG;
  if (batch == 0) return DC_OK;
  if (batch < 0 || !q || !g_x || !g_q) return DC_ERR_INVALID_ARG;
  const long long blocks = ceil_div64(batch, 128);
  if (dtype == DC_F32)
    fk_vjp_kernel<float><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*fk, (const float*)q, batch, (const float*)g_x,
                                                                            (float*)g_q);
  else if (dtype == DC_F64)
    fk_vjp_kernel<double><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*fk, (const double*)q, batch,
                                                                             (const double*)g_x, (double*)g_q);
  else
    return DC_ERR_INVALID_ARG;
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
