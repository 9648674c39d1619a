// Tensor-core score kernel (dc_score_tc.cuh): instantiations, support-set packing entry points and the dispatch test.
#include <cmath>
#include <cstdlib>

#include "dc_score_tc.cuh"

namespace dc {

extern double g_peer_timeout_s;  // dc_peer.cu

// process-wide knobs (dc_set_option)
static double g_tc_enable = -1.0;  // -1: not initialised (reads DIFFCO_B200_TC once)
static double g_tc_err_coef = 5e-7;
static double g_tc_tol_pair = 2e-7;
static double g_tc_min_batch = 4096;
static unsigned long long* g_tc_stats = nullptr;  // device counter of near pairs (DC_OPT_TC_STATS), lazily allocated

static bool tc_enabled() {
  if (g_tc_enable < 0) {
    const char* e = std::getenv("DIFFCO_B200_TC");
    g_tc_enable = (e != nullptr && e[0] == '0') ? 0.0 : 1.0;
  }
  return g_tc_enable > 0;
}

int tc_set_option(int option, double value) {
  switch (option) {
    case DC_OPT_TC_ENABLE: g_tc_enable = value != 0.0 ? 1.0 : 0.0; return DC_OK;
    case DC_OPT_TC_ERR_COEF: if (!(value > 0)) return DC_ERR_INVALID_ARG; g_tc_err_coef = value; return DC_OK;
    case DC_OPT_TC_TOL_PAIR: if (!(value > 0)) return DC_ERR_INVALID_ARG; g_tc_tol_pair = value; return DC_OK;
    case DC_OPT_TC_MIN_BATCH: if (!(value >= 1)) return DC_ERR_INVALID_ARG; g_tc_min_batch = value; return DC_OK;
    case DC_OPT_PEER_TIMEOUT_S: if (!(value > 0)) return DC_ERR_INVALID_ARG; g_peer_timeout_s = value; return DC_OK;
    case DC_OPT_TC_STATS: {
      // value != 0: start (or reset) counting the pairs the tensor-core kernel re-evaluates exactly; 0: stop
      if (value == 0.0) {
        if (g_tc_stats) (void)cudaFree(g_tc_stats);
        g_tc_stats = nullptr;
        return DC_OK;
      }
      if (!g_tc_stats && cudaMalloc(&g_tc_stats, sizeof(unsigned long long)) != cudaSuccess) {
        (void)cudaGetLastError();
        g_tc_stats = nullptr;
        return DC_ERR_CUDA;
      }
      return cudaMemset(g_tc_stats, 0, sizeof(unsigned long long)) == cudaSuccess ? DC_OK : DC_ERR_CUDA;
    }
    default: return DC_ERR_INVALID_ARG;
  }
}
double tc_get_option(int option) {
  switch (option) {
    case DC_OPT_TC_ENABLE: return tc_enabled() ? 1.0 : 0.0;
    case DC_OPT_TC_ERR_COEF: return g_tc_err_coef;
    case DC_OPT_TC_TOL_PAIR: return g_tc_tol_pair;
    case DC_OPT_TC_MIN_BATCH: return g_tc_min_batch;
    case DC_OPT_PEER_TIMEOUT_S: return g_peer_timeout_s;
    case DC_OPT_TC_STATS: {  // synchronises the device: a debugging / test facility
      if (!g_tc_stats) return -1.0;
      unsigned long long v = 0;
      if (cudaMemcpy(&v, g_tc_stats, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
        (void)cudaGetLastError();
        return NAN;
      }
      return (double)v;
    }
    default: return NAN;
  }
}

static bool tc_shape_ok(int n_features, int n_class, int dtype) {
  return dtype == DC_F32 && n_class == 1 && n_features >= 1 && n_features <= TcLayoutT<32>::FMAX;
}

// Does dc_score_grad send this call to the tensor-core kernel?  (DiffCo.score with RQKernel(p = 2), one class,
// F <= 30 (<= 14: two CTAs per SM, the BASELINE shape; 15..30: one CTA per SM), fp32, score or score + summed gradient, a batch large enough to fill the SMs, and — when the caller told us
// max|s|^2 — a kernel narrow enough that only a small fraction of the pairs falls under the near-pair threshold.)
bool takes_tensor_core_kernel(const dc_fk_desc& fk, const dc_kernel_desc& kernel, const dc_supports& sv, int64_t batch,
                              int grad_mode) {
  if (!tc_enabled() || sv.tc_blob == nullptr) return false;
  const int F = fk.type == DC_FK_NONE ? fk.dof : fk.n_points * fk.point_dim;
  if (!tc_shape_ok(F, sv.n_class, sv.dtype) || (fk.type != DC_FK_NONE && fk.dof > DC_MAX_DOF)) return false;  // NONE: dof == F
  if (fk.n_repeat > 1 || fk.time_last) return false;  // composite maps (line / temporal kernels): lane-split or thread-per-query
  if (kernel.kind != DC_K_RQ || kernel.order != 2 || !(kernel.param > 0)) return false;
  if ((float)kernel.param != (float)sv.tc_gamma) return false;  // the operand image has the kernel width folded in
  if (fk.type != DC_FK_NONE && fk.dof > (tc_group(F) == 16 ? TcLayoutT<16>::QS_DOF : TcLayoutT<32>::QS_DOF)) return false;
  if (grad_mode != DC_GRAD_NONE && grad_mode != DC_GRAD_SUM) return false;
  if (batch < (int64_t)g_tc_min_batch || sv.n >= (1 << 24)) return false;
  if (sv.tc_s2max > 0) {
    // near threshold on rho for a typical query (|x|^2 ~ max|s|^2 / 2) against the spread of the supports
    const double drho = g_tc_err_coef * 1.5 * sv.tc_s2max;
    const double tcrit = std::cbrt(std::fmax(kernel.param * drho / g_tc_tol_pair, 1.0));
    const double thr = (tcrit - 1.0) / (kernel.param / 2.0);
    // (measured for the wide instantiation too, bench.py configs.wide_tc: Panda at gamma = 10 has 11 % near pairs and the
    // exact path then costs more than the whole lane-split kernel — 1.6e7 vs 3.4e7 evals/s — so the same limit applies)
    if (thr > 0.04 * sv.tc_s2max) return false;
  }
  return true;
}

// Device-visible alias of a pinned (page-locked, mapped) host pointer; device pointers are returned unchanged; nullptr for
// pageable host memory.
static void* device_alias(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  if (at.type == cudaMemoryTypeHost) return at.devicePointer;
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return const_cast<void*>(p);
  return nullptr;
}

// device counter of finished CTAs for the in-kernel step barrier (one sync launch at a time per device: PeerExchange
// issues them on one stream); the kernel leaves it at 0
__device__ unsigned int g_tc_done_counter = 0;
extern double g_peer_timeout_s;  // dc_peer.cu (DC_OPT_PEER_TIMEOUT_S)

static int tc_score_grad_sync(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                              int64_t batch, void* score, int64_t score_ld, void* grad, int64_t grad_ld, const void* grad_out,
                              int32_t grad_mode, int num_sms, cudaStream_t stream, const dc_peer_table* bcast, int n_bcast,
                              void* mirror, const dc_peer_table* flags, int rank, uint32_t epoch) {
  TcArgs a;
  a.sync_world = 0;
  a.flag_mine = nullptr;
  a.done = nullptr;
  a.epoch = epoch;
  a.sync_timeout_cycles = (long long)(g_peer_timeout_s * 2.0e9);
  for (int k = 0; k < DC_MAX_PEERS; ++k) a.flag_peer[k] = nullptr;
  if (flags != nullptr) {
    void* cnt = nullptr;
    if (cudaGetSymbolAddress(&cnt, g_tc_done_counter) != cudaSuccess) {
      (void)cudaGetLastError();
      return DC_ERR_CUDA;
    }
    a.done = static_cast<unsigned int*>(cnt);
    a.sync_world = n_bcast;
    a.flag_mine = static_cast<const uint32_t*>(flags->ptr[rank]);
    for (int k = 0; k < n_bcast; ++k) a.flag_peer[k] = static_cast<uint32_t*>(flags->ptr[k]) + rank;
  }
  a.mirror = static_cast<float*>(mirror);
  a.n_bcast = n_bcast;
  for (int k = 0; k < DC_MAX_PEERS; ++k) a.bcast[k] = (bcast && k < n_bcast) ? static_cast<float*>(bcast->ptr[k]) : nullptr;
  a.fk = *fk;
  if (!make_radial_consts<float>(*kernel, &a.rc)) return DC_ERR_INVALID_ARG;
  a.blob = static_cast<const unsigned char*>(sv->tc_blob);
  a.table = static_cast<const float*>(sv->table);
  a.table_lo = static_cast<const float*>(sv->table_lo);
  a.q = static_cast<const float*>(q);
  a.score = static_cast<float*>(score);
  a.grad = static_cast<float*>(grad);
  a.grad_out = (grad_mode == DC_GRAD_SUM) ? static_cast<const float*>(grad_out) : nullptr;
  a.trace = nullptr;
  a.dbg = nullptr;
  a.stats = g_tc_stats;
  a.batch = batch;
  a.score_ld = score_ld;
  a.grad_ld = grad_ld;
  a.n_sv = (int)sv->n;
  a.n_feat = sv->n_features;
  a.n_in = fk->dof;
  a.row_stride = sv->row_stride;
  a.f_pad = sv->f_pad;
  a.n_tiles = 0;
  a.n_chunks = 0;
  a.err_coef = (float)g_tc_err_coef;
  a.tol_pair = (float)g_tc_tol_pair;
  return grad_mode == DC_GRAD_NONE ? launch_score_tc<TC_SCORE>(a, num_sms, stream)
                                   : launch_score_tc<TC_GRAD>(a, num_sms, stream);
}

int tc_score_grad(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q, int64_t batch,
                  void* score, int64_t score_ld, void* grad, int64_t grad_ld, const void* grad_out, int32_t grad_mode,
                  int num_sms, cudaStream_t stream, const dc_peer_table* bcast, int n_bcast, void* mirror) {
  return tc_score_grad_sync(fk, kernel, sv, q, batch, score, score_ld, grad, grad_ld, grad_out, grad_mode, num_sms, stream, bcast,
                            n_bcast, mirror, nullptr, 0, 0);
}

}  // namespace dc

using namespace dc;

extern "C" {

int dc_supports_tc_bytes(int64_t n, int32_t n_features, int32_t n_class, int32_t dtype, int64_t* bytes) {
  if (n < 1 || !bytes) return DC_ERR_INVALID_ARG;
  if (!tc_shape_ok(n_features, n_class, dtype) || n >= (1 << 24)) return DC_ERR_UNSUPPORTED;
  *bytes = (int64_t)tc_blob_bytes(n, tc_group(n_features));
  return DC_OK;
}

int dc_pack_supports_tc(const void* s_feat, const void* w, int64_t n, int32_t n_features, const dc_kernel_desc* kernel,
                        void* blob, dc_stream_t stream) {
  if (n < 1 || !s_feat || !w || !blob || !kernel || (reinterpret_cast<uintptr_t>(blob) & 127) != 0) return DC_ERR_INVALID_ARG;
  if (!tc_shape_ok(n_features, 1, DC_F32) || n >= (1 << 24)) return DC_ERR_UNSUPPORTED;
  if (kernel->kind != DC_K_RQ || kernel->order != 2 || !(kernel->param > 0)) return DC_ERR_UNSUPPORTED;
  return launch_pack_supports_tc(static_cast<const float*>(s_feat), static_cast<const float*>(w), n, n_features,
                                 (float)kernel->param, static_cast<unsigned char*>(blob), (cudaStream_t)stream);
}

int dc_supports_tc_info(const void* blob, int64_t n, int32_t n_features, double* s2max, int32_t* valid) {
  if (!blob || n < 1 || n >= (1 << 24) || !tc_shape_ok(n_features, 1, DC_F32)) return DC_ERR_INVALID_ARG;
  float t[TcLayout::TRAILER_FLOATS];
  if (cudaMemcpy(t, tc_trailer(blob, n, tc_group(n_features)), sizeof(t), cudaMemcpyDeviceToHost) != cudaSuccess) {  // synchronises (pack time)
    (void)cudaGetLastError();
    return DC_ERR_CUDA;
  }
  if (s2max) *s2max = (double)t[0];
  if (valid) *valid = t[8] != 0.f ? 1 : 0;
  return DC_OK;
}

int dc_score_grad_bcast(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                        int64_t batch, const dc_peer_table* outs, int32_t n_outs, int64_t row_offset, int32_t grad_mode,
                        void* mirror, dc_stream_t stream) {
  if (!fk || !kernel || !sv || !outs || n_outs < 1 || n_outs > DC_MAX_PEERS || row_offset < 0) return DC_ERR_INVALID_ARG;
  if (batch == 0) return DC_OK;
  if (batch < 0 || !q || !sv->table) return DC_ERR_INVALID_ARG;
  if (grad_mode != DC_GRAD_NONE && grad_mode != DC_GRAD_SUM) return DC_ERR_INVALID_ARG;
  for (int k = 0; k < n_outs; ++k)
    if (!outs->ptr[k] || (reinterpret_cast<uintptr_t>(outs->ptr[k]) & 15) != 0) return DC_ERR_INVALID_ARG;
  if (!takes_tensor_core_kernel(*fk, *kernel, *sv, batch, grad_mode)) return DC_ERR_UNSUPPORTED;
  int dev = 0, num_sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return DC_ERR_NO_DEVICE;
  }
  // q and mirror may be pinned host buffers (zero-copy: the kernel moves whole tiles with coalesced 16-byte accesses)
  const void* qd = device_alias(q);
  void* md = mirror ? device_alias(mirror) : nullptr;
  if (!qd || (mirror && !md)) return DC_ERR_INVALID_ARG;
  const int64_t rec = 1 + (grad_mode == DC_GRAD_SUM ? fk->dof : 0);
  float* mine = static_cast<float*>(outs->ptr[0]) + row_offset * rec;  // the kernel adds (mine - outs[0]) to every base
  return tc_score_grad(fk, kernel, sv, qd, batch, mine, rec, grad_mode == DC_GRAD_SUM ? mine + 1 : nullptr, rec, nullptr,
                       grad_mode, num_sms, (cudaStream_t)stream, outs, n_outs, md);
}

int dc_score_grad_bcast_sync(const dc_fk_desc* fk, const dc_kernel_desc* kernel, const dc_supports* sv, const void* q,
                             int64_t batch, const dc_peer_table* outs, int32_t n_outs, int64_t row_offset, int32_t grad_mode,
                             void* mirror, const dc_peer_table* flags, int32_t rank, uint32_t epoch, dc_stream_t stream) {
  if (!fk || !kernel || !sv || !outs || !flags || n_outs < 1 || n_outs > DC_MAX_PEERS || row_offset < 0 || rank < 0 ||
      rank >= n_outs)
    return DC_ERR_INVALID_ARG;
  if (batch <= 0 || !q || !sv->table) return DC_ERR_INVALID_ARG;  // every rank must launch: an empty shard cannot take part
  if (grad_mode != DC_GRAD_NONE && grad_mode != DC_GRAD_SUM) return DC_ERR_INVALID_ARG;
  for (int k = 0; k < n_outs; ++k)
    if (!outs->ptr[k] || (reinterpret_cast<uintptr_t>(outs->ptr[k]) & 15) != 0 || !flags->ptr[k]) return DC_ERR_INVALID_ARG;
  if (!takes_tensor_core_kernel(*fk, *kernel, *sv, batch, grad_mode)) return DC_ERR_UNSUPPORTED;
  int dev = 0, num_sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return DC_ERR_NO_DEVICE;
  }
  const void* qd = device_alias(q);
  void* md = mirror ? device_alias(mirror) : nullptr;
  if (!qd || (mirror && !md)) return DC_ERR_INVALID_ARG;
  const int64_t rec = 1 + (grad_mode == DC_GRAD_SUM ? fk->dof : 0);
  float* mine = static_cast<float*>(outs->ptr[0]) + row_offset * rec;
  return tc_score_grad_sync(fk, kernel, sv, qd, batch, mine, rec, grad_mode == DC_GRAD_SUM ? mine + 1 : nullptr, rec, nullptr,
                            grad_mode, num_sms, (cudaStream_t)stream, outs, n_outs, md, flags, rank, epoch);
}

int dc_set_option(int32_t option, double value) { return tc_set_option(option, value); }
double dc_get_option(int32_t option) { return tc_get_option(option); }

}  // extern "C"
