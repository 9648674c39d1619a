// Greedy kernel-perceptron training, the whole sequential loop on the device in ONE launch.
//
// Reference: DiffCo.train_perceptron, diffco/kernel_perceptrons.py:98-133 (single class) and the legacy
// MultiDiffCo.train_perceptron, diffco/deprecated/MultiDiffCo.py:50-83 (classes visited in order inside every
// outer iteration, one kernel matrix shared by all classes, `complete` flags never reset).  The reference runs
// ~10 small tensor ops per iteration from Python (~0.4 ms/iteration published); here one
// 1024-thread CTA keeps the loop on chip: block-wide first-index arg-min / arg-max, lazy kernel rows computed on
// demand (kernel_perceptrons.py:117-119) straight from the feature matrix, and the hypothesis axpy.
//
// Arithmetic is kept un-fused (explicit mul / add roundings, no FMA contraction) and in the dtype of the inputs
// so that, in float64, the sequence of selected indices reproduces the float64 reference (bit-exact support-vector
// selection is the parity gate, BASELINE.json north_star).  Tie-breaks return the FIRST extremal index like
// torch.min / torch.max on CPU.
#include "dc_common.cuh"
#include "dc_radial.cuh"

namespace dc {

constexpr int kTrainThreads = 1024;

template <typename T>
struct TrainArgs {
  RadialConsts<T> rc;
  const T* x;   // [N,F]
  const T* y;   // [N,C]
  T* gains;     // [N,C]
  T* hyp;       // [N,C]
  T* kmat;      // [N,N], or compact [cap,N]: only the rows the loop asks for (slot != nullptr)
  int* slot;    // compact storage: slot[i] = row of kmat holding K[i,:], -1 = not computed; nullptr = full N x N matrix
  long long cap;
  T* diag;      // [N]  (== diagonal of kmat; 0 marks "row not computed yet", kernel_perceptrons.py:117)
  long long* iters_out;  // [2]: last iteration index, number of kernel rows computed
  long long n;
  long long max_iteration;
  int n_feat;
  int n_class;
  int legacy_multi;
  T beta;
};

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T>
struct ValIdx {
  T v;
  long long i;
};

// first-index arg-min (SIGN=+1) / arg-max (SIGN=-1) over the block; every thread returns the result.
template <typename T, int SIGN>
__device__ ValIdx<T> block_argext(T v, long long i, ValIdx<T>* smem) {
  auto better = [](T av, long long ai, T bv, long long bi) {
    // is (a) strictly preferable to (b)?  NaNs never win, matching "first extremal index" on well-formed data.
    if (SIGN > 0) return (av < bv) || (av == bv && ai < bi);
    return (av > bv) || (av == bv && ai < bi);
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T ov = __shfl_xor_sync(0xffffffffu, v, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (better(ov, oi, v, i)) {
      v = ov;
      i = oi;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = ValIdx<T>{v, i};
  __syncthreads();
  if (warp == 0) {
    ValIdx<T> r = smem[lane];  // kTrainThreads / 32 == 32 warps
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, r.v, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, r.i, o);
      if (better(ov, oi, r.v, r.i)) {
        r.v = ov;
        r.i = oi;
      }
    }
    if (lane == 0) smem[32] = r;
  }
  __syncthreads();
  const ValIdx<T> out = smem[32];
  __syncthreads();
  return out;
}

__device__ long long block_sum_ll(long long v, long long* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    long long r = smem[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) smem[32] = r;
  }
  __syncthreads();
  const long long out = smem[32];
  __syncthreads();
  return out;
}

template <typename T>
__global__ void __launch_bounds__(kTrainThreads, 1) perceptron_train_kernel(const TrainArgs<T> a) {
  __shared__ ValIdx<T> s_vi[33];
  __shared__ long long s_ll[33];
  __shared__ T s_xi[DC_MAX_FEATURES];
  __shared__ int s_complete[DC_MAX_CLASSES];
  __shared__ long long s_used;  // compact storage: rows handed out so far

  const int tid = threadIdx.x;
  const long long N = a.n;
  const int C = a.n_class, F = a.n_feat;
  const T inf = (T)INFINITY;
  if (tid < DC_MAX_CLASSES) s_complete[tid] = 0;
  if (a.slot != nullptr) {
    // rows pre-loaded by the caller (jump start) occupy the slots 0 .. max(slot)
    long long mx = -1;
    for (long long j = tid; j < N; j += kTrainThreads) mx = max(mx, (long long)a.slot[j]);
    const ValIdx<T> top = block_argext<T, -1>((T)mx, 0, s_vi);
    if (tid == 0) s_used = (long long)top.v + 1;
  }
  __syncthreads();

  long long rows_computed = 0;
  long long it = 0;
  bool stop = false;
  for (it = 0; it < a.max_iteration && !stop; ++it) {
    bool any_progress = false;
    for (int c = 0; c < C; ++c) {
      // ---- margin = y * h ; (min_margin, min_i) = min(margin)  [first index] ---------------------------
      T bv = inf;
      long long bi = N;
      for (long long j = tid; j < N; j += kTrainThreads) {
        const T m = mul_rn(a.y[j * C + c], a.hyp[j * C + c]);
        if (m < bv) {  // strided ascending j: strict '<' keeps the first index per thread
          bv = m;
          bi = j;
        }
      }
      const ValIdx<T> mn = block_argext<T, +1>(bv, bi, s_vi);
      const long long i = mn.i;

      // ---- lazy kernel row: K[i] = k(X_t[i], X_t); K[:, i] = K[i] ----------------------------------------
      const bool need_row = (a.diag[i] == (T)0);
      __syncthreads();  // every thread has sampled the flag before the row's owner thread overwrites it
      if (need_row && a.slot != nullptr) {
        if (s_used >= a.cap) {  // out of row storage: report and stop (the caller retries with a larger capacity)
          if (tid == 0) {
            a.iters_out[0] = it;
            a.iters_out[1] = -1;
          }
          return;
        }
        if (tid == 0) a.slot[i] = (int)s_used;
        __syncthreads();
        if (tid == 0) s_used += 1;
      }
      const long long ri = (a.slot != nullptr) ? (long long)a.slot[i] : i;  // row of kmat that holds K[i,:]
      if (need_row) {
        for (int f = tid; f < F; f += kTrainThreads) s_xi[f] = a.x[i * F + f];
        __syncthreads();
        for (long long j = tid; j < N; j += kTrainThreads) {
          const T* xj = a.x + j * F;
          const bool temporal = a.rc.kind == DC_K_RQ_TEMPORAL;
          const int Fx = temporal ? F - 1 : F;
          T rho = (T)0;
          for (int f = 0; f < Fx; ++f) {
            const T d = s_xi[f] - xj[f];
            rho = fma(d, d, rho);
          }
          T k, coef, coef_t;
          if (temporal)
            radial_eval_temporal<T>(a.rc, rho, s_xi[F - 1] - xj[F - 1], k, coef, coef_t);
          else
            radial_eval<KR_GENERIC, T>(a.rc, rho, k, coef);
          k = k * a.rc.score_scale;
          a.kmat[ri * N + j] = k;
          if (a.slot == nullptr) a.kmat[j * N + i] = k;  // K[:, i] = K[i] (kernel_perceptrons.py:119); compact: rows only
          if (j == i) a.diag[i] = k;
        }
        ++rows_computed;
        __syncthreads();
      }
      const T* krow = a.kmat + ri * N;

      if (mn.v <= (T)0) {
        // delta = (beta^((1+y_i)/2) * y_i - h_i) / K_ii ; gains_i += delta ; h += delta * K[i]
        const T yi = a.y[i * C + c];
        const T hi = a.hyp[i * C + c];
        const T kii = krow[i];
        const T target = mul_rn(pow_t(a.beta, (T)0.5 * ((T)1 + yi)), yi);
        const T delta = add_rn(target, -hi) / kii;
        __syncthreads();  // everyone has read h_i before it is updated
        if (tid == 0) a.gains[i * C + c] = add_rn(a.gains[i * C + c], delta);
        for (long long j = tid; j < N; j += kTrainThreads) a.hyp[j * C + c] = add_rn(a.hyp[j * C + c], mul_rn(delta, krow[j]));
        __syncthreads();
        any_progress = true;
        continue;
      }

      // ---- modified margin: y * (h - gains * diag K) * (gains != 0) ; drop the largest if positive ----------
      T xv = -inf;
      long long xi = N;
      long long nz = 0;
      for (long long j = tid; j < N; j += kTrainThreads) {
        const T gj = a.gains[j * C + c];
        const T ind = (gj != (T)0) ? (T)1 : (T)0;
        nz += (gj != (T)0);
        const T m = mul_rn(mul_rn(a.y[j * C + c], add_rn(a.hyp[j * C + c], -mul_rn(gj, a.diag[j]))), ind);
        if (m > xv) {
          xv = m;
          xi = j;
        }
      }
      const ValIdx<T> mx = block_argext<T, -1>(xv, xi, s_vi);
      const long long count = block_sum_ll(nz, s_ll);
      if (mx.v > (T)0 && count > 1) {
        const long long r = mx.i;
        const T gr = a.gains[r * C + c];
        const T* rrow = a.kmat + ((a.slot != nullptr) ? (long long)a.slot[r] : r) * N;  // nonzero gain => its row exists
        __syncthreads();
        for (long long j = tid; j < N; j += kTrainThreads) a.hyp[j * C + c] = add_rn(a.hyp[j * C + c], -mul_rn(gr, rrow[j]));
        if (tid == 0) a.gains[r * C + c] = (T)0;
        __syncthreads();
        any_progress = true;
        continue;
      }
      if (a.legacy_multi) {
        if (tid == 0) s_complete[c] = 1;
        __syncthreads();
      }
    }
    if (a.legacy_multi) {
      // deprecated/MultiDiffCo.py:78-79: stop once every class has been marked complete (flags are sticky)
      bool all = true;
      for (int c = 0; c < C; ++c) all = all && (s_complete[c] != 0);
      if (all) stop = true;
    } else {
      // kernel_perceptrons.py:133: break when neither update applied
      if (!any_progress) stop = true;
    }
    if (stop) break;
  }
  if (tid == 0) {
    // `it` of the reference after the loop: index of the last iteration executed
    a.iters_out[0] = (it >= a.max_iteration) ? a.max_iteration - 1 : it;
    a.iters_out[1] = rows_computed;
  }
}

template <typename T>
static int train_typed(const dc_kernel_desc* kernel, const void* x, const void* y, int64_t n, int32_t F, int32_t C,
                       double beta, int64_t max_iteration, void* gains, void* hyp, void* kmat, int32_t* slot, int64_t cap,
                       void* diag, int32_t legacy_multi, int64_t* iters_out, cudaStream_t stream) {
  TrainArgs<T> a;
  if (!make_radial_consts<T>(*kernel, &a.rc)) return DC_ERR_INVALID_ARG;
  a.x = (const T*)x;
  a.y = (const T*)y;
  a.gains = (T*)gains;
  a.hyp = (T*)hyp;
  a.kmat = (T*)kmat;
  a.slot = slot;
  a.cap = cap;
  a.diag = (T*)diag;
  a.iters_out = (long long*)iters_out;
  a.n = n;
  a.max_iteration = max_iteration;
  a.n_feat = F;
  a.n_class = C;
  a.legacy_multi = legacy_multi;
  a.beta = (T)beta;
  perceptron_train_kernel<T><<<1, kTrainThreads, 0, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dc

extern "C" int dc_perceptron_train(const dc_kernel_desc* kernel, const void* x_feat, const void* y, int64_t n,
                                   int32_t n_features, int32_t n_class, int32_t dtype, double beta, int64_t max_iteration,
                                   void* gains, void* hypothesis, void* kernel_matrix, void* diag, int32_t legacy_multi,
                                   int64_t* iterations_out, dc_stream_t stream) {
  return dc_perceptron_train_rows(kernel, x_feat, y, n, n_features, n_class, dtype, beta, max_iteration, gains, hypothesis,
                                  kernel_matrix, nullptr, n, diag, legacy_multi, iterations_out, stream);
}

extern "C" int dc_perceptron_train_rows(const dc_kernel_desc* kernel, const void* x_feat, const void* y, int64_t n,
                                        int32_t n_features, int32_t n_class, int32_t dtype, double beta, int64_t max_iteration,
                                        void* gains, void* hypothesis, void* kernel_rows, int32_t* row_slot, int64_t row_capacity,
                                        void* diag, int32_t legacy_multi, int64_t* iterations_out, dc_stream_t stream) {
  void* kernel_matrix = kernel_rows;
  if (!kernel || !x_feat || !y || !gains || !hypothesis || !kernel_matrix || !diag || !iterations_out)
    return DC_ERR_INVALID_ARG;
  if (row_slot != nullptr && (row_capacity < 1 || row_capacity > n)) return DC_ERR_INVALID_ARG;
  if (n < 1 || n_features < 1 || n_features > DC_MAX_FEATURES || n_class < 1 || n_class > DC_MAX_CLASSES || max_iteration < 0)
    return DC_ERR_INVALID_ARG;
  if (!legacy_multi && n_class != 1) return DC_ERR_INVALID_ARG;
  if (dtype == DC_F32)
    return dc::train_typed<float>(kernel, x_feat, y, n, n_features, n_class, beta, max_iteration, gains, hypothesis,
                                  kernel_matrix, row_slot, row_capacity, diag, legacy_multi, iterations_out, (cudaStream_t)stream);
  if (dtype == DC_F64)
    return dc::train_typed<double>(kernel, x_feat, y, n, n_features, n_class, beta, max_iteration, gains, hypothesis,
                                   kernel_matrix, row_slot, row_capacity, diag, legacy_multi, iterations_out, (cudaStream_t)stream);
  return DC_ERR_INVALID_ARG;
}
