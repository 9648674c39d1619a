// Fused collision-score kernel, lane-split form: any feature count F <= 64, any class count C <= 8, any
// radial kernel order, fp32 or fp64.  This is the path for
//   * the small batches the reference optimisers actually issue (B ~ 18..256 float64 waypoints per call,
//     diffco/optim.py:86-103,190-218),
//   * float64 callers in general (the reference scripts run in double, scripts/speed_compare.py:209-211),
//   * shapes the thread-per-query kernel is not instantiated for (e.g. DualPandaFK, F = 42).
//
// One query is owned by WPQ warps of an 8-warp CTA (WPQ in {1,2,4,8}, chosen on the host so that small
// batches still spread over the SMs).  The 32*WPQ lanes stride over the rows of the packed support table,
// read straight from L2/HBM with 16-byte loads (each lane streams whole rows, a warp covers a contiguous
// 32-row block), keep per-lane partial sums of the C scores and the F feature-gradients in registers, and
// combine them with warp-shuffle reductions followed by a fixed-order pass over the WPQ warp leaders.
// FK and its J^T product run on the leader lane before / after the loop.
#pragma once

#include "dc_common.cuh"
#include "dc_fk.cuh"
#include "dc_radial.cuh"

namespace dc {

template <typename T>
struct LsArgs {
  dc_fk_desc fk;
  RadialConsts<T> rc;
  const T* table;
  const T* q;
  T* score;
  T* grad;
  const T* grad_out;
  long long batch;
  long long score_ld;
  long long grad_ld;
  int n_sv;
  int n_feat;
  int n_class;
  int n_in;
  int f_pad;
  int row_stride;
  int wpq;        // warps per query
  int grad_mode;  // dc_grad_mode
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, T* out);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* out) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  out[0] = v.x;
  out[1] = v.y;
  out[2] = v.z;
  out[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<double>(const double* p, double* out) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(p));
  const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
  out[0] = a.x;
  out[1] = a.y;
  out[2] = b.x;
  out[3] = b.y;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLsWarps = 8;

// RCH = row_stride / 4 (number of 4-element chunks per table row), compile time so that the row and the
// gradient accumulators live in registers.
template <typename T, int RCH, bool GRAD>
__global__ void __launch_bounds__(kLsWarps * 32) score_ls_kernel(const __grid_constant__ LsArgs<T> a) {
  constexpr int RW = 4 * RCH;
  __shared__ T xs[kLsWarps][DC_MAX_FEATURES + 4];
  __shared__ T part[kLsWarps][DC_MAX_FEATURES + DC_MAX_CLASSES];
  __shared__ T gos[kLsWarps][DC_MAX_CLASSES];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wpq = a.wpq;
  const int qpc = kLsWarps / wpq;  // queries per CTA
  const int grp = warp / wpq;      // query slot inside the CTA
  const int sub = warp - grp * wpq;
  const int lead = grp * wpq;      // leader warp of the group
  const long long b = (long long)blockIdx.x * qpc + grp;
  const bool active = b < a.batch;
  const int F = a.n_feat, C = a.n_class;
  T qv[DC_MAX_DOF];

  // ---- FK on the leader lane -------------------------------------------------------------------------
  if (sub == 0 && lane == 0) {
    if (active) {
      const T* qp = a.q + (size_t)b * a.n_in;
      if (a.fk.type == DC_FK_NONE) {
        for (int f = 0; f < F; ++f) xs[lead][f] = qp[f];
      } else {
#pragma unroll
        for (int i = 0; i < DC_MAX_DOF; ++i) qv[i] = (i < a.n_in) ? qp[i] : (T)0;
        fk_features(a.fk, qv, &xs[lead][0], 1);
      }
    } else {
      for (int f = 0; f < F; ++f) xs[lead][f] = (T)0;
    }
  }
  __syncthreads();

  const int n_pass = (GRAD && a.grad_mode == DC_GRAD_JAC) ? C : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    if (sub == 0 && lane < DC_MAX_CLASSES) {
      T v = (T)0;
      if (lane < C) {
        if (a.grad_mode == DC_GRAD_JAC)
          v = (lane == pass) ? (T)1 : (T)0;
        else
          v = (a.grad_out != nullptr && active) ? a.grad_out[(size_t)b * C + lane] : (T)1;
      }
      gos[lead][lane] = v;
    }
    __syncthreads();

    T sc[DC_MAX_CLASSES];
    T g[GRAD ? RW : 1];
#pragma unroll
    for (int c = 0; c < DC_MAX_CLASSES; ++c) sc[c] = (T)0;
#pragma unroll
    for (int e = 0; e < (GRAD ? RW : 1); ++e) g[e] = (T)0;
    T gt = (T)0;
    const bool temporal = a.rc.kind == DC_K_RQ_TEMPORAL;
    const int Fx = temporal ? F - 1 : F;  // features that enter rho

    // Small rows: keep the query features and the difference vector in registers.  Large rows (F > 32 fp32 /
    // F > 16 fp64): re-read x from shared memory and the row from L1 for the gradient update instead.
    constexpr bool KEEP = (RW * sizeof(T) <= 128);
    T xr[KEEP ? RW : 1];
    if constexpr (KEEP) {
#pragma unroll
      for (int e = 0; e < RW; ++e) xr[e] = (e < F) ? xs[lead][e] : (T)0;
    }
    const T* xq = &xs[lead][0];

    for (int n = sub * 32 + lane; n < a.n_sv; n += 32 * wpq) {
      const T* rp = a.table + (size_t)n * a.row_stride;
      T d[KEEP ? RW : 4];
      T rho = (T)0;
      if constexpr (KEEP) {
        T row[RW];
#pragma unroll
        for (int j = 0; j < RCH; ++j) load4<T>(rp + 4 * j, row + 4 * j);
#pragma unroll
        for (int e = 0; e < RW; ++e) {
          d[e] = (e < F) ? xr[e] + row[e] : (T)0;  // table holds -s; entries past F are weights / padding
          rho = fma(d[e], (e < Fx) ? d[e] : (T)0, rho);
        }
      } else {
#pragma unroll
        for (int j = 0; j < RCH; ++j) {
          load4<T>(rp + 4 * j, d);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const T dv = (4 * j + e < Fx) ? xq[4 * j + e] + d[e] : (T)0;
            rho = fma(dv, dv, rho);
          }
        }
      }
      T k, coef, coef_t = (T)0, dt = (T)0;
      if (temporal) {  // the last feature is time (kernel.py:175-202): it leaves rho and gets its own coefficient
        dt = xq[F - 1] + __ldg(rp + F - 1);
        radial_eval_temporal<T>(a.rc, rho, dt, k, coef, coef_t);
      } else {
        radial_eval<KR_GENERIC, T>(a.rc, rho, k, coef);
      }
      T om = (T)0;
#pragma unroll
      for (int c = 0; c < DC_MAX_CLASSES; ++c) {
        if (c < C) {
          const T w = __ldg(rp + a.f_pad + c);
          sc[c] = fma(w, k, sc[c]);
          if (GRAD) om = fma(gos[lead][c], w, om);
        }
      }
      if (GRAD) {
        const T cv = om * coef;
        gt = fma(om * (coef_t - coef), dt, gt);  // corrects the last feature's term below; 0 unless temporal
        if constexpr (KEEP) {
#pragma unroll
          for (int e = 0; e < RW; ++e) g[e] = fma(cv, d[e], g[e]);
        } else {
#pragma unroll
          for (int j = 0; j < RCH; ++j) {
            load4<T>(rp + 4 * j, d);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const T dv = (4 * j + e < F) ? xq[4 * j + e] + d[e] : (T)0;
              g[4 * j + e] = fma(cv, dv, g[4 * j + e]);
            }
          }
        }
      }
    }

    // ---- reductions: lanes (shuffles), then the WPQ warps of the group (fixed order) -----------------
#pragma unroll
    for (int c = 0; c < DC_MAX_CLASSES; ++c) {
      if (c < C) {
        const T v = warp_sum(sc[c]);
        if (lane == 0) part[warp][c] = v;
      }
    }
    if (GRAD) {
#pragma unroll
      for (int e = 0; e < RW; ++e) {
        if (e < F) {
          const T v = warp_sum(g[e]);
          if (lane == 0) part[warp][DC_MAX_CLASSES + e] = v;
        }
      }
      if (temporal) {
        const T v = warp_sum(gt);
        if (lane == 0) part[warp][DC_MAX_CLASSES + F - 1] += v;
      }
    }
    __syncthreads();
    if (sub == 0 && active) {
      // scores (first pass only), one class per lane
      if (pass == 0 && lane < C) {
        T s = (T)0;
        for (int w = 0; w < wpq; ++w) s += part[lead + w][lane];
        a.score[(size_t)b * a.score_ld + lane] = s * a.rc.score_scale;
      }
      if (GRAD) {
        // reduced feature gradient back into part[lead] (lanes stride over features)
        for (int e = lane; e < F; e += 32) {
          T s = (T)0;
          for (int w = 0; w < wpq; ++w) s += part[lead + w][DC_MAX_CLASSES + e];
          part[lead][DC_MAX_CLASSES + e] = s * a.rc.grad_scale;
        }
        __syncwarp();
        if (lane == 0) {
          T* out = a.grad + (size_t)b * a.grad_ld + (a.grad_mode == DC_GRAD_JAC ? (size_t)pass * a.n_in : 0);
          if (a.fk.type == DC_FK_NONE) {
            for (int f = 0; f < F; ++f) out[f] = part[lead][DC_MAX_CLASSES + f];
          } else {
            T gq[DC_MAX_DOF];
#pragma unroll
            for (int i = 0; i < DC_MAX_DOF; ++i) gq[i] = (T)0;
            fk_vjp<T>(a.fk, qv, &xs[lead][0], 1, &part[lead][DC_MAX_CLASSES], 1, gq);
            for (int i = 0; i < a.n_in; ++i) out[i] = gq[i];
          }
        }
      }
    }
    __syncthreads();
  }
}

#ifdef DC_LS_INSTANTIATE
template <typename T, int RCH, bool GRAD>
int launch_score_ls_inst(const LsArgs<T>& a, cudaStream_t stream) {
  const int qpc = kLsWarps / a.wpq;
  const long long grid = ceil_div64(a.batch, qpc);
  if (grid > 0x7fffffffLL) return DC_ERR_UNSUPPORTED;
  score_ls_kernel<T, RCH, GRAD><<<(int)grid, kLsWarps * 32, 0, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

template <typename T, bool GRAD>
int launch_score_ls_rch(const LsArgs<T>& a, cudaStream_t stream) {
  switch (a.row_stride / 4) {
#define DC_LS_CASE(R) \
  case R:             \
    return launch_score_ls_inst<T, R, GRAD>(a, stream);
    DC_LS_CASE(1)
    DC_LS_CASE(2)
    DC_LS_CASE(3)
    DC_LS_CASE(4)
    DC_LS_CASE(5)
    DC_LS_CASE(6)
    DC_LS_CASE(7)
    DC_LS_CASE(8)
    DC_LS_CASE(9)
    DC_LS_CASE(10)
    DC_LS_CASE(11)
    DC_LS_CASE(12)
    DC_LS_CASE(13)
    DC_LS_CASE(14)
    DC_LS_CASE(15)
    DC_LS_CASE(16)
    DC_LS_CASE(17)
    DC_LS_CASE(18)
#undef DC_LS_CASE
    default:
      return DC_ERR_UNSUPPORTED;
  }
}

template <typename T>
int launch_score_ls(LsArgs<T>& a, int num_sms, cudaStream_t stream) {
  // spread small batches: use more warps per query until the grid covers the SMs about twice
  int wpq = 1;
  while (wpq < kLsWarps && ceil_div64(a.batch * wpq, kLsWarps) < 2LL * num_sms) wpq <<= 1;
  a.wpq = wpq;
  return (a.grad_mode == DC_GRAD_NONE) ? launch_score_ls_rch<T, false>(a, stream) : launch_score_ls_rch<T, true>(a, stream);
}

#endif  // DC_LS_INSTANTIATE

}  // namespace dc
