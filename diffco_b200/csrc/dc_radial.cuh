// Radial kernels k(rho), rho = |x - s|^2, and the coefficient of (x - s) in dk/dx (diffco/kernel.py).
//
// eval() returns UNSCALED values; the constant factors are applied once per query after the sum over
// support vectors:   score = score_scale * sum w k      g_x = grad_scale * sum w coef (x - s).
//
//   RQKernel(gamma,p)      kernel.py:17-29   u = 1/(1+gamma/p rho)  k = u^p        coef = u^(p+1)   grad_scale = -2 gamma
//   Polyharmonic(k=1,eps)  kernel.py:59-79   k = sqrt(rho)          coef = 1/sqrt(rho) (0 at rho=0)  both scales 1/eps
//   Polyharmonic(k odd)    k = r^k           coef = k r^(k-2)        scales 1/eps
//   Polyharmonic(k even)   k = r^k log r (0 at 0)   coef = r^(k-2) (k log r + 1) (0 at 0)            scales 1/eps
//   MultiQuadratic(eps)    kernel.py:45-57   k = sqrt(rho/eps^2+1)  coef = 1/k     grad_scale = 1/eps^2
//   TemporalFKKernel       kernel.py:175-202 the last feature is time: rho_x over the first F-1 features, dt the last
//                          difference; ux = 1/(1+gx/px rho_x), ut = 1/(1+gt/pt dt^2), k = ux^px (ut^pt)^alpha,
//                          coef = ux k for the space features, coef_t = (gt/gx) alpha ut k for the time feature
//                          (radial_eval_temporal); grad_scale = -2 gx
//
// The r = 0 sub-gradient of Polyharmonic(1) is 0, matching torch.cdist's backward: rho is
// clamped to a tiny positive value before MUFU.RSQ, so coef stays finite and multiplies an exactly-zero
// difference vector.
#pragma once

#include "dc_common.cuh"

namespace dc {

enum RadialKind { KR_RQ2 = 0, KR_PH1 = 1, KR_MQ = 2, KR_GENERIC = 3 };

template <typename T>
struct RadialConsts {
  T c0;           // RQ: gamma/p ; MQ: 1/eps^2 ; PH: unused
  T score_scale;  // applied to sum w k
  T grad_scale;   // applied to sum w coef (x - s)
  int kind;       // dc_kernel_kind (generic path)
  int order;      // p or k       (generic path)
  // DC_K_RQ_TEMPORAL only
  T c0_t;         // gamma_t / p_t
  T alpha_pt;     // alpha * p_t: (ut^pt)^alpha = ut^(alpha pt)
  T coef_t_scale; // (gamma_t / gamma_x) * alpha
};

template <typename T>
__host__ inline bool make_radial_consts(const dc_kernel_desc& k, RadialConsts<T>* out) {
  out->kind = k.kind;
  out->order = k.order;
  out->c0_t = out->alpha_pt = out->coef_t_scale = (T)0;
  switch (k.kind) {
    case DC_K_RQ:
      if (k.order < 1) return false;
      out->c0 = (T)(k.param / k.order);
      out->score_scale = (T)1;
      out->grad_scale = (T)(-2.0 * k.param);
      return true;
    case DC_K_POLYHARMONIC:
      if (k.order < 1 || k.param == 0.0) return false;
      out->c0 = (T)0;
      out->score_scale = (T)(1.0 / k.param);
      out->grad_scale = (T)(1.0 / k.param);
      return true;
    case DC_K_MULTIQUADRIC:
      if (k.param == 0.0) return false;
      out->c0 = (T)(1.0 / (k.param * k.param));
      out->score_scale = (T)1;
      out->grad_scale = (T)(1.0 / (k.param * k.param));
      return true;
    case DC_K_RQ_TEMPORAL:
      if (k.order < 1 || k.order2 < 1 || k.param == 0.0) return false;
      out->c0 = (T)(k.param / k.order);
      out->score_scale = (T)1;
      out->grad_scale = (T)(-2.0 * k.param);
      out->c0_t = (T)(k.param2 / k.order2);
      out->alpha_pt = (T)(k.alpha * k.order2);
      out->coef_t_scale = (T)(k.param2 / k.param * k.alpha);
      return true;
    default:
      return false;
  }
}

inline int fast_radial_kind(const dc_kernel_desc& k) {
  if (k.kind == DC_K_RQ && k.order == 2) return KR_RQ2;
  if (k.kind == DC_K_POLYHARMONIC && k.order == 1) return KR_PH1;
  if (k.kind == DC_K_MULTIQUADRIC) return KR_MQ;
  return KR_GENERIC;
}

__device__ __forceinline__ float log_t(float x) { return logf(x); }
__device__ __forceinline__ double log_t(double x) { return log(x); }
__device__ __forceinline__ float pow_t(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double pow_t(double x, double y) { return pow(x, y); }

template <typename T>
__device__ __forceinline__ T ipow(T b, int e) {
  T r = (T)1;
  while (e > 0) {
    if (e & 1) r *= b;
    b *= b;
    e >>= 1;
  }
  return r;
}

template <int KIND, typename T>
__device__ __forceinline__ void radial_eval(const RadialConsts<T>& rc, T rho, T& k, T& coef) {
  if constexpr (KIND == KR_RQ2) {
    const T u = fast_rcp(fma(rho, rc.c0, (T)1));
    k = u * u;
    coef = k * u;
  } else if constexpr (KIND == KR_PH1) {
    const T ri = fast_rsqrt(max(rho, Tiny<T>::v));
    k = rho * ri;
    coef = ri;
  } else if constexpr (KIND == KR_MQ) {
    const T t = fma(rho, rc.c0, (T)1);
    const T ri = fast_rsqrt(t);
    k = t * ri;
    coef = ri;
  } else {
    if (rc.kind == DC_K_RQ) {
      const T u = fast_rcp(fma(rho, rc.c0, (T)1));
      k = ipow(u, rc.order);
      coef = k * u;
    } else if (rc.kind == DC_K_MULTIQUADRIC) {
      const T t = fma(rho, rc.c0, (T)1);
      const T ri = fast_rsqrt(t);
      k = t * ri;
      coef = ri;
    } else {  // polyharmonic, any order
      const T ri = fast_rsqrt(max(rho, Tiny<T>::v));
      const T r = rho * ri;
      const int n = rc.order;
      if (n & 1) {
        k = ipow(r, n);
        coef = (n == 1) ? ri : (T)n * ipow(r, n - 2);
      } else {
        const bool pos = rho > (T)0;
        const T lr = pos ? (T)0.5 * log_t(rho) : (T)0;
        const T rk2 = ipow(r, n - 2);
        k = pos ? rk2 * rho * lr : (T)0;
        coef = pos ? rk2 * ((T)n * lr + (T)1) : (T)0;
      }
    }
  }
}

// DC_K_RQ_TEMPORAL: rho_x = |dx|^2 over the space features, dt = the time difference.  coef multiplies the space
// differences and coef_t the time difference, both under the common grad_scale = -2 gamma_x.
template <typename T>
__device__ __forceinline__ void radial_eval_temporal(const RadialConsts<T>& rc, T rho_x, T dt, T& k, T& coef, T& coef_t) {
  const T ux = fast_rcp(fma(rho_x, rc.c0, (T)1));
  const T ut = fast_rcp(fma(dt * dt, rc.c0_t, (T)1));
  k = ipow(ux, rc.order) * pow_t(ut, rc.alpha_pt);
  coef = ux * k;
  coef_t = rc.coef_t_scale * ut * k;
}

// Packed form for the thread-per-query kernel: both halves of the pair are different QUERIES against the same support
// vector, so every step stays a 2-wide instruction (MUFU is scalar and runs on its own pipe).
template <int KIND>
__device__ __forceinline__ void radial_eval2(const RadialConsts<float>& rc, P2 rho, P2& k, P2& coef) {
  if constexpr (KIND == KR_RQ2) {
    const P2 t = pfma_bb(rho, rc.c0, 1.0f);
    const P2 u(fast_rcp(t.lo()), fast_rcp(t.hi()));
    k = pmul(u, u);
    coef = pmul(k, u);
  } else if constexpr (KIND == KR_PH1) {
    const P2 ri(fast_rsqrt(fmaxf(rho.lo(), Tiny<float>::v)), fast_rsqrt(fmaxf(rho.hi(), Tiny<float>::v)));
    k = pmul(rho, ri);
    coef = ri;
  } else {
    static_assert(KIND == KR_MQ, "thread-per-query kernel: RQ(p=2), Polyharmonic(1) and MultiQuadratic only");
    const P2 t = pfma_bb(rho, rc.c0, 1.0f);
    const P2 ri(fast_rsqrt(t.lo()), fast_rsqrt(t.hi()));
    k = pmul(t, ri);
    coef = ri;
  }
}

}  // namespace dc
