// dc_traj_step — one iteration of the reference's penalty trajectory optimiser (Weighted.step, diffco/optim.py:706-752;
// same terms as adam_traj_optimize, optim.py:86-127) for W waypoints in ONE launch of ONE CTA:
//
//   control points cp = FK(p)                                   (robot.fkine, model.py)
//   path length  sum |cp[i+1]-cp[i]|^2                          optim.py:715-717
//   max move     sum clamp(|cp[i+1,m]-cp[i,m]|^2 - v^2, 0)      optim.py:718-719
//   joint limit  sum clamp(lo-p,0) + clamp(p-hi,0)              optim.py:720-722
//   collision    sum clamp(score + bias, 0)                     optim.py:711-713 (score and d score/dp come from dc_score_grad)
//   gradient of  dif*length + w_mm*max_move + w_jl*joint_limit + w_col*collision  w.r.t. p, analytically (J_FK^T products),
//   masked, then torch.optim.Adam's update and robot.wrap.
//
// The reference builds this from ~40 tensor ops + autograd per step; the loop is launch-latency bound there
// (SURVEY.md §3.4).  Thread w owns waypoint w; control points are exchanged through shared memory; the four sums are
// reduced in a fixed order.
#include "dc_common.cuh"
#include "dc_fk.cuh"

namespace dc {

template <typename T>
struct TrajArgs {
  dc_fk_desc fk;
  dc_traj_params prm;
  T* p;
  const T* score;
  const T* score_grad;
  const T* mask;
  T* exp_avg;
  T* exp_avg_sq;
  double* step;
  T* terms;
  int n_wp;
  int n_feat;
};

template <typename T>
__device__ __forceinline__ T wrap2pi_t(T th) {  // (pi + th) % (2 pi) - pi with Python's sign convention (utils.py:51-52)
  const T pi = (T)3.14159265358979323846, two_pi = (T)(2.0 * 3.14159265358979323846);
  const T a = pi + th;
  T m = fmod(a, two_pi);
  if (m != (T)0 && m < (T)0) m += two_pi;
  return m - pi;
}

template <typename T>
__global__ void __launch_bounds__(256, 1) traj_step_kernel(const __grid_constant__ TrajArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cps = reinterpret_cast<T*>(smem_raw);                  // [W][F]
  T* red = cps + (size_t)a.n_wp * a.n_feat;                 // [5][8] per-warp partial sums
  const int tid = threadIdx.x, W = a.n_wp, F = a.n_feat, D = a.fk.dof;
  const int M = a.fk.type == DC_FK_NONE ? F : a.fk.n_points, dim = a.fk.type == DC_FK_NONE ? 1 : a.fk.point_dim;
  const dc_traj_params& P = a.prm;
  const double step_d = *a.step + 1.0;

  for (int w = tid; w < W; w += blockDim.x) {
    T q[DC_MAX_DOF];
#pragma unroll
    for (int i = 0; i < DC_MAX_DOF; ++i) q[i] = (i < D) ? a.p[(size_t)w * D + i] : (T)0;
    fk_forward<T>(a.fk, q, cps + (size_t)w * F, 1);
  }
  __syncthreads();

  const T v2 = (T)(P.max_speed * P.max_speed);
  const T bc1 = (T)(1.0 - pow(P.beta1, step_d)), bc2_sqrt = (T)sqrt(1.0 - pow(P.beta2, step_d));
  const T step_size = (T)P.lr / bc1;
  T acc_len = 0, acc_col = 0, acc_jl = 0, acc_mm = 0, acc_g2 = 0;
  for (int w = tid; w < W; w += blockDim.x) {
    T q[DC_MAX_DOF], gq[DC_MAX_DOF], gcp[DC_MAX_FEATURES];
#pragma unroll
    for (int i = 0; i < DC_MAX_DOF; ++i) {
      q[i] = (i < D) ? a.p[(size_t)w * D + i] : (T)0;
      gq[i] = (T)0;
    }
    for (int f = 0; f < F; ++f) gcp[f] = (T)0;
    const T* c0 = cps + (size_t)w * F;
    for (int m = 0; m < M; ++m) {
      if (w > 0) {
        const T* cm = c0 - F;
        T ss = 0;
        for (int k = 0; k < dim; ++k) {
          const T d = c0[m * dim + k] - cm[m * dim + k];
          ss = fma(d, d, ss);
        }
        const T coef = (T)2 * ((T)P.dif_weight + ((P.max_move_weight != 0 && ss - v2 > (T)0) ? (T)P.max_move_weight : (T)0));
        for (int k = 0; k < dim; ++k) gcp[m * dim + k] += coef * (c0[m * dim + k] - cm[m * dim + k]);
      }
      if (w < W - 1) {
        const T* cn = c0 + F;
        T ss = 0;
        for (int k = 0; k < dim; ++k) {
          const T d = cn[m * dim + k] - c0[m * dim + k];
          ss = fma(d, d, ss);
        }
        acc_len += ss;
        if (P.max_move_weight != 0) acc_mm += max(ss - v2, (T)0);
        const T coef = (T)2 * ((T)P.dif_weight + ((P.max_move_weight != 0 && ss - v2 > (T)0) ? (T)P.max_move_weight : (T)0));
        for (int k = 0; k < dim; ++k) gcp[m * dim + k] -= coef * (cn[m * dim + k] - c0[m * dim + k]);
      }
    }
    fk_vjp<T>(a.fk, q, const_cast<T*>(c0), 1, gcp, 1, gq);
    T hinge_on = (T)0;
    if (a.score != nullptr && P.collision_weight != 0) {
      const T hinge = a.score[w] + (T)P.safety_bias;
      acc_col += max(hinge, (T)0);
      hinge_on = hinge > (T)0 ? (T)P.collision_weight : (T)0;
    }
    for (int i = 0; i < D; ++i) {
      const size_t idx = (size_t)w * D + i;
      T g = gq[i];
      if (hinge_on != (T)0) g += hinge_on * a.score_grad[idx];
      if (P.joint_limit_weight != 0) {
        const T lo = (T)P.limits[i][0], hi = (T)P.limits[i][1];
        acc_jl += max(lo - q[i], (T)0) + max(q[i] - hi, (T)0);
        g += (T)P.joint_limit_weight * ((q[i] > hi ? (T)1 : (T)0) - (q[i] < lo ? (T)1 : (T)0));
      }
      if (a.mask != nullptr) g *= a.mask[idx];
      acc_g2 = fma(g, g, acc_g2);
      // torch.optim.Adam (no weight decay / amsgrad): lerp, addcmul, addcdiv
      T m1 = a.exp_avg[idx], m2 = a.exp_avg_sq[idx];
      m1 = m1 + (g - m1) * (T)(1.0 - P.beta1);
      m2 = m2 * (T)P.beta2 + (T)(1.0 - P.beta2) * g * g;
      a.exp_avg[idx] = m1;
      a.exp_avg_sq[idx] = m2;
      const T denom = sqrt(m2) / bc2_sqrt + (T)P.eps;
      T pn = q[i] - step_size * (m1 / denom);
      if (P.wrap[i]) pn = wrap2pi_t(pn);
      a.p[idx] = pn;
    }
  }
  // fixed-order block reduction of the four sums
  T vals[5] = {acc_len, acc_col, acc_jl, acc_mm, acc_g2};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    T v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[k * 8 + (tid >> 5)] = v;
  }
  __syncthreads();
  if (tid == 0) {
    T s[5];
    for (int k = 0; k < 5; ++k) {
      s[k] = 0;
      for (int wp = 0; wp < (int)(blockDim.x >> 5); ++wp) s[k] += red[k * 8 + wp];
    }
    a.terms[0] = s[0];
    a.terms[1] = s[1];
    a.terms[2] = s[2];
    a.terms[3] = s[3];
    a.terms[4] = (T)P.collision_weight * s[1] + (T)P.max_move_weight * s[3] + (T)P.joint_limit_weight * s[2];
    a.terms[5] = s[4];  // |masked gradient|^2 (adam_traj_optimize's stopping test, optim.py:121-122)
    *a.step = step_d;
  }
}

template <typename T>
static int launch_traj_step(const dc_fk_desc* fk, const dc_traj_params* prm, int64_t n_wp, void* p, const void* score,
                            const void* score_grad, const void* mask, void* exp_avg, void* exp_avg_sq, double* step, void* terms,
                            cudaStream_t stream) {
  TrajArgs<T> a;
  a.fk = *fk;
  a.prm = *prm;
  a.p = static_cast<T*>(p);
  a.score = static_cast<const T*>(score);
  a.score_grad = static_cast<const T*>(score_grad);
  a.mask = static_cast<const T*>(mask);
  a.exp_avg = static_cast<T*>(exp_avg);
  a.exp_avg_sq = static_cast<T*>(exp_avg_sq);
  a.step = step;
  a.terms = static_cast<T*>(terms);
  a.n_wp = (int)n_wp;
  a.n_feat = fk->type == DC_FK_NONE ? fk->dof : fk->n_points * fk->point_dim;
  const size_t smem = sizeof(T) * ((size_t)n_wp * a.n_feat + 40);
  if (smem > 200 * 1024) return DC_ERR_UNSUPPORTED;
  auto kern = traj_step_kernel<T>;
  DC_SET_FUNC_ATTR_ONCE(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  kern<<<1, 256, smem, stream>>>(a);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // namespace dc

extern "C" int dc_traj_step(const dc_fk_desc* fk, const dc_traj_params* prm, int64_t n_waypoints, int32_t dtype, void* p,
                            const void* score, const void* score_grad, const void* mask, void* exp_avg, void* exp_avg_sq,
                            double* step, void* terms, dc_stream_t stream) {
  if (!fk || !prm || !p || !exp_avg || !exp_avg_sq || !step || !terms) return DC_ERR_INVALID_ARG;
  if (n_waypoints < 2 || n_waypoints > (1 << 20) || fk->dof < 1 || fk->dof > DC_MAX_DOF) return DC_ERR_INVALID_ARG;
  if ((score == nullptr) != (score_grad == nullptr)) return DC_ERR_INVALID_ARG;
  const int F = fk->type == DC_FK_NONE ? fk->dof : fk->n_points * fk->point_dim;
  if (F < 1 || F > DC_MAX_FEATURES) return DC_ERR_INVALID_ARG;
  if (!(prm->lr > 0) || !(prm->beta1 >= 0 && prm->beta1 < 1) || !(prm->beta2 >= 0 && prm->beta2 < 1)) return DC_ERR_INVALID_ARG;
  if (dtype == DC_F32)
    return dc::launch_traj_step<float>(fk, prm, n_waypoints, p, score, score_grad, mask, exp_avg, exp_avg_sq, step, terms,
                                       (cudaStream_t)stream);
  if (dtype == DC_F64)
    return dc::launch_traj_step<double>(fk, prm, n_waypoints, p, score, score_grad, mask, exp_avg, exp_avg_sq, step, terms,
                                        (cudaStream_t)stream);
  return DC_ERR_INVALID_ARG;
}
