// One translation unit per (radial kind, class width, mode): compiled several times by diffco_b200/build.py with
//   -DDC_TQ_KIND=<KR_*> -DDC_TQ_CW=<1|4> -DDC_TQ_MODE=<M_*> -DDC_TQ_NAME=<symbol>
// so the thread-per-query instantiations build in parallel.
#include "dc_score_tq.cuh"

#ifndef DC_TQ_KIND
#error "compile with -DDC_TQ_KIND -DDC_TQ_CW -DDC_TQ_MODE -DDC_TQ_NAME"
#endif

namespace dc {

// Feature counts of the reference's feature maps (diffco/model.py): planar chains 2..8 links, SE(2)/SE(3) bodies,
// Baxter (12 / 24), Panda (21; 15 in robot_fkine.py; 27 through its URDF), raw configurations.  The per-lane state (x, d,
// g: 3F packed registers) must fit the 128 registers a 512-thread CTA leaves per thread — F <= 16 — or the 255 of a
// 256-thread CTA (F = 21, 24, 27); other widths take the lane-split kernel.
template <int F>
static int launch_f(ScoreArgs<float>& a, int num_sms, cudaStream_t stream) {
  // four concurrent tiles per SM when there are enough tiles to go around, one 16-warp tile otherwise
  const long long tiles = (a.batch + 63) / 64;
  constexpr int NW = F > 16 ? 8 : 16;  // warps per CTA (TqCfg::NW)
  if (tiles >= 3LL * num_sms) return launch_score_tq<F, DC_TQ_KIND, DC_TQ_CW, DC_TQ_MODE, 4, 3>(a, num_sms, stream);
  return launch_score_tq<F, DC_TQ_KIND, DC_TQ_CW, DC_TQ_MODE, NW, 3>(a, num_sms, stream);
}

int DC_TQ_NAME(int n_feat, ScoreArgs<float>& a, int num_sms, cudaStream_t stream) {
  switch (n_feat) {
#define DC_TQ_CASE(F) \
  case F:             \
    return launch_f<F>(a, num_sms, stream);
    DC_TQ_CASE(2)
    DC_TQ_CASE(3)
    DC_TQ_CASE(4)
    DC_TQ_CASE(6)
    DC_TQ_CASE(7)
    DC_TQ_CASE(8)
    DC_TQ_CASE(10)
    DC_TQ_CASE(12)
    DC_TQ_CASE(14)
    DC_TQ_CASE(15)
    DC_TQ_CASE(16)
    DC_TQ_CASE(21)
    DC_TQ_CASE(24)
    DC_TQ_CASE(27)
#undef DC_TQ_CASE
    default:
      return DC_ERR_UNSUPPORTED;
  }
}

}  // namespace dc
