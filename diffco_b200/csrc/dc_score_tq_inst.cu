// One translation unit per (radial kind, class width, mode): compiled several times by diffco_b200/build.py with
//   -DDC_TQ_KIND=<KR_*> -DDC_TQ_CW=<1|4> -DDC_TQ_MODE=<M_*> -DDC_TQ_NAME=<symbol>
// so the thread-per-query instantiations build in parallel.
#include "dc_score_tq.cuh"

#ifndef DC_TQ_KIND
#error "compile with -DDC_TQ_KIND -DDC_TQ_CW -DDC_TQ_MODE -DDC_TQ_NAME"
#endif

namespace dc {

// Q (queries per lane): 2 where the register file allows it, 1 for the wide multi-class modes.
template <int FP>
static int launch_fp(ScoreArgs<float>& a, int num_sms, cudaStream_t stream) {
  constexpr int CW = DC_TQ_CW;
  constexpr int MODE = DC_TQ_MODE;
  constexpr int NG = (MODE == M_JAC) ? CW : (MODE == M_GRAD ? 1 : 0);
  constexpr int Q = ((2 + NG) * FP * 2 * 2 + 2 * CW * 2 > 96) ? 1 : 2;
  return launch_score_tq<FP, DC_TQ_KIND, CW, MODE, Q, 16, 3>(a, num_sms, stream);
}

int DC_TQ_NAME(int fp, ScoreArgs<float>& a, int num_sms, cudaStream_t stream) {
  switch (fp) {
    case 1: return launch_fp<1>(a, num_sms, stream);
    case 2: return launch_fp<2>(a, num_sms, stream);
    case 3: return launch_fp<3>(a, num_sms, stream);
    case 4: return launch_fp<4>(a, num_sms, stream);
    case 6: return launch_fp<6>(a, num_sms, stream);
    case 7: return launch_fp<7>(a, num_sms, stream);
    case 8: return launch_fp<8>(a, num_sms, stream);
    case 11: return launch_fp<11>(a, num_sms, stream);
    case 12: return launch_fp<12>(a, num_sms, stream);
    default: return DC_ERR_UNSUPPORTED;
  }
}

}  // namespace dc
