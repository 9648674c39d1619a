// Instantiations of the lane-split score kernel for double (see dc_score_ls.cuh).
#define DC_LS_INSTANTIATE
#include "dc_score_ls.cuh"

namespace dc {
int ls_launch_f64(LsArgs<double>& a, int num_sms, cudaStream_t stream) { return launch_score_ls<double>(a, num_sms, stream); }
}  // namespace dc
