// Instantiations of the lane-split score kernel for float (see dc_score_ls.cuh).
#define DC_LS_INSTANTIATE
#include "dc_score_ls.cuh"

namespace dc {
int ls_launch_f32(LsArgs<float>& a, int num_sms, cudaStream_t stream) { return launch_score_ls<float>(a, num_sms, stream); }
}  // namespace dc
