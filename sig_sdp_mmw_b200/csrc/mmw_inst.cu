// One (sketch dtype, lanes-per-row) instantiation of the kernels; compiled eight times by
// sig_sdp_mmw_b200/build.py with -DSIGSDP_T=<double|float> -DSIGSDP_G=<4|8|16|32>
// -DSIGSDP_NAME=ks_<f64|f32>_g<G>.
#include "mmw_kernels.cuh"

namespace sigsdp {
using L = Launchers<SIGSDP_T, SIGSDP_G>;
const KernelSet SIGSDP_NAME = {L::prepare, L::fused, L::dual, L::exp_, L::loss, L::term,
                               L::copy,    L::gram,  L::record, L::batch, L::batch_occupancy};
}  // namespace sigsdp
