// placeholder until the tcgen05 kernels land
#include "kernels.h"
namespace rnvp {
int k_conv_fwd_tf32(const ConvArgs& a, cudaStream_t st) { (void)a; (void)st; set_error("tf32 conv not built"); return RNVP_ERR_INVALID; }
int k_conv_wgrad_tf32(const WgradArgs& a, cudaStream_t st) { (void)a; (void)st; set_error("tf32 wgrad not built"); return RNVP_ERR_INVALID; }
}
