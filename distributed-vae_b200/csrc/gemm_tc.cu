// gemm_tc.cu — placeholder until the tcgen05 kernels land: reports "unsupported" so that api.cu
// takes the fp32 SIMT kernels.
#include "gemm_tc.h"

namespace mvae {
bool gemm_tc_supported(int, int, int) { return false; }
int tc_fc1_forward(const mvae_dims&, const mvae_hparams&, const mvae_state&, const mvae_inputs&, const DropSpec&,
                   const Work&, cudaStream_t, Fc1EpiArgs*) { set_error("tensor-core path not built"); return -3; }
int tc_fc11_loss_grad(const mvae_dims&, const mvae_hparams&, const mvae_state&, const mvae_inputs&, const Work&, float,
                      int, cudaStream_t) { set_error("tensor-core path not built"); return -3; }
int tc_fc1_wgrad(const mvae_dims&, const mvae_hparams&, const mvae_state&, const mvae_inputs&, const DropSpec&,
                 const Work&, cudaStream_t) { set_error("tensor-core path not built"); return -3; }
}  // namespace mvae
