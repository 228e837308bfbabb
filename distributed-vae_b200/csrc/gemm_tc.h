// gemm_tc.h — tcgen05/TMA kernels for the gene-dimension GEMMs (fc1 forward, fc11 fused
// forward+loss+backward, fc1 weight gradient).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace mvae {

// true when the tensor-core kernels can run this shape (TMA needs 16-byte aligned row pitches)
bool gemm_tc_supported(int B, int D, int H);

// fc11 GEMM fused with the reconstruction loss and (want_grad) d fc11.weight / d fc11.bias / d h10.
int tc_fc11_loss_grad(const mvae_dims& d, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                      const Work& w, float gscale, int want_grad, cudaStream_t s, bool defer_gene_fix = false);

// d fc1.weight = delta1^T * dropout(x)
int tc_fc1_wgrad(const mvae_dims& d, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                 const DropSpec& drop, const Work& w, cudaStream_t s);

// ts_gemm.cu: stream-K kernels with the x operand transformed in registers and fed to the MMA from tensor memory
int64_t ts_part_floats(int A, int Bpad, int Dpad);
int ts_fc1_forward(const mvae_dims& d, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                   const DropSpec& drop, const Work& w, float* a1_out, double* stats_out, Fc1Deferred* defer, cudaStream_t s);
int ts_fc1_wgrad(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const DropSpec& drop, const Work& w,
                 cudaStream_t s);

// fc11_ts.cu: second-generation fused fc11 passes (resident operand and dY in tensor memory, stream-K)
int ts_fc11_rows(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale, int want_grad,
                 float* x_rec, double* recon_acc, cudaStream_t s);
// defer_gene_fix: only the row pass's partials (d h10) are summed on s; ts_fc11_gene_fixup then sums the gene pass's
// (d fc11.weight, d fc11.bias) on the stream it is given -- the side branch of the fused step (returns 1 if nothing is pending)
int ts_fc11_loss_grad(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale,
                      double* recon_acc, cudaStream_t s, bool defer_gene_fix = false);
int ts_fc11_gene_fixup(cudaStream_t s);
// gemm_tc.cu: grouped tcgen05 weight-gradient GEMM of the wide narrow-layer problems (delta^T . bn(input), bias gradient
// through a column of ones); partials in the layout of wgrad_reduce2_kernel
bool tc_narrow_wgrad_ok(const WgArgs& a, const WgProblem& q);
int tc_narrow_wgrad(WgArgs& a, const int* idx, int n, int split3, cudaStream_t s, bool share_sm = false);
// augmenter forward pieces (udagan.py:217-329, eval mode): folded BatchNorm/bias affine, Linear with fused affine + activation
// epilogue on tcgen05, and the row-wise fma used by the reparameterisation
int launch_fold_affine(const float* bias, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                       int n, float* scale, float* shift, cudaStream_t s);
int launch_fma_rows(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, float* out, int64_t ldo,
                    int64_t rows, int n, float a_scale, cudaStream_t s);
int tc_linear_act(const float* x, int64_t x_pitch, const float* w, int64_t w_pitch, float* y, int64_t y_pitch, int64_t rows,
                  int n_out, int k, const float* scale, const float* shift, int act, int split3, cudaStream_t s);

}  // namespace mvae
