// kernels.h — argument blocks and launchers of the CUDA kernels behind libmixvae_b200.
#pragma once
#include "common.cuh"

namespace mvae {

struct BnOff { int64_t off[6]; };

// ---- narrow dense layers ---------------------------------------------------------------------
struct DenseFwdArgs {
  const float* in;  int64_t in_arm_stride;    // [A][B][nin]
  float* out;       int64_t out_arm_stride;   // [A][B][nout]
  const float* params; int64_t p_arm_stride, offW, offB;
  int B, nin, nout;
  int bn_mode;                // 0: none, 1: batch statistics from bn_sums_in, 2: given mean/rstd (eval)
  const double* bn_sums_in;   // [A][2][128] column sums / sums of squares of `in`
  float* bn_mean; float* bn_rstd;  // [A][128] finalised statistics of `in` (written when bn_mode==1)
  double* stats_out;          // [A][2][128] accumulators for `out` or nullptr
  float eps; int relu;
};
int launch_dense_fwd(const DenseFwdArgs& a, int A, cudaStream_t s);

struct DenseBwdArgs {
  const float* g_out; const float* act_out; float* delta; float* g_in;
  const float* params; int64_t p_arm_stride, offW;
  int B, nin, nout;
  int bn_out; const double* bnb_sums; const float* mean_out; const float* rstd_out;
  int bn_in; const float* act_in; const float* mean_in; const float* rstd_in; double* bnb_sums_next;
  float* delta_t; int64_t delta_t_ld, delta_t_arm_stride;
};
int launch_dense_bwd(const DenseBwdArgs& a, int A, cudaStream_t s);
int launch_dense_fwd_mma(const DenseFwdArgs& a, int A, cudaStream_t s);   // warp-MMA 3xTF32 versions (kernels_mma.cu)
int launch_dense_bwd_mma(const DenseBwdArgs& a, int A, int split3, cudaStream_t s);

struct Fc1EpiArgs {
  const float* part; int64_t split_stride, arm_stride, ld; int nsplit;
  const float* params; int64_t p_arm_stride, offB;
  float* out; double* stats_out; int B, H;
};
int launch_fc1_epilogue(const Fc1EpiArgs& a, int A, cudaStream_t s);

int launch_bn_eval_prep(const float* bn_running, int64_t bn_stride, BnOff off, float* bn_mean, float* bn_rstd,
                        int A, int H, int L, float eps, cudaStream_t s);
int launch_bn_update_running(float* bn_running, int64_t bn_stride, BnOff off, int64_t* nbt, const double* bn_sums,
                             int A, int B, int H, int L, float momentum, cudaStream_t s);

// ---- categorical / state heads ---------------------------------------------------------------
struct HeadArgs {
  int A, At, arm_off, B, H, L, C, S;
  const float* params; int64_t p_arm_stride;
  int64_t oWc, oBc, oWmu, oBmu, oWsig, oBsig, oW6, oB6;
  // forward inputs
  const float* a5;            // [A][B][L]
  int bn_mode;                // 1 batch sums, 2 given
  const double* bn_sums5;     // [A][2][128]
  float* bn_mean5; float* bn_rstd5;  // [A][128]
  const uint8_t* cat_mask;    // [C] 1 = category kept, or nullptr (forward(mask=...), nn_model.py:332-335)
  const float* U; const float* E; const uint8_t* keep_s;   // U / E nullptr: in-kernel counter-based draws
  const uint64_t* ukeys; const uint64_t* ekeys;   // device tables of generator keys per LOCAL arm (Work::keys, streams 1 / 2)
  float tau, temp, eps, s_scale; int hard, training;
  // forward outputs
  float *x_low, *c_prob, *qc, *c_smp, *s_mean, *s_logvar, *s_smp;
  float *ysoft, *svar, *yy, *zc, *d6;
  double* kl_sums;            // [A][16]: sum_b(1+lv-mu^2-e^lv) per state dim
  // optional (training): running-statistics update of batch_l1..l5 by block 0 of every arm (nn.BatchNorm1d momentum
  // update, unbiased variance; replaces a separate launch).  bn_running == nullptr: skipped.
  float* bn_running; int64_t bn_stride; BnOff bn_off; int64_t* nbt; const double* bn_sums_all; float momentum;
  // backward-only
  const float* g_d6;          // [A][B][L]
  const float* gdiff;         // [A][B][C]  sum_b (r_a - r_b) of the local arms
  const float* colc;          // [A][4][128]  w, cvar, mean, T
  float kl_coef, ent_coef, g_coef;   // max(At-1,1)*beta/B, (At-1)/B, 2*lam/B
  float *delta6, *delta_mu, *delta_sig, *delta_z, *g_xlow;
  double* bnb_sums5;          // [A][2][128]
};
int launch_head_fwd(const HeadArgs& a, cudaStream_t s);
int launch_head_bwd(const HeadArgs& a, cudaStream_t s);

// ---- loss ------------------------------------------------------------------------------------
struct CouplingArgs {
  int A, At, arm_off, B, C;
  const float* qc_all;    // [At][B][C]
  const float* csmp_all;  // [At][B][C]
  double* acc;            // acc_loss block
  float* gdiff;           // [A][B][C]  sum_b (r_a - r_b) of the local arms, r = log(q+eps)*w
  float* wcat;            // [At][128]
  float eps, lam;
};
int launch_qstats(const CouplingArgs& a, cudaStream_t s);
int launch_coupling_rows(const CouplingArgs& a, cudaStream_t s);

struct LossFinalArgs {
  int A, At, arm_off, B, D, C, S;
  const double* acc_loss; const double* kl_sums;
  float* colc;            // [A][4][128]
  float* loss_out;
  float eps, lam, beta;
};
int launch_loss_finalize(const LossFinalArgs& a, cudaStream_t s);

// elementwise reconstruction loss (+ gradient) on a materialised pre-activation (SIMT path / x_rec)
struct ReconElemArgs {
  float* pre;             // [A][B][D] in: h10*W11^T (no bias); out: dY (if want_grad)
  const float* x; int64_t x_arm_stride, x_row_stride;
  const float* params; int64_t p_arm_stride, offB;
  float* x_rec;           // optional [A][B][D]
  double* recon_acc;      // acc_loss block (accl_recon) or nullptr
  int B, D; float gscale; int want_grad;
};
int launch_recon_elem(const ReconElemArgs& a, int A, cudaStream_t s);
int launch_colsum(const float* src, int64_t src_arm_stride, float* dst_base, int64_t dst_arm_stride, int B, int D,
                  int A, cudaStream_t s);

// ---- weight gradients of the narrow layers ---------------------------------------------------
struct WgProblem {
  int64_t delta_off, in_off;           // offsets into work (arm 0)
  int64_t delta_arm_stride, in_arm_stride;
  int nout, nin, in_ld;
  int bn_layer;                        // -1: raw input, else normalise the input with bn stats of that layer
  int64_t poffW, poffB;                // param offsets (grads written there)
  int nsplit, cta_begin;               // row splits of this problem / its first CTA (filled by launch_wgrad_mma)
};
struct WgArgs {
  WgProblem prob[13];
  int nprob, A, B, rows_per_split, nsplit;
  const float* work; const float* bn_mean; const float* bn_rstd;
  float* part; int64_t part_split_stride, part_arm_stride, base_off;
  float* grads; int64_t g_arm_stride;
};
int launch_wgrad(const WgArgs& a, cudaStream_t s);
// side branch for the weight-gradient group of the fused step: the thin problems (wgrad2) run beside the grouped tcgen05
// GEMM, the fixed-order sum of the partials beside the fc1 weight gradient (the caller joins reduce_done before Adam)
struct WgFork {
  cudaStream_t side;
  cudaEvent_t fork_ev, thin_done, wide_done, reduce_done;
};
int launch_wgrad_mma(const WgArgs& a, int split3, cudaStream_t s, const WgFork* fork = nullptr);
// decoder stack fc7..fc10 fused per direction (kernels_chain.cu)
int launch_dec_chain_fwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* h6, float* const hout[4], int split3, cudaStream_t s);
int launch_dec_chain_bwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* g10, const float* const act[4], float* const delta[4], float* g6, int split3,
                         cudaStream_t s);

// encoder middle fc2..fc5 (+ batch_l1..l4 folded in, column sums for batch_l2..l5) as ONE cooperative kernel; returns 1 when
// the shape does not fit a single co-resident wave (the caller then runs the per-layer kernels)
// fc1's stream-K partial tiles handed to the encoder chain, which then forms a1 = relu(scale * sum + b1) itself (first phase of
// the single-tile chain kernel) instead of a separate fix-up launch: ts_gemm.cu fills it, kernels_chain.cu consumes it
struct Fc1Deferred {
  const float* part; int batch, ktiles; int64_t U, G; float scale; int64_t offB; int valid;
};
// 1: launch_enc_chain_fwd will run the one-tile-per-CTA cooperative kernel for this shape (and can take an Fc1Deferred)
int enc_chain_fwd_is_single(int A, int B, int H, int L);
int launch_enc_chain_fwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* a1, float* const aout[4], double* acc_fwd, float* bn_mean, float* bn_rstd, float eps, const Fc1Deferred* fc1,
                         cudaStream_t s);

// encoder middle backward (batch_l5 .. batch_l1 and fc5 .. fc2): deltas of layers 5..1, one cooperative kernel
int launch_enc_chain_bwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* g_xlow, const float* const act[5], float* const delta[5], double* acc_bwd,
                         const float* bn_mean, const float* bn_rstd, float* g_scratch, int split3, cudaStream_t s);

// ---- optimiser / misc ------------------------------------------------------------------------
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                float wd, int adamw, int64_t step, const uint64_t* step_dev, cudaStream_t s);
int launch_adam_peer(float* const* peer_params, const float* const* peer_grads, float* m, float* v, int64_t n, int rank,
                     int world, float lr, float b1, float b2, float eps, int64_t step, const uint64_t* step_dev,
                     cudaStream_t s);
int launch_scale(float* p, int64_t n, const float* scale_dev, cudaStream_t s);
int launch_counter_inc(uint64_t* counter, cudaStream_t s);
int launch_unpack_rows(const uint32_t* bitmap, const float* values, const int64_t* row_ptr, int64_t rows, int D, float* out,
                       int64_t out_ld, cudaStream_t s);
int launch_argmax(const float* q, int32_t* labels, int64_t rows, int cols, cudaStream_t s);
int launch_confmat(const int32_t* labels, int64_t n, int A, int K, int32_t* counts, cudaStream_t s);
int launch_transpose(const float* src, int64_t src_ld, int64_t src_batch_stride, float* dst, int64_t dst_ld,
                     int64_t dst_batch_stride, int rows, int cols, int batch, cudaStream_t s);

// ---- gene-dimension GEMMs, fp32 SIMT (precision==3 or shapes the tensor-core path rejects) ----
struct GemmArgs {
  const float* A; int64_t sAm, sAk, A_batch;   // A(m,k) = A[m*sAm + k*sAk]
  const float* Bm; int64_t sBk, sBn, B_batch;  // B(k,n)
  float* C; int64_t sCm, sCn, C_batch;
  int M, N, K;
  DropSpec drop; int drop_operand;             // 0 none, 1: A is x with (row,col)=(m,k), 2: B is x with (row,col)=(k,n)
};
int launch_sgemm_simt(const GemmArgs& a, int batch, cudaStream_t s);
int launch_dropout_mask(const DropSpec& d, uint64_t key, int B, uint8_t* out, cudaStream_t s);
// generator keys of one step -> keys[kNumStreams][MVAE_MAX_ARMS] (local arm a uses global arm a + arm_off), counters -> keys[3*16 .. +2).
// counters (nullable): device {rng step, Adam step}; when given the kernel INCREMENTS counters[0] (and counters[1] if
// bump_adam) and uses the new values instead of step_host -- the form a replayed CUDA graph needs.
int launch_step_prep(uint64_t seed, uint64_t step_host, uint64_t* counters, int bump_adam, int arm_off, uint64_t* keys_out,
                     cudaStream_t s);

}  // namespace mvae
