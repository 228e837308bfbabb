/*
 * mixvae_b200.h — C ABI of libmixvae_b200.so: the coupled mixture-VAE (cpl-mixVAE / MMIDAS)
 * training step as hand-written sm_100a CUDA kernels.
 *
 * The reference (AllenInstitute/distributed-vae) has no FFI for this path: its boundary is the
 * Python class API of mmidas/nn_model.py (mixVAE_model.forward :297, .loss :495), the autograd
 * backward at mmidas/cpl_mixvae.py:462 and torch.optim.Adam.step at :463.  Each entry point below
 * names the reference call it stands in for; distributed-vae_b200/mmidas_b200/ is the Python
 * mirror that binds them with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every device pointer is owned by the caller (PyTorch) and
 *     borrowed for the duration of a call; the library allocates and frees nothing on the device.
 *   - all work is enqueued on the cudaStream_t passed in (as void*); no host synchronisation, so
 *     every call is CUDA-graph capturable.
 *   - return value: 0 = ok, <0 = argument/shape/arch error, >0 = cudaError_t.  Text of the last
 *     error of the calling thread: mvae_last_error().
 *   - fp32 storage everywhere (the reference is fp32); row-major; weights are [out,in] like
 *     nn.Linear.
 */
#ifndef MIXVAE_B200_H_
#define MIXVAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVAE_ABI_VERSION 2
#define MVAE_N_PARAM_TENSORS 28 /* 14 Linear layers x {weight, bias} per arm */
#define MVAE_MAX_ARMS 16

/* Shapes.  mixVAE_model.__init__ arguments (nn_model.py:112-134). */
typedef struct mvae_dims {
  int32_t n_arm;        /* A  n_arm                          */
  int32_t batch;        /* B  cells per step                 */
  int32_t input_dim;    /* D  genes                          */
  int32_t fc_dim;       /* H  hidden width          (<=128)  */
  int32_t lowD_dim;     /* L  latent width          (<=32)   */
  int32_t n_categories; /* C  categories            (<=128)  */
  int32_t state_dim;    /* S  continuous state dim  (<=8)    */
  int32_t n_arm_total;  /* arms in the whole model (>= n_arm; > n_arm when arms are sharded over ranks) */
  int32_t arm_offset;   /* index of this rank's first arm in the whole model */
} mvae_dims;

/* Scalars.  mixVAE_model.__init__ (nn_model.py:160-178) + forward(temp) + Adam defaults. */
typedef struct mvae_hparams {
  float tau, temp, beta, lam, eps, momentum, x_drop, s_drop;
  int32_t hard;       /* straight-through one-hot Gumbel sample (nn_model.py:486-493) */
  int32_t precision;  /* 0: fc1 3xTF32 + TF32 elsewhere (default), 1: 3xTF32 in every gene GEMM,
                         2: plain TF32 everywhere, 3: fp32 SIMT kernels (no tensor cores) */
} mvae_hparams;

/* Per-arm flat parameter layout, in floats.  Tensor order = the reference's per-arm parameter
 * order: fc1.w fc1.b fc2.w ... fc5.b fcc.w fcc.b fc_mu.w fc_mu.b fc_sigma.w fc_sigma.b fc6.w ...
 * fc11.w fc11.b.  Arm a's tensor t lives at params + a*arm_stride + offset[t]. */
typedef struct mvae_layout {
  int64_t offset[MVAE_N_PARAM_TENSORS];
  int64_t numel[MVAE_N_PARAM_TENSORS];
  int64_t arm_stride;   /* floats per arm incl. padding (multiple of 256)                     */
  int64_t bn_stride;    /* floats per arm of BN running stats: l1..l5,s each {mean,var}       */
  int64_t bn_offset[6]; /* offset of running_mean of batch_l1..l5, batch_s; var follows mean  */
  int64_t work_floats;  /* workspace size in floats for (dims)                                */
} mvae_layout;

/* Device buffers that persist across steps (all owned by the caller). */
typedef struct mvae_state {
  float* params;        /* [A][arm_stride]                                           */
  float* grads;         /* [A][arm_stride]   written (not accumulated) by backward   */
  float* adam_m;        /* [A][arm_stride]                                           */
  float* adam_v;        /* [A][arm_stride]                                           */
  float* bn_running;    /* [A][bn_stride]    running_mean / running_var             */
  int64_t* bn_batches;  /* [A][6]            num_batches_tracked                     */
  float* work;          /* [work_floats]     activations + partials                  */
} mvae_state;

/* Per-step inputs.  The reference draws its noise with torch's RNG inside forward (nn_model.py:264 dropout, :440 Gumbel
 * uniforms, :427 state noise) and has no injection hook; here every noise tensor may be injected (parity tests) or left
 * NULL, in which case it is drawn in-kernel by a counter-based generator keyed on (seed, step, global arm index, stream):
 * arms sharded over ranks (arm_offset > 0) draw the streams of their global index, and a step counter that lives on the
 * device (counters) lets a captured CUDA graph draw fresh noise on every replay. */
typedef struct mvae_inputs {
  const float* x;          /* [A or 1][B][D]                                                  */
  int64_t x_arm_stride;    /* floats between arms' inputs; 0 = all arms share x (x.expand)     */
  int64_t x_row_stride;    /* floats between rows (>= D)                                       */
  const float* U;          /* [A][B][C] uniforms for the Gumbel noise (nn_model.py:440), or NULL */
  const float* E;          /* [A][B][S] uniforms for the state sample (nn_model.py:427), or NULL */
                           /* NULL: drawn in-kernel, counter-based on (seed, step, arm, cell, k) */
  const uint8_t* keep_x;   /* [A][B][D] input-dropout keep mask or NULL                        */
  const uint8_t* keep_s;   /* [A][B][S] state-dropout keep mask or NULL (s_drop == 0)          */
  const uint8_t* cat_mask; /* [C] 1 = category kept, or NULL: forward(mask=...) of the pruning path (nn_model.py:332-335):
                              q = softmax(c_prob[:, kept] / tau) on the kept categories, exactly 0 elsewhere */
  uint64_t seed;           /* in-kernel generators                                             */
  uint64_t step;           /* step index of this forward (ignored when counters != NULL)       */
  uint64_t* counters;      /* NULL, or device {forward counter, Adam step counter}: mvae_forward increments
                              counters[0] and uses the new value as `step`; mvae_train_step also increments
                              counters[1] and uses it as Adam's step (CUDA-graph replay)        */
  int32_t training;        /* 1: batch-stat BN, dropout, Gumbel; 0: mixVAE_model.forward(eval=True) on .eval() */
} mvae_inputs;

/* Forward outputs, the tensors mixVAE_model.forward returns (nn_model.py:368); [A][B][.] each. */
typedef struct mvae_outputs {
  float* x_low;     /* [A][B][L] */
  float* c_prob;    /* [A][B][C] */
  float* qc;        /* [A][B][C] */
  float* c_smp;     /* [A][B][C] */
  float* s_mean;    /* [A][B][S] */
  float* s_logvar;  /* [A][B][S] */
  float* s_smp;     /* [A][B][S] */
  float* x_rec;     /* [A][B][D] or NULL: the reconstruction is only materialised on request */
} mvae_outputs;

/* Layout of the loss vector written by mvae_loss (floats):
 *   [0] total  [1] joint  [2] neg. joint entropy (avg over pairs)  [3] simplex distance (avg)
 *   [4] l2 distance of samples (avg)  [5 .. 5+At) rec per arm  [5+At .. 5+2At) KL per arm
 *   [5+2At .. 5+3At) log-likelihood metric per arm;  At = n_arm_total.
 * (mixVAE_model.loss return tuple, nn_model.py:588-598) */
#define MVAE_LOSS_FLOATS(At) (5 + 3 * (At))

const char* mvae_last_error(void);
int mvae_abi_version(void);

/* Shape bookkeeping (host only). */
int mvae_compute_layout(const mvae_dims* dims, mvae_layout* out);

/* mixVAE_model.forward (nn_model.py:297-368): encoder, categorical head, Gumbel-softmax, state
 * head, decoder stack up to fc10; x_rec = relu(fc11(.)) only if out->x_rec != NULL. */
int mvae_forward(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st,
                 const mvae_inputs* in, const mvae_outputs* out, void* stream);

/* mixVAE_model.loss (nn_model.py:495-598).  Fuses the fc11 GEMM with the reconstruction loss and,
 * if want_grad, with its own backward (d fc11.weight, d fc11.bias, d h10), and the coupling terms
 * with the first half of their gradient.  qc_all: [n_arm_total][B][C] posteriors of every arm of
 * the model in global arm order (== out->qc when arms are not sharded); c_smp_all likewise.
 * loss_out: MVAE_LOSS_FLOATS(n_arm_total) floats on the device. */
int mvae_loss(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st,
              const mvae_inputs* in, const mvae_outputs* out, const float* qc_all,
              const float* c_smp_all, float* loss_out, int want_grad, void* stream);

/* Tensor.backward of the loss (cpl_mixvae.py:462): the rest of the backward chain; fills
 * st->grads.  grad_scale: device pointer to the upstream gradient of the total loss (NULL = 1). */
int mvae_backward(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st,
                  const mvae_inputs* in, const mvae_outputs* out, const float* grad_scale,
                  void* stream);

/* torch.optim.Adam.step (cpl_mixvae.py:463, defaults of :274) over a flat buffer.
 * step: 1-based index of this update (already incremented).  step_counter: NULL, or a device counter that is
 * incremented on the stream and then used instead of `step` (CUDA-graph replay). */
int mvae_adam(float* params, const float* grads, float* m, float* v, int64_t n, float lr,
              float beta1, float beta2, float eps, float weight_decay, int32_t adamw,
              int64_t step, uint64_t* step_counter, void* stream);

/* Data-parallel replicas (the reference's FSDP wrap, train.py:140-143, would reduce-scatter the gradients and all-gather
 * the parameters around torch.optim.Adam.step, cpl_mixvae.py:463): gradient averaging fused with Adam over peer memory.
 * peer_params / peer_grads: HOST arrays of `world` device pointers, entry q = the flat parameter / gradient buffer of
 * replica q (peer-accessible: CUDA IPC / symmetric memory; entry `rank` is this replica's own).  Rank `rank` sums its
 * 1/world share of all replicas' gradients in rank order, divides by `world`, applies Adam with its share of m / v and
 * stores the new parameters into every replica's buffer.  The caller orders the call between two cross-replica barriers
 * on `stream`: every replica's gradients final before, every replica's parameters final after.  step / step_counter
 * as for mvae_adam. */
int mvae_adam_peer(float* const* peer_params, const float* const* peer_grads, float* m, float* v, int64_t n, int32_t rank,
                   int32_t world, float lr, float beta1, float beta2, float eps, int64_t step, uint64_t* step_counter,
                   void* stream);

/* zero_grad + forward + loss + backward + Adam in one call (cpl_mixvae.py:434-463). */
int mvae_train_step(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st,
                    const mvae_inputs* in, const mvae_outputs* out, float* loss_out, float lr,
                    float beta1, float beta2, float adam_eps, int64_t step, void* stream);

/* zero_grad + forward + loss + backward in one call, WITHOUT the optimiser step: the step of a data-parallel replica
 * (all arms local), whose gradients are then averaged and applied by mvae_adam_peer (or an all-reduce + mvae_adam).
 * Same kernels and side branches as mvae_train_step; in->counters[0] is advanced as by mvae_forward. */
int mvae_grad_step(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st,
                   const mvae_inputs* in, const mvae_outputs* out, float* loss_out, void* stream);

/* Device-side argmax of q(c|x) -> int32 labels [A][B] (replaces the per-step D2H of
 * cpl_mixvae.py:476 + mmidas/_utils.py:78 classify). */
int mvae_argmax(const float* q, int32_t* labels, int64_t rows, int32_t cols, void* stream);

/* Device-side confusion counts between the arms' labels (mmidas/_utils.py:83 compute_confmat, used by the
 * consensus bookkeeping cpl_mixvae.py:512-523, :640-657, :750-763): labels [n_arm][n_cells] int32 (mvae_argmax output),
 * counts [n_pairs][K][K] int32 with pairs (a < b) in order; counts are ACCUMULATED (zero them first), so the
 * labels of an epoch never leave the device: only n_pairs*K*K integers do. */
int mvae_confmat(const int32_t* labels, int64_t n_cells, int32_t n_arm, int32_t n_categories, int32_t* counts, void* stream);

/* The keep-mask [A][B][D] that the in-kernel dropout generator applies for (in->seed, in->step):
 * the fc1 forward and fc1 weight-gradient kernels regenerate it on the fly instead of reading it. */
int mvae_dropout_mask(const mvae_dims* dims, const mvae_hparams* hp, const mvae_inputs* in,
                      uint8_t* keep_out, void* stream);

/* The host->device hand-over of a batch (`x = x.to(rank)`, cpl_mixvae.py:416) for a row-packed host batch: expression
 * matrices are sparse (Smart-seq ~35 % non-zeros, 10x ~8 %), so the host keeps a batch as a bitmap (bit j of word w of a
 * row = gene 32 w + j is non-zero; ceil(n_cols / 32) uint32 words per row), the non-zero fp32 values in row-major order and
 * row_ptr[rows + 1] (int64 offsets into values).  This expands the three device arrays into the dense [rows][n_cols] fp32
 * matrix, bit-exactly (zeros come back as +0.0f). */
int mvae_unpack_rows(const uint32_t* bitmap, const float* values, const int64_t* row_ptr, int64_t rows, int32_t n_cols,
                     float* out, int64_t out_row_stride, void* stream);

/* Number of kernels launched by the library in this process (for bench.py's gpu_launches). */
int64_t mvae_launch_count(void);

/* Programmatic dependent launch between the library's own kernels (on by default): a kernel launched right behind another
 * kernel of the same call carries the programmatic-stream-serialization attribute (a programmatic edge in a captured
 * graph) and runs its set-up and the loads of step constants while its predecessor drains; results do not depend on it.
 * mvae_pdl_enable(0) launches every kernel fully ordered (diagnostics, tests); returns the previous setting.  Affects
 * launches made (and graphs captured) afterwards, process-wide. */
int mvae_pdl_enable(int on);

/* Test hook for the tcgen05 GEMM kernel behind the gene-dimension layers:
 * C[split] (M x N, ldc) = A . B over K with fp32 storage and TF32 (flags 0) or error-compensated
 * 3xTF32 (flags 1|2) tensor-core math.  a_mn / b_mn: 0 = operand stored [M or N][K] (K contiguous),
 * 1 = stored [K][M or N].  Pitches in floats (multiples of 4).  Split-K partials land
 * c_split_stride floats apart; the caller sums them. */
int mvae_debug_tc_gemm(const float* A, int a_mn, int64_t a_pitch, const float* B, int b_mn, int64_t b_pitch,
                       int M, int N, int K, int BN, int nsplit, int flags, float* C, int64_t ldc,
                       int64_t c_split_stride, void* stream);

/* ---- Augmenter forward: mmidas/augmentation/udagan.py:217-329 (Augmenter_smartseq.forward) as called by the
 * training loop in eval mode (cpl_mixvae.py:184, :422-423).  With running statistics every `relu(batch_fcN(fcN(x)))` is a
 * Linear followed by a per-column affine and an activation:
 *   mvae_fold_affine : scale = gamma / sqrt(var + eps), shift = (bias - mean) * scale + beta   (null mean/var: shift = bias)
 *   mvae_linear_act  : y = act((x . w^T) * scale + shift) on tcgen05 (TF32, or error-compensated 3xTF32 when split3 != 0);
 *                      x [rows][k] and w [n_out][k] row-major, pitches multiples of 4 floats (pad columns hold zeros),
 *                      act: 0 none, 1 relu, 2 elu, 3 sigmoid
 *   mvae_fma_rows    : out = a_scale * a * b + c on [rows][n] views (reparam_trick, aug_utils.py:51-65; b / c may be null)
 * replacing nn.Linear / nn.BatchNorm1d(eval) / F.relu / F.elu / torch.sigmoid of the reference. */
int mvae_fold_affine(const float* bias, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                     int32_t n, float* scale, float* shift, void* stream);
int mvae_linear_act(const float* x, int64_t x_pitch, const float* w, int64_t w_pitch, float* y, int64_t y_pitch, int64_t rows,
                    int32_t n_out, int32_t k, const float* scale, const float* shift, int32_t act, int32_t split3, void* stream);
int mvae_fma_rows(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, float* out, int64_t ldo,
                  int64_t rows, int32_t n, float a_scale, void* stream);

/* Device-time accounting for bench.py's roofline: when enabled, the library brackets each kernel
 * group with CUDA events on the launching stream.  Groups (index into ms_out / count_out):
 *   0 fc1 forward  1 fc11 fused loss+grad  2 fc1 weight gradient  3 narrow layers forward
 *   4 narrow layers backward  5 coupling loss  6 narrow weight gradients  7 Adam.
 * mvae_timing_enable(on) resets the records; mvae_timing_read synchronises on the recorded events. */
#define MVAE_TIMING_GROUPS 8
int mvae_timing_enable(int on);
int mvae_timing_read(float* ms_out, int32_t* count_out, int32_t n_groups);

#ifdef __cplusplus
}
#endif
#endif /* MIXVAE_B200_H_ */
