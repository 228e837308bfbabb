// api.cu — the C ABI (include/mixvae_b200.h): sequencing of the kernels of one cpl-mixVAE step.
//
// Step order follows mmidas/cpl_mixvae.py:434-463: zero_grad (implicit: every gradient is written,
// never accumulated), forward, loss, backward, Adam.
#include <math.h>

#include "common.cuh"
#include "gemm_tc.h"
#include "kernels.h"

namespace mvae {

struct Plan {
  mvae_dims d;
  mvae_layout L;
  Work w;
  int A, At, B, D, H, Ld, C, S;
};

static int make_plan(const mvae_dims* dims, Plan* p) {
  MVAE_CHECK_ARG(dims != nullptr, "dims is null");
  p->d = *dims;
  int rc = compute_layout(*dims, &p->L);
  if (rc) return rc;
  p->w = make_work(*dims);
  p->A = dims->n_arm; p->At = dims->n_arm_total; p->B = dims->batch; p->D = dims->input_dim;
  p->H = dims->fc_dim; p->Ld = dims->lowD_dim; p->C = dims->n_categories; p->S = dims->state_dim;
  return 0;
}

static int check_device() {
  tl_pdl = 0;            // every entry point passes through here: its first launch is an ordinary one
  static int ok = -1;
  if (ok < 0) {
    int dev = 0;
    MVAE_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MVAE_CUDA(cudaGetDeviceProperties(&prop, dev));
    ok = (prop.major == 10) ? 1 : 0;
    if (!ok) set_error("libmixvae_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
  }
  return ok == 1 ? 0 : -2;
}

// Programmatic dependent launch (see common.cuh).  mvae_pdl_enable(0) turns the attribute off (the test suite compares
// both ways); per-group timing (events between the launches) measures the serial order and runs without it.
thread_local int tl_pdl = 0;
static int g_pdl_on = 1;
bool pdl_enabled() { return g_pdl_on && !timing_enabled(); }

// Side branch of the fused step (mvae_train_step): the coupling kernels run beside the decoder chain and the first fc11
// pass, the loss finalisation beside the decoder's backward chain -- small latency-bound kernels that need no SM the main
// branch can use at those points.  Fork and join are event edges, so the step still captures into one CUDA graph.
// One set of resources per host thread and device, created outside any stream capture (every entry point passes through
// check_device first; a step captured before the resources exist simply runs unforked).
struct SideBranch {
  cudaStream_t side = nullptr;
  cudaEvent_t head_done = nullptr, qstats_done = nullptr, fc11_done = nullptr, final_done = nullptr;
  cudaEvent_t wg_fork = nullptr, wg_thin = nullptr, wg_wide = nullptr, wg_reduce = nullptr;
};
static SideBranch* side_branch(cudaStream_t s) {
  static thread_local SideBranch sb[64];
  static thread_local int state[64];      // 0: not tried, 1: ready, -1: unavailable
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (state[dev] == 0) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;
    SideBranch& b = sb[dev];
    const bool ok = cudaStreamCreateWithFlags(&b.side, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.head_done, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.qstats_done, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.fc11_done, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.final_done, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.wg_fork, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.wg_thin, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.wg_wide, cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&b.wg_reduce, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) cudaGetLastError();
    state[dev] = ok ? 1 : -1;
  }
  return state[dev] == 1 ? &sb[dev] : nullptr;
}

static uint64_t* step_keys(const Plan& p, const mvae_state& st) { return reinterpret_cast<uint64_t*>(st.work + p.w.keys); }

static DropSpec make_drop(const Plan& p, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in) {
  DropSpec d;
  memset(&d, 0, sizeof(d));
  d.D = p.D;
  d.rows = p.B;
  if (!in.training || hp.x_drop <= 0.f) {
    d.mode = 0;
    return d;
  }
  d.scale = 1.0f / (1.0f - hp.x_drop);
  d.keep_arm_stride = (int64_t)p.B * p.D;
  if (in.keep_x) {
    d.mode = 1;
    d.keep = in.keep_x;
  } else {
    d.mode = 2;
    d.keys = step_keys(p, st) + kStreamDrop * MVAE_MAX_ARMS;   // written by step_prep_kernel at the top of the forward
    double t = (double)hp.x_drop * 256.0;
    d.thresh16 = (uint32_t)(t + 0.5);
  }
  return d;
}

static BnOff bn_off(const Plan& p) {
  BnOff o;
  for (int i = 0; i < 6; ++i) o.off[i] = p.L.bn_offset[i];
  return o;
}

#define RC(x)            \
  do {                   \
    int _rc = (x);       \
    if (_rc) return _rc; \
  } while (0)

static bool use_tc(const Plan& p, const mvae_hparams& hp) {
  return hp.precision != 3 && gemm_tc_supported(p.B, p.D, p.H);
}

// ---------------------------------------------------------------------------------------------
static CouplingArgs coupling_args(const Plan& p, const mvae_hparams& hp, const mvae_state& st, const float* qc_all,
                                  const float* csmp_all) {
  CouplingArgs c;
  memset(&c, 0, sizeof(c));
  c.A = p.A; c.At = p.At; c.arm_off = p.d.arm_offset; c.B = p.B; c.C = p.C;
  c.qc_all = qc_all; c.csmp_all = csmp_all; c.acc = reinterpret_cast<double*>(st.work + p.w.acc_loss);
  c.gdiff = st.work + p.w.rsum; c.wcat = st.work + p.w.wcat; c.eps = hp.eps; c.lam = hp.lam;
  return c;
}

static int forward_impl(const Plan& p, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                        const mvae_outputs& out, int bump_adam, cudaStream_t s, SideBranch* fork = nullptr,
                        bool clear_all = false) {
  const int A = p.A, B = p.B, D = p.D, H = p.H, Ld = p.Ld, C = p.C, S = p.S;
  const Work& w = p.w;
  float* work = st.work;
  double* acc_fwd = reinterpret_cast<double*>(work + w.acc_fwd);
  const int training = in.training;
  MVAE_CHECK_ARG(in.x != nullptr, "x is required");   // U / E may be null: drawn in-kernel from (seed, step)
  MVAE_CHECK_ARG(!training || B >= 2, "training needs at least 2 cells (batch statistics)");
  MVAE_CHECK_ARG(!(training && hp.s_drop > 0.f) || in.keep_s != nullptr, "keep_s is required when s_drop > 0");
  MVAE_CHECK_ARG(in.x_row_stride >= D, "x_row_stride < D");
  // the accumulator blocks of the forward, the loss and the backward are adjacent in the work buffer: the fused step
  // clears all three with one node
  MVAE_CUDA(cudaMemsetAsync(acc_fwd, 0, (size_t)(clear_all ? w.acc_bwd + w.acc_bwd_floats - w.acc_fwd : w.acc_fwd_floats) * 4, s));
  tl_pdl = 0;            // (a memset node: the step's first kernel is fully ordered behind everything before it)
  RC(launch_step_prep(in.seed, in.step, in.counters, bump_adam, p.d.arm_offset, step_keys(p, st), s));
  float* bn_mean = work + w.bn_mean;
  float* bn_rstd = work + w.bn_rstd;
  if (!training)
    RC(launch_bn_eval_prep(st.bn_running, p.L.bn_stride, bn_off(p), bn_mean, bn_rstd, A, H, Ld, hp.eps, s));

  // ---- fc1: [B,D] x [D,H]  (nn_model.py:264)
  DropSpec drop = make_drop(p, hp, st, in);
  Fc1EpiArgs epi;
  memset(&epi, 0, sizeof(epi));
  timing_begin(TG_FC1_FWD, s);
  const bool ts_path = use_tc(p, hp);
  // the single-tile encoder chain takes fc1's partial tiles and forms a1 itself (no separate fix-up launch)
  Fc1Deferred fc1d;
  memset(&fc1d, 0, sizeof(fc1d));
#ifdef MVAE_NO_FC1_FUSE      // (compile-time switch for A/B measurements)
  const bool defer_fix = false;
#else
  const bool defer_fix = ts_path && training && enc_chain_fwd_is_single(A, B, H, Ld);
#endif
  if (ts_path) {
    RC(ts_fc1_forward(p.d, hp, st, in, drop, w, work + w.a[0], acc_fwd + acc_bn(0, A, 0), defer_fix ? &fc1d : nullptr, s));
  } else {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = in.x; g.sAm = in.x_row_stride; g.sAk = 1; g.A_batch = in.x_arm_stride;
    g.Bm = st.params + p.L.offset[FC1_W]; g.sBk = 1; g.sBn = D; g.B_batch = p.L.arm_stride;
    g.C = work + w.fc1_part; g.sCm = H; g.sCn = 1; g.C_batch = (int64_t)B * H;
    g.M = B; g.N = H; g.K = D;
    g.drop = drop; g.drop_operand = drop.mode ? 1 : 0;
    RC(launch_sgemm_simt(g, A, s));
    epi.part = work + w.fc1_part; epi.split_stride = 0; epi.arm_stride = (int64_t)B * H; epi.ld = H; epi.nsplit = 1;
  }
  epi.params = st.params; epi.p_arm_stride = p.L.arm_stride; epi.offB = p.L.offset[FC1_B];
  epi.out = work + w.a[0]; epi.stats_out = acc_fwd + acc_bn(0, A, 0); epi.B = B; epi.H = H;
  if (!ts_path) RC(launch_fc1_epilogue(epi, A, s));
  timing_end(TG_FC1_FWD, s);
  timing_begin(TG_NARROW_FWD, s);

  // ---- fc2..fc5 with the BatchNorm of the previous layer folded into the load (:265-268)
  int chain_rc = 1;
  if (training && hp.precision != 3) {
    float* aout[4] = {work + w.a[1], work + w.a[2], work + w.a[3], work + w.a[4]};
    chain_rc = launch_enc_chain_fwd(st.params, p.L.arm_stride, p.L.offset, A, B, H, Ld, work + w.a[0], aout, acc_fwd, bn_mean,
                                    bn_rstd, hp.eps, &fc1d, s);
    if (chain_rc < 0 || chain_rc > 1) return chain_rc;
  }
  for (int l = 1; l <= 4 && chain_rc == 1; ++l) {
    DenseFwdArgs a;
    memset(&a, 0, sizeof(a));
    const int nout = l < 4 ? H : Ld;
    a.in = work + w.a[l - 1]; a.in_arm_stride = (int64_t)B * H;
    a.out = work + w.a[l]; a.out_arm_stride = (int64_t)B * nout;
    a.params = st.params; a.p_arm_stride = p.L.arm_stride;
    a.offW = p.L.offset[FC1_W + 2 * l]; a.offB = p.L.offset[FC1_B + 2 * l];
    a.B = B; a.nin = H; a.nout = nout;
    a.bn_mode = training ? 1 : 2;
    a.bn_sums_in = acc_fwd + acc_bn(l - 1, A, 0);
    a.bn_mean = bn_mean + (int64_t)(l - 1) * A * 128;
    a.bn_rstd = bn_rstd + (int64_t)(l - 1) * A * 128;
    a.stats_out = training ? acc_fwd + acc_bn(l, A, 0) : nullptr;
    a.eps = hp.eps; a.relu = 1;
    RC(hp.precision == 3 ? launch_dense_fwd(a, A, s) : launch_dense_fwd_mma(a, A, s));
  }

  // ---- categorical head, Gumbel-softmax, state head, fc6 (:269, :337-351, :278-280)
  HeadArgs h;
  memset(&h, 0, sizeof(h));
  h.A = A; h.At = p.At; h.arm_off = p.d.arm_offset; h.B = B; h.H = H; h.L = Ld; h.C = C; h.S = S;
  h.params = st.params; h.p_arm_stride = p.L.arm_stride;
  h.oWc = p.L.offset[FCC_W]; h.oBc = p.L.offset[FCC_B]; h.oWmu = p.L.offset[FCMU_W]; h.oBmu = p.L.offset[FCMU_B];
  h.oWsig = p.L.offset[FCSIG_W]; h.oBsig = p.L.offset[FCSIG_B]; h.oW6 = p.L.offset[FC6_W]; h.oB6 = p.L.offset[FC6_B];
  h.a5 = work + w.a[4];
  h.bn_mode = training ? 1 : 2;
  h.bn_sums5 = acc_fwd + acc_bn(4, A, 0);
  h.bn_mean5 = bn_mean + (int64_t)4 * A * 128; h.bn_rstd5 = bn_rstd + (int64_t)4 * A * 128;
  h.U = in.U; h.E = in.E; h.keep_s = (training && hp.s_drop > 0.f) ? in.keep_s : nullptr;
  h.cat_mask = in.cat_mask;
  h.ukeys = step_keys(p, st) + kStreamU * MVAE_MAX_ARMS; h.ekeys = step_keys(p, st) + kStreamE * MVAE_MAX_ARMS;
  h.tau = hp.tau; h.temp = hp.temp; h.eps = hp.eps; h.s_scale = 1.0f / (1.0f - hp.s_drop);
  h.hard = hp.hard; h.training = training;
  h.x_low = out.x_low; h.c_prob = out.c_prob; h.qc = out.qc; h.c_smp = out.c_smp;
  h.s_mean = out.s_mean; h.s_logvar = out.s_logvar; h.s_smp = out.s_smp;
  h.ysoft = work + w.ysoft; h.svar = work + w.svar; h.yy = work + w.yy; h.zc = work + w.zc; h.d6 = work + w.d[0];
  h.kl_sums = acc_fwd + acc_kl(A, 0);
  if (training) {   // running statistics of batch_l1..l5 updated by the head kernel (one launch less)
    h.bn_running = st.bn_running; h.bn_stride = p.L.bn_stride; h.bn_off = bn_off(p); h.nbt = st.bn_batches;
    h.bn_sums_all = acc_fwd; h.momentum = hp.momentum;
  }
  RC(launch_head_fwd(h, s));
  if (fork) {
    // q is final: its column statistics (inv_var of the coupling term) start now, beside the decoder chain; the loss
    // accumulators they add to are cleared first
    if (!clear_all) {
      MVAE_CUDA(cudaMemsetAsync(work + w.acc_loss, 0, (size_t)w.acc_loss_floats * 4, s));
      tl_pdl = 0;
    }
    MVAE_CUDA(cudaEventRecord(fork->head_done, s));
    MVAE_CUDA(cudaStreamWaitEvent(fork->side, fork->head_done, 0));
    const int main_pdl = tl_pdl;       // the side branch's launch is an ordinary one; the main branch goes on behind head_fwd
    tl_pdl = 0;
    RC(launch_qstats(coupling_args(p, hp, st, out.qc, out.c_smp), fork->side));
    MVAE_CUDA(cudaEventRecord(fork->qstats_done, fork->side));
    tl_pdl = main_pdl;
  }

  // ---- fc7..fc10 (:281-284)
  if (hp.precision != 3) {
    float* hout[4] = {work + w.d[1], work + w.d[2], work + w.d[3], work + w.d[4]};
    RC(launch_dec_chain_fwd(st.params, p.L.arm_stride, p.L.offset, A, B, H, Ld, work + w.d[0], hout, hp.precision == 1, s));
  } else {
    for (int l = 1; l <= 4; ++l) {
      DenseFwdArgs a;
      memset(&a, 0, sizeof(a));
      const int nin = l == 1 ? Ld : H;
      a.in = work + w.d[l - 1]; a.in_arm_stride = (int64_t)B * nin;
      a.out = work + w.d[l]; a.out_arm_stride = (int64_t)B * H;
      a.params = st.params; a.p_arm_stride = p.L.arm_stride;
      a.offW = p.L.offset[FC7_W + 2 * (l - 1)]; a.offB = p.L.offset[FC7_B + 2 * (l - 1)];
      a.B = B; a.nin = nin; a.nout = H; a.bn_mode = 0; a.eps = hp.eps; a.relu = 1;
      RC(launch_dense_fwd(a, A, s));
    }
  }
  timing_end(TG_NARROW_FWD, s);

  // ---- optional materialised reconstruction x_rec = relu(fc11(h10)) (:287)
  if (out.x_rec && use_tc(p, hp) && hp.precision != 1) {
    RC(ts_fc11_rows(p.d, st, in, w, 0.f, 0, out.x_rec, nullptr, s));
  } else if (out.x_rec) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = work + w.d[4]; g.sAm = H; g.sAk = 1; g.A_batch = (int64_t)B * H;
    g.Bm = st.params + p.L.offset[FC11_W]; g.sBk = 1; g.sBn = H; g.B_batch = p.L.arm_stride;
    g.C = out.x_rec; g.sCm = D; g.sCn = 1; g.C_batch = (int64_t)B * D;
    g.M = B; g.N = D; g.K = H;
    RC(launch_sgemm_simt(g, A, s));
    ReconElemArgs r;
    memset(&r, 0, sizeof(r));
    r.pre = out.x_rec; r.x = in.x; r.x_arm_stride = in.x_arm_stride; r.x_row_stride = in.x_row_stride;
    r.params = st.params; r.p_arm_stride = p.L.arm_stride; r.offB = p.L.offset[FC11_B];
    r.x_rec = out.x_rec; r.recon_acc = nullptr; r.B = B; r.D = D; r.gscale = 0.f; r.want_grad = 0;
    RC(launch_recon_elem(r, A, s));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
static int loss_impl(const Plan& p, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                     const mvae_outputs& out, const float* qc_all, const float* csmp_all, float* loss_out,
                     int want_grad, cudaStream_t s, SideBranch* fork = nullptr, bool cleared = false) {
  const int A = p.A, At = p.At, B = p.B, D = p.D, H = p.H, C = p.C, S = p.S;
  const Work& w = p.w;
  float* work = st.work;
  MVAE_CHECK_ARG(B >= 2, "the loss needs at least 2 cells (unbiased batch variance in inv_var, nn_model.py:75)");
  MVAE_CHECK_ARG(At >= 2, "the coupled loss needs at least 2 arms (the reference divides by the number of arm pairs)");
  MVAE_CHECK_ARG(qc_all && csmp_all && loss_out, "null argument");
  double* acc_loss = reinterpret_cast<double*>(work + w.acc_loss);
  double* acc_fwd = reinterpret_cast<double*>(work + w.acc_fwd);
  const float gscale = (float)(At - 1 > 1 ? At - 1 : 1) / (float)B;

  {
    timing_begin(TG_COUPLING, s);
    // ---- coupling terms over every arm of the model (:558-569): first, while q / c_smp (just written by the head kernel
    // or gathered) are in L2 -- the fc11 passes below stream the gene matrix through it
    const CouplingArgs c = coupling_args(p, hp, st, qc_all, csmp_all);
    if (fork) {          // cleared and summed on the side branch since the head kernel (forward_impl)
      MVAE_CUDA(cudaStreamWaitEvent(s, fork->qstats_done, 0));
    } else {
      if (!cleared) MVAE_CUDA(cudaMemsetAsync(acc_loss, 0, (size_t)w.acc_loss_floats * 4, s));
      RC(launch_qstats(c, s));
    }
    if (!fork) tl_pdl = 0;          // (memset / ordinary launches in front; behind the join the attribute stays)
    RC(launch_coupling_rows(c, s));
    timing_end(TG_COUPLING, s);
  }

  // ---- reconstruction term: fc11 GEMM fused with loss (+ its own backward)
  timing_begin(TG_FC11, s);
  if (use_tc(p, hp)) {
    RC(tc_fc11_loss_grad(p.d, hp, st, in, w, gscale, want_grad, s, fork != nullptr));
  } else {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = work + w.d[4]; g.sAm = H; g.sAk = 1; g.A_batch = (int64_t)B * H;
    g.Bm = st.params + p.L.offset[FC11_W]; g.sBk = 1; g.sBn = H; g.B_batch = p.L.arm_stride;
    g.C = work + w.big; g.sCm = D; g.sCn = 1; g.C_batch = (int64_t)B * D;
    g.M = B; g.N = D; g.K = H;
    RC(launch_sgemm_simt(g, A, s));
    ReconElemArgs r;
    memset(&r, 0, sizeof(r));
    r.pre = work + w.big; r.x = in.x; r.x_arm_stride = in.x_arm_stride; r.x_row_stride = in.x_row_stride;
    r.params = st.params; r.p_arm_stride = p.L.arm_stride; r.offB = p.L.offset[FC11_B];
    r.x_rec = nullptr; r.recon_acc = acc_loss; r.B = B; r.D = D; r.gscale = gscale; r.want_grad = want_grad;
    RC(launch_recon_elem(r, A, s));
    if (want_grad) {
      // d h10 = dY * W11
      memset(&g, 0, sizeof(g));
      g.A = work + w.big; g.sAm = D; g.sAk = 1; g.A_batch = (int64_t)B * D;
      g.Bm = st.params + p.L.offset[FC11_W]; g.sBk = H; g.sBn = 1; g.B_batch = p.L.arm_stride;
      g.C = work + w.g_d10; g.sCm = H; g.sCn = 1; g.C_batch = (int64_t)B * H;
      g.M = B; g.N = H; g.K = D;
      RC(launch_sgemm_simt(g, A, s));
      // d W11 = dY^T * h10
      memset(&g, 0, sizeof(g));
      g.A = work + w.big; g.sAm = 1; g.sAk = D; g.A_batch = (int64_t)B * D;
      g.Bm = work + w.d[4]; g.sBk = H; g.sBn = 1; g.B_batch = (int64_t)B * H;
      g.C = st.grads + p.L.offset[FC11_W]; g.sCm = H; g.sCn = 1; g.C_batch = p.L.arm_stride;
      g.M = D; g.N = H; g.K = B;
      RC(launch_sgemm_simt(g, A, s));
      RC(launch_colsum(work + w.big, (int64_t)B * D, st.grads + p.L.offset[FC11_B], p.L.arm_stride, B, D, A, s));
    }
  }

  timing_end(TG_FC11, s);
  timing_begin(TG_COUPLING, s);
  LossFinalArgs f;
  memset(&f, 0, sizeof(f));
  f.A = A; f.At = At; f.arm_off = p.d.arm_offset; f.B = B; f.D = D; f.C = C; f.S = S;
  f.acc_loss = acc_loss; f.kl_sums = acc_fwd + acc_kl(A, 0);
  f.colc = work + w.colc; f.loss_out = loss_out; f.eps = hp.eps; f.lam = hp.lam; f.beta = hp.beta;
  if (fork) {
    // the side branch turns the sums into the loss vector and the coupling constants while the main branch goes on with
    // the decoder's backward chain; join before the head kernel, which needs the constants
    MVAE_CUDA(cudaEventRecord(fork->fc11_done, s));
    MVAE_CUDA(cudaStreamWaitEvent(fork->side, fork->fc11_done, 0));
    RC(launch_loss_finalize(f, fork->side));        // (an ordinary launch; tl_pdl still describes the main branch)
    {
      const int rc = ts_fc11_gene_fixup(fork->side);   // d fc11.weight / d fc11.bias: nothing before Adam reads them
      if (rc != 0 && rc != 1) return rc;
    }
    MVAE_CUDA(cudaEventRecord(fork->final_done, fork->side));
  } else {
    RC(launch_loss_finalize(f, s));
    tl_pdl = 0;
  }
  timing_end(TG_COUPLING, s);
  return 0;
}

// ---------------------------------------------------------------------------------------------
static int backward_impl(const Plan& p, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                         const mvae_outputs& out, const float* grad_scale, cudaStream_t s, SideBranch* fork = nullptr,
                         bool cleared = false) {
  const int A = p.A, At = p.At, B = p.B, D = p.D, H = p.H, Ld = p.Ld, C = p.C, S = p.S;
  const Work& w = p.w;
  float* work = st.work;
  MVAE_CHECK_ARG(in.training, "backward needs a training-mode forward");
  double* acc_bwd = reinterpret_cast<double*>(work + w.acc_bwd);
  if (!cleared) {
    MVAE_CUDA(cudaMemsetAsync(acc_bwd, 0, (size_t)w.acc_bwd_floats * 4, s));
    tl_pdl = 0;
  }
  float* bn_mean = work + w.bn_mean;
  float* bn_rstd = work + w.bn_rstd;
  const bool tc = use_tc(p, hp);

  // ---- decoder fc10..fc7
  timing_begin(TG_NARROW_BWD, s);
  const float* g_cur = work + w.g_d10;
  if (hp.precision != 3) {
    const float* act[4] = {work + w.d[1], work + w.d[2], work + w.d[3], work + w.d[4]};
    float* dl[4] = {work + w.delta_dec[1], work + w.delta_dec[2], work + w.delta_dec[3], work + w.delta_dec[4]};
    RC(launch_dec_chain_bwd(st.params, p.L.arm_stride, p.L.offset, A, B, H, Ld, work + w.g_d10, act, dl, work + w.gtmp[1],
                            hp.precision == 1, s));
    g_cur = work + w.gtmp[1];
  } else {
    for (int l = 4; l >= 1; --l) {
      DenseBwdArgs a;
      memset(&a, 0, sizeof(a));
      const int nin = l == 1 ? Ld : H;
      a.g_out = g_cur; a.act_out = work + w.d[l]; a.delta = work + w.delta_dec[l];
      a.g_in = work + w.gtmp[l & 1];
      a.params = st.params; a.p_arm_stride = p.L.arm_stride; a.offW = p.L.offset[FC7_W + 2 * (l - 1)];
      a.B = B; a.nin = nin; a.nout = H;
      RC(launch_dense_bwd(a, A, s));
      g_cur = a.g_in;
    }
  }
  // g_cur = d loss / d d6, [A][B][L] in gtmp[1]

  // ---- heads
  HeadArgs h;
  memset(&h, 0, sizeof(h));
  h.A = A; h.At = At; h.arm_off = p.d.arm_offset; h.B = B; h.H = H; h.L = Ld; h.C = C; h.S = S;
  h.params = st.params; h.p_arm_stride = p.L.arm_stride;
  h.oWc = p.L.offset[FCC_W]; h.oBc = p.L.offset[FCC_B]; h.oWmu = p.L.offset[FCMU_W]; h.oBmu = p.L.offset[FCMU_B];
  h.oWsig = p.L.offset[FCSIG_W]; h.oBsig = p.L.offset[FCSIG_B]; h.oW6 = p.L.offset[FC6_W]; h.oB6 = p.L.offset[FC6_B];
  h.E = in.E; h.keep_s = hp.s_drop > 0.f ? in.keep_s : nullptr;
  h.ukeys = step_keys(p, st) + kStreamU * MVAE_MAX_ARMS; h.ekeys = step_keys(p, st) + kStreamE * MVAE_MAX_ARMS;   // same step: the forward's keys
  h.tau = hp.tau; h.temp = hp.temp; h.eps = hp.eps; h.s_scale = 1.0f / (1.0f - hp.s_drop);
  h.hard = hp.hard; h.training = 1;
  h.x_low = out.x_low; h.c_prob = out.c_prob; h.qc = out.qc; h.c_smp = out.c_smp;
  h.s_mean = out.s_mean; h.s_logvar = out.s_logvar; h.s_smp = out.s_smp;
  h.ysoft = work + w.ysoft; h.svar = work + w.svar; h.yy = work + w.yy; h.zc = work + w.zc; h.d6 = work + w.d[0];
  h.g_d6 = g_cur; h.gdiff = work + w.rsum; h.colc = work + w.colc;
  const float scale = (float)(At - 1 > 1 ? At - 1 : 1);
  h.kl_coef = scale * hp.beta / (float)B;
  h.ent_coef = (float)(At - 1) / (float)B;
  h.g_coef = 2.f * hp.lam / (float)B;
  h.delta6 = work + w.delta_dec[0]; h.delta_mu = work + w.delta_mu; h.delta_sig = work + w.delta_sig;
  h.delta_z = work + w.delta_z; h.g_xlow = work + w.g_xlow;
  h.bnb_sums5 = acc_bwd + accb_bn(4, A, 0);
  if (fork) {
    MVAE_CUDA(cudaStreamWaitEvent(s, fork->final_done, 0));      // join: coupling constants, the loss vector, d fc11.*
  }
  RC(launch_head_bwd(h, s));

  // ---- encoder fc5..fc2, then the BatchNorm+ReLU backward of layer 1
  g_cur = work + w.g_xlow;
  int bchain_rc = 1;
  if (hp.precision != 3) {
    const float* act[5] = {work + w.a[0], work + w.a[1], work + w.a[2], work + w.a[3], work + w.a[4]};
    float* dl[5] = {work + w.delta_enc[0], work + w.delta_enc[1], work + w.delta_enc[2], work + w.delta_enc[3],
                    work + w.delta_enc[4]};
    bchain_rc = launch_enc_chain_bwd(st.params, p.L.arm_stride, p.L.offset, A, B, H, Ld, work + w.g_xlow, act, dl, acc_bwd,
                                     bn_mean, bn_rstd, work + w.gtmp[0], hp.precision == 1, s);
    if (bchain_rc < 0 || bchain_rc > 1) return bchain_rc;
  }
  for (int l = 4; l >= 0 && bchain_rc == 1; --l) {
    DenseBwdArgs a;
    memset(&a, 0, sizeof(a));
    const int nout = l == 4 ? Ld : H;
    a.g_out = g_cur; a.act_out = work + w.a[l]; a.delta = work + w.delta_enc[l];
    a.g_in = l > 0 ? work + w.gtmp[l & 1] : nullptr;
    a.params = st.params; a.p_arm_stride = p.L.arm_stride; a.offW = p.L.offset[FC1_W + 2 * l];
    a.B = B; a.nin = H; a.nout = nout;
    a.bn_out = 1; a.bnb_sums = acc_bwd + accb_bn(l, A, 0);
    a.mean_out = bn_mean + (int64_t)l * A * 128; a.rstd_out = bn_rstd + (int64_t)l * A * 128;
    if (l > 0) {
      a.bn_in = 1; a.act_in = work + w.a[l - 1];
      a.mean_in = bn_mean + (int64_t)(l - 1) * A * 128; a.rstd_in = bn_rstd + (int64_t)(l - 1) * A * 128;
      a.bnb_sums_next = acc_bwd + accb_bn(l - 1, A, 0);
    } else if (tc) {
      a.delta_t = work + w.delta1_t; a.delta_t_ld = w.Bpad; a.delta_t_arm_stride = (int64_t)w.Hpad * w.Bpad;
    }
    RC(hp.precision == 3 ? launch_dense_bwd(a, A, s) : launch_dense_bwd_mma(a, A, hp.precision == 1, s));
    g_cur = a.g_in;
  }

  timing_end(TG_NARROW_BWD, s);
  // Order: the narrow weight gradients run FIRST, while the activations and deltas that narrow_bwd just touched are
  // still in L2 -- the fc1 weight gradient streams the whole gene matrix through L2 and would evict them.
  // ---- weight gradients of every narrow layer
  timing_begin(TG_WGRAD, s);
  WgArgs wg;
  memset(&wg, 0, sizeof(wg));
  int np = 0;
  auto add = [&](int64_t doff, int nout, int64_t ioff, int nin, int in_ld, int bn_layer, int pw, int pb) {
    WgProblem& q = wg.prob[np++];
    q.delta_off = doff; q.delta_arm_stride = (int64_t)B * nout; q.nout = nout;
    q.in_off = ioff; q.in_arm_stride = (int64_t)B * in_ld; q.nin = nin; q.in_ld = in_ld;
    q.bn_layer = bn_layer; q.poffW = p.L.offset[pw]; q.poffB = p.L.offset[pb];
  };
  add(w.delta_enc[0], H, 0, 0, 1, -1, FC1_W, FC1_B);  // bias only
  for (int l = 1; l <= 4; ++l) add(w.delta_enc[l], l == 4 ? Ld : H, w.a[l - 1], H, H, l - 1, FC1_W + 2 * l, FC1_B + 2 * l);
  add(w.delta_z, C, w.yy, Ld, Ld + C, -1, FCC_W, FCC_B);
  add(w.delta_mu, S, w.yy, Ld + C, Ld + C, -1, FCMU_W, FCMU_B);
  add(w.delta_sig, S, w.yy, Ld + C, Ld + C, -1, FCSIG_W, FCSIG_B);
  add(w.delta_dec[0], Ld, w.zc, C + S, C + S, -1, FC6_W, FC6_B);
  add(w.delta_dec[1], H, w.d[0], Ld, Ld, -1, FC7_W, FC7_B);
  for (int l = 2; l <= 4; ++l) add(w.delta_dec[l], H, w.d[l - 1], H, H, -1, FC7_W + 2 * (l - 1), FC7_B + 2 * (l - 1));
  wg.nprob = np; wg.A = A; wg.B = B; wg.rows_per_split = w.wg_rows; wg.nsplit = w.wg_nsplit;
  wg.work = work; wg.bn_mean = bn_mean; wg.bn_rstd = bn_rstd;
  wg.part = work + w.wg_part; wg.part_arm_stride = w.wg_floats; wg.part_split_stride = (int64_t)A * w.wg_floats;
  wg.base_off = p.L.offset[FC1_B];
  wg.grads = st.grads; wg.g_arm_stride = p.L.arm_stride;
  WgFork wf;
  memset(&wf, 0, sizeof(wf));
  const bool wg_forked = fork != nullptr && hp.precision != 3;
  if (wg_forked) {
    wf.side = fork->side; wf.fork_ev = fork->wg_fork; wf.thin_done = fork->wg_thin; wf.wide_done = fork->wg_wide;
    wf.reduce_done = fork->wg_reduce;
  }
  RC(hp.precision == 3 ? launch_wgrad(wg, s) : launch_wgrad_mma(wg, hp.precision == 1, s, wg_forked ? &wf : nullptr));
  timing_end(TG_WGRAD, s);
  // ---- d fc1.weight = delta1^T * dropout(x)
  DropSpec drop = make_drop(p, hp, st, in);
  timing_begin(TG_FC1_WGRAD, s);
  if (tc && hp.precision != 1) {
    RC(ts_fc1_wgrad(p.d, st, in, drop, w, s));
  } else if (tc) {
    RC(tc_fc1_wgrad(p.d, hp, st, in, drop, w, s));
  } else {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = work + w.delta_enc[0]; g.sAm = 1; g.sAk = H; g.A_batch = (int64_t)B * H;
    g.Bm = in.x; g.sBk = in.x_row_stride; g.sBn = 1; g.B_batch = in.x_arm_stride;
    g.C = st.grads + p.L.offset[FC1_W]; g.sCm = D; g.sCn = 1; g.C_batch = p.L.arm_stride;
    g.M = H; g.N = D; g.K = B;
    g.drop = drop; g.drop_operand = drop.mode ? 2 : 0;
    RC(launch_sgemm_simt(g, A, s));
  }

  timing_end(TG_FC1_WGRAD, s);
  if (wg_forked) MVAE_CUDA(cudaStreamWaitEvent(s, fork->wg_reduce, 0));      // join: the narrow layers' gradients are final

  if (grad_scale) RC(launch_scale(st.grads, (int64_t)A * p.L.arm_stride, grad_scale, s));
  return 0;
}

}  // namespace mvae

using namespace mvae;

extern "C" {

int mvae_pdl_enable(int on) {
  const int prev = g_pdl_on;
  g_pdl_on = on ? 1 : 0;
  return prev;
}

int mvae_forward(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st, const mvae_inputs* in,
                 const mvae_outputs* out, void* stream) {
  Plan p;
  RC(make_plan(dims, &p));
  MVAE_CHECK_ARG(hp && st && in && out, "null argument");
  RC(check_device());
  return forward_impl(p, *hp, *st, *in, *out, 0, (cudaStream_t)stream);
}

int mvae_loss(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st, const mvae_inputs* in,
              const mvae_outputs* out, const float* qc_all, const float* c_smp_all, float* loss_out, int want_grad,
              void* stream) {
  Plan p;
  RC(make_plan(dims, &p));
  MVAE_CHECK_ARG(hp && st && in && out, "null argument");
  RC(check_device());
  return loss_impl(p, *hp, *st, *in, *out, qc_all, c_smp_all, loss_out, want_grad, (cudaStream_t)stream);
}

int mvae_backward(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st, const mvae_inputs* in,
                  const mvae_outputs* out, const float* grad_scale, void* stream) {
  Plan p;
  RC(make_plan(dims, &p));
  MVAE_CHECK_ARG(hp && st && in && out, "null argument");
  RC(check_device());
  return backward_impl(p, *hp, *st, *in, *out, grad_scale, (cudaStream_t)stream);
}

int mvae_adam(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
              float eps, float weight_decay, int32_t adamw, int64_t step, uint64_t* step_counter, void* stream) {
  MVAE_CHECK_ARG(params && grads && m && v, "null argument");
  RC(check_device());
  if (step_counter) RC(launch_counter_inc(step_counter, (cudaStream_t)stream));
  return launch_adam(params, grads, m, v, n, lr, beta1, beta2, eps, weight_decay, adamw, step, step_counter,
                     (cudaStream_t)stream);
}

int mvae_adam_peer(float* const* peer_params, const float* const* peer_grads, float* m, float* v, int64_t n, int32_t rank,
                   int32_t world, float lr, float beta1, float beta2, float eps, int64_t step, uint64_t* step_counter,
                   void* stream) {
  MVAE_CHECK_ARG(peer_params && peer_grads && m && v, "null argument");
  RC(check_device());
  if (step_counter) RC(launch_counter_inc(step_counter, (cudaStream_t)stream));
  return launch_adam_peer(peer_params, peer_grads, m, v, n, rank, world, lr, beta1, beta2, eps, step, step_counter,
                          (cudaStream_t)stream);
}

int mvae_train_step(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st, const mvae_inputs* in,
                    const mvae_outputs* out, float* loss_out, float lr, float beta1, float beta2, float adam_eps,
                    int64_t step, void* stream) {
  Plan p;
  RC(make_plan(dims, &p));
  MVAE_CHECK_ARG(hp && st && in && out && loss_out, "null argument");
  MVAE_CHECK_ARG(dims->n_arm == dims->n_arm_total, "mvae_train_step needs every arm local; with sharded arms call forward/loss/backward around the all-gather");
  RC(check_device());
  cudaStream_t s = (cudaStream_t)stream;
  mvae_outputs o = *out;
  o.x_rec = nullptr;
  SideBranch* fork = timing_enabled() ? nullptr : side_branch(s);     // (per-group timing measures the serial order)
  RC(forward_impl(p, *hp, *st, *in, o, 1, s, fork, true));
  RC(loss_impl(p, *hp, *st, *in, o, o.qc, o.c_smp, loss_out, 1, s, fork, true));
  RC(backward_impl(p, *hp, *st, *in, o, nullptr, s, fork, true));
  TimedScope ts(TG_ADAM, s);
  return launch_adam(st->params, st->grads, st->adam_m, st->adam_v, (int64_t)p.A * p.L.arm_stride, lr, beta1, beta2,
                     adam_eps, 0.f, 0, step, in->counters ? in->counters + 1 : nullptr, s);
}

int mvae_grad_step(const mvae_dims* dims, const mvae_hparams* hp, const mvae_state* st, const mvae_inputs* in,
                   const mvae_outputs* out, float* loss_out, void* stream) {
  Plan p;
  RC(make_plan(dims, &p));
  MVAE_CHECK_ARG(hp && st && in && out && loss_out, "null argument");
  MVAE_CHECK_ARG(dims->n_arm == dims->n_arm_total, "mvae_grad_step needs every arm local; with sharded arms call forward/loss/backward around the all-gather");
  RC(check_device());
  cudaStream_t s = (cudaStream_t)stream;
  mvae_outputs o = *out;
  o.x_rec = nullptr;
  SideBranch* fork = timing_enabled() ? nullptr : side_branch(s);
  RC(forward_impl(p, *hp, *st, *in, o, 0, s, fork, true));
  RC(loss_impl(p, *hp, *st, *in, o, o.qc, o.c_smp, loss_out, 1, s, fork, true));
  return backward_impl(p, *hp, *st, *in, o, nullptr, s, fork, true);
}

int mvae_dropout_mask(const mvae_dims* dims, const mvae_hparams* hp, const mvae_inputs* in, uint8_t* keep_out,
                      void* stream) {
  Plan p;
  RC(make_plan(dims, &p));
  MVAE_CHECK_ARG(hp && in && keep_out, "null argument");
  RC(check_device());
  mvae_inputs i2 = *in;
  i2.keep_x = nullptr;
  i2.training = 1;
  mvae_state nost;
  memset(&nost, 0, sizeof(nost));
  DropSpec d = make_drop(p, *hp, nost, i2);
  MVAE_CHECK_ARG(d.mode == 2, "x_drop is 0: there is no mask");
  d.keys = nullptr;    // the keys are derived here, on the host, from the same (seed, step, global arm) function
  for (int a = 0; a < p.A; ++a)
    RC(launch_dropout_mask(d, stream_key(in->seed, in->step, (uint32_t)(a + p.d.arm_offset), kStreamDrop), p.B,
                           keep_out + (int64_t)a * p.B * p.D, (cudaStream_t)stream));
  return 0;
}

int mvae_argmax(const float* q, int32_t* labels, int64_t rows, int32_t cols, void* stream) {
  MVAE_CHECK_ARG(q && labels && cols >= 1, "bad argument");
  RC(check_device());
  return launch_argmax(q, labels, rows, cols, (cudaStream_t)stream);
}

int mvae_confmat(const int32_t* labels, int64_t n_cells, int32_t n_arm, int32_t n_categories, int32_t* counts, void* stream) {
  MVAE_CHECK_ARG(labels && counts && n_arm >= 2 && n_categories >= 1 && n_cells >= 0, "bad argument");
  RC(check_device());
  return launch_confmat(labels, n_cells, n_arm, n_categories, counts, (cudaStream_t)stream);
}

int mvae_unpack_rows(const uint32_t* bitmap, const float* values, const int64_t* row_ptr, int64_t rows, int32_t n_cols,
                     float* out, int64_t out_row_stride, void* stream) {
  MVAE_CHECK_ARG(bitmap && row_ptr && out, "null argument");     // values may be null when the batch has no non-zero
  MVAE_CHECK_ARG(rows >= 0 && n_cols >= 1 && out_row_stride >= n_cols, "bad shape");
  RC(check_device());
  return launch_unpack_rows(bitmap, values, row_ptr, rows, n_cols, out, out_row_stride, (cudaStream_t)stream);
}

// ---- augmenter forward: see include/mixvae_b200.h
int mvae_fold_affine(const float* bias, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                     int32_t n, float* scale, float* shift, void* stream) {
  MVAE_CHECK_ARG(scale && shift && n >= 1, "bad argument");
  MVAE_CHECK_ARG((mean == nullptr) == (var == nullptr), "mean and var come together");
  RC(check_device());
  return launch_fold_affine(bias, mean, var, gamma, beta, eps, n, scale, shift, (cudaStream_t)stream);
}

int mvae_fma_rows(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, float* out, int64_t ldo,
                  int64_t rows, int32_t n, float a_scale, void* stream) {
  MVAE_CHECK_ARG(a && out && rows >= 0 && n >= 1, "bad argument");
  RC(check_device());
  return launch_fma_rows(a, lda, b, ldb, c, ldc, out, ldo, rows, n, a_scale, (cudaStream_t)stream);
}

int mvae_linear_act(const float* x, int64_t x_pitch, const float* w, int64_t w_pitch, float* y, int64_t y_pitch, int64_t rows,
                    int32_t n_out, int32_t k, const float* scale, const float* shift, int32_t act, int32_t split3, void* stream) {
  MVAE_CHECK_ARG(x && w && y && scale && shift, "null argument");
  MVAE_CHECK_ARG(rows >= 1 && rows < (1ll << 31) && n_out >= 1 && k >= 1, "bad shape");
  MVAE_CHECK_ARG(x_pitch % 4 == 0 && w_pitch % 4 == 0 && x_pitch >= k && w_pitch >= k && y_pitch >= n_out,
                 "pitches: x and w multiples of 4 floats and >= k, y >= n_out");
  MVAE_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "x and w must be 16-byte aligned");
  MVAE_CHECK_ARG(act >= 0 && act <= 3, "act must be 0 (none), 1 (relu), 2 (elu) or 3 (sigmoid)");
  RC(check_device());
  return tc_linear_act(x, x_pitch, w, w_pitch, y, y_pitch, rows, n_out, k, scale, shift, act, split3, (cudaStream_t)stream);
}

}  // extern "C"
