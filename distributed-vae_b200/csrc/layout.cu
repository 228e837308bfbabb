// layout.cu — host-side shape bookkeeping: flat parameter layout and workspace map.
#include <stdarg.h>

#include <vector>

#include "common.cuh"

namespace mvae {

static thread_local char g_err[512] = "";
int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- per-group timing ---------------------------------------------------------------------------
struct TimedSpan { cudaEvent_t a, b; int group; };
static bool g_timing_on = false;
static std::vector<TimedSpan> g_spans;
static cudaEvent_t g_open[TG_COUNT];

bool timing_enabled() { return g_timing_on; }
void timing_begin(int group, cudaStream_t s) {
  if (!g_timing_on) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  g_open[group] = e;
}
void timing_end(int group, cudaStream_t s) {
  if (!g_timing_on) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, s);
  g_spans.push_back({g_open[group], e, group});
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static int check_dims(const mvae_dims& d) {
  MVAE_CHECK_ARG(d.n_arm >= 1 && d.n_arm <= MVAE_MAX_ARMS, "n_arm=%d out of range [1,%d]", d.n_arm, MVAE_MAX_ARMS);
  MVAE_CHECK_ARG(d.n_arm_total >= d.n_arm && d.n_arm_total <= MVAE_MAX_ARMS, "n_arm_total=%d invalid", d.n_arm_total);
  MVAE_CHECK_ARG(d.arm_offset >= 0 && d.arm_offset + d.n_arm <= d.n_arm_total, "arm_offset=%d invalid", d.arm_offset);
  MVAE_CHECK_ARG(d.batch >= 1, "batch=%d", d.batch);
  MVAE_CHECK_ARG(d.input_dim >= 1, "input_dim=%d", d.input_dim);
  MVAE_CHECK_ARG(d.fc_dim >= 1 && d.fc_dim <= kMaxH, "fc_dim=%d unsupported (1..%d)", d.fc_dim, kMaxH);
  MVAE_CHECK_ARG(d.lowD_dim >= 1 && d.lowD_dim <= kMaxL, "lowD_dim=%d unsupported (1..%d)", d.lowD_dim, kMaxL);
  MVAE_CHECK_ARG(d.n_categories >= 2 && d.n_categories <= kMaxC, "n_categories=%d unsupported (2..%d)", d.n_categories, kMaxC);
  MVAE_CHECK_ARG(d.state_dim >= 1 && d.state_dim <= kMaxS, "state_dim=%d unsupported (1..%d)", d.state_dim, kMaxS);
  // the narrow weight-gradient kernels carry the bias gradient as one extra input column: inputs + 1 <= 128
  MVAE_CHECK_ARG(d.lowD_dim + d.n_categories <= 127 && d.n_categories + d.state_dim <= 127,
                 "lowD_dim + n_categories and n_categories + state_dim must be <= 127");
  return 0;
}

int compute_layout(const mvae_dims& d, mvae_layout* L) {
  int rc = check_dims(d);
  if (rc) return rc;
  const int64_t D = d.input_dim, H = d.fc_dim, Ld = d.lowD_dim, C = d.n_categories, S = d.state_dim;
  const int64_t shp[14][2] = {{H, D}, {H, H}, {H, H}, {H, H}, {Ld, H}, {C, Ld}, {S, Ld + C}, {S, Ld + C},
                              {Ld, C + S}, {H, Ld}, {H, H}, {H, H}, {H, H}, {D, H}};
  int64_t off = 0;
  for (int l = 0; l < 14; ++l) {
    L->offset[2 * l] = off;
    L->numel[2 * l] = shp[l][0] * shp[l][1];
    off = round_up(off + L->numel[2 * l], 32);
    L->offset[2 * l + 1] = off;
    L->numel[2 * l + 1] = shp[l][0];
    off = round_up(off + L->numel[2 * l + 1], 32);
  }
  L->arm_stride = round_up(off, 256);
  const int64_t bnf[6] = {H, H, H, H, Ld, S};
  off = 0;
  for (int i = 0; i < 6; ++i) {
    L->bn_offset[i] = off;
    off += 2 * bnf[i];
  }
  L->bn_stride = round_up(off, 32);
  Work w = make_work(d);
  L->work_floats = w.total;
  return 0;
}

Work make_work(const mvae_dims& d) {
  Work w;
  memset(&w, 0, sizeof(w));
  const int64_t A = d.n_arm, At = d.n_arm_total, B = d.batch, D = d.input_dim, H = d.fc_dim, Ld = d.lowD_dim,
                C = d.n_categories, S = d.state_dim;
  int64_t off = 0;
  auto take = [&](int64_t n) {
    int64_t o = off;
    off = round_up(off + n, 64);  // 256-byte aligned blocks
    return o;
  };
  w.Bpad = (int32_t)round_up(B, 128);
  w.Dpad = (int32_t)round_up(D, 128);
  w.Hpad = 128;
  for (int i = 0; i < 4; ++i) w.a[i] = take(A * B * H);
  w.a[4] = take(A * B * Ld);
  w.bn_mean = take(5 * A * 128);
  w.bn_rstd = take(5 * A * 128);
  w.ysoft = take(A * B * C);
  w.svar = take(A * B * S);
  w.yy = take(A * B * (Ld + C));
  w.zc = take(A * B * (C + S));
  w.d[0] = take(A * B * Ld);
  for (int i = 1; i < 5; ++i) w.d[i] = take(A * B * H);
  w.g_d10 = take(A * B * H);
  w.gtmp[0] = take(A * B * H);
  w.gtmp[1] = take(A * B * H);
  w.delta_dec[0] = take(A * B * Ld);
  for (int i = 1; i < 5; ++i) w.delta_dec[i] = take(A * B * H);
  w.delta_mu = take(A * B * S);
  w.delta_sig = take(A * B * S);
  w.delta_z = take(A * B * C);
  w.g_xlow = take(A * B * Ld);
  for (int i = 0; i < 4; ++i) w.delta_enc[i] = take(A * B * H);
  w.delta_enc[4] = take(A * B * Ld);
  w.delta1_t = take(A * (int64_t)w.Hpad * w.Bpad);
  w.d10_t = take(A * (int64_t)w.Hpad * w.Bpad);
  w.w11_t = take(A * (int64_t)w.Hpad * w.Dpad);
  w.rsum = take(A * B * C);
  w.colc = take(A * 4 * 128);
  w.wcat = take(At * 128);
  w.fc1_splitk = 8;
  {
    // split-K partials of the tensor-core GEMMs: fc1 / d h10 use [split<=4][A][Bpad][128], d W11 [split<=2][A][Dpad][128]
    const int64_t p1 = (int64_t)w.fc1_splitk * A * w.Bpad * 128, p2 = (int64_t)8 * A * w.Dpad * 128;
    // stream-K partial tiles of the fc1 kernels (ts_gemm.cu): one [128][128] tile per (CTA, output tile) pair
    const int64_t t1 = A * (w.Bpad / 128), t2 = A * (w.Dpad / 128);
    const int64_t p3 = ((t1 > t2 ? t1 : t2) + 2 * A + 320) * 128 * 128;
    int64_t pm = p1 > p2 ? p1 : p2;
    if (p3 > pm) pm = p3;
    // fc11_ts.cu keeps the row pass's partial tiles until the gene pass's fix-up: two regions
    const int64_t p4 = (f11_gene_slot0((int)A, (int)B) + t2 + 160) * 128 * 128;
    if (p4 > pm) pm = p4;
    w.fc1_part = take(pm);
  }
  w.db_part = take((int64_t)8 * A * w.Dpad + 160 * 512);   // also [slot][4 groups][128] of fc11_ts.cu
  w.big = take(A * B * D);
  // narrow-layer weight-gradient partials mirror the parameter range [offset(fc1.b), offset(fc11.w))
  mvae_layout L;
  {
    // duplicate of compute_layout's offsets without recursion
    const int64_t shp[14][2] = {{H, D}, {H, H}, {H, H}, {H, H}, {Ld, H}, {C, Ld}, {S, Ld + C}, {S, Ld + C},
                                {Ld, C + S}, {H, Ld}, {H, H}, {H, H}, {H, H}, {D, H}};
    int64_t o = 0;
    for (int l = 0; l < 14; ++l) {
      L.offset[2 * l] = o;
      o = round_up(o + shp[l][0] * shp[l][1], 32);
      L.offset[2 * l + 1] = o;
      o = round_up(o + shp[l][0], 32);
    }
  }
  w.wg_floats = L.offset[FC11_W] - L.offset[FC1_B];
  w.wg_rows = 256;
  w.wg_nsplit = (int32_t)((B + w.wg_rows - 1) / w.wg_rows);
  // the cp.async kernel (wgrad2) picks its split counts per problem from the SM count, at most kWgMaxSplit
  w.wg_part = take((int64_t)(w.wg_nsplit > kWgMaxSplit ? w.wg_nsplit : kWgMaxSplit) * A * w.wg_floats);
  w.acc_fwd_floats = 2 * acc_fwd_doubles((int)A);
  w.acc_fwd = take(w.acc_fwd_floats);
  w.acc_loss_floats = 2 * acc_loss_doubles();
  w.acc_loss = take(w.acc_loss_floats);
  w.acc_bwd_floats = 2 * acc_bwd_doubles((int)A);
  w.acc_bwd = take(w.acc_bwd_floats);
  w.keys = take(2 * ((int64_t)kNumStreams * MVAE_MAX_ARMS + 2));
  w.total = off;
  return w;
}

}  // namespace mvae

extern "C" {
const char* mvae_last_error(void) { return mvae::g_err; }
int mvae_abi_version(void) { return MVAE_ABI_VERSION; }
int mvae_compute_layout(const mvae_dims* dims, mvae_layout* out) {
  if (!dims || !out) {
    mvae::set_error("null argument");
    return -1;
  }
  return mvae::compute_layout(*dims, out);
}
int64_t mvae_launch_count(void) { return mvae::g_launches; }
int mvae_timing_enable(int on) {
  for (auto& sp : mvae::g_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  mvae::g_spans.clear();
  mvae::g_timing_on = on != 0;
  return 0;
}
int mvae_timing_read(float* ms_out, int32_t* count_out, int32_t n_groups) {
  if (!ms_out || !count_out || n_groups < mvae::TG_COUNT) { mvae::set_error("timing_read: need %d slots", mvae::TG_COUNT); return -1; }
  for (int i = 0; i < n_groups; ++i) { ms_out[i] = 0.f; count_out[i] = 0; }
  for (auto& sp : mvae::g_spans) {
    cudaError_t e = cudaEventSynchronize(sp.b);
    if (e != cudaSuccess) { mvae::set_error("timing_read: %s", cudaGetErrorString(e)); return (int)e; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, sp.a, sp.b);
    ms_out[sp.group] += ms;
    count_out[sp.group] += 1;
  }
  return 0;
}
}
