// fc11_ts.cu — the last decoder layer fused with the reconstruction loss and its own backward, third generation.
//
// Reference ops (mmidas/nn_model.py): x_hat = relu(fc11(h10)) :287; 0.5*mse_sum/B + 0.5*BCE(bin(x_hat), bin(x))
// :542-546; autograd of both (dY = max(A-1,1)/B * (x_hat - x) * [x_hat > 0]; the BCE half acts on constants).
// Two passes over x, nothing [B,D]-sized is ever written:
//
//   ROW  pass (thread = cell):  X[128 cells x 64 genes] = h10 . W11^T ; loss sums ; dY ; d h10 += dY . W11
//   GENE pass (thread = gene):  X^T[128 genes x 64 cells] = W11 . h10^T ; dY^T ; d fc11.weight += dY^T . h10 ; d fc11.bias
//
// Same skeleton for both (roles of h10 and W11 swapped through the tensor maps):
//   * the resident operand R (128 rows x H: h10 block / W11 block) lives in TENSOR MEMORY and feeds MMA1 in the TS form;
//   * a unit is a PAIR of 32-column halves.  Streamed per unit: two raw x tiles (16 KB each, own deep ring: they come
//     from HBM) and ONE image of the small operand tile T (64 rows x H, L2-resident) that serves both MMAs: the
//     SWIZZLE_128B_BASE32B layout is the only one MN-major TF32 operands may use, and the tensor core reads it K-major
//     as well when the descriptor's stride between row groups is that of its 4-row atoms (512 B) -- measured, the
//     second-generation kernel pulled every T tile from L2 twice, once per swizzle;
//   * MMA1 (N = 64: 51 cycles of issue per K = 8 step instead of 2 x 47 at N = 32) -> acc1[pair] (TMEM) -> 4 epilogue
//     groups (half-unit j belongs to group j % 4; lane = row of R) compute x_hat, the loss terms and dY in registers and
//     store dY as the A operand of MMA2 straight into TMEM (tcgen05.st);
//   * MMA2 (TS form, per half) accumulates into acc2 (TMEM) over the whole segment;
//   * stream-K over (tile, unit): one CTA per SM, one wave; partial tiles are summed in a fixed order by a fix-up kernel.
// TMEM columns: acc2 [0,128) | acc1 2 x 64 [128,256) | dY stages 4 x 32 [256,384) | R [384,512).
#include "gemm_tc.h"
#include "tc_common.cuh"

namespace mvae {

namespace {
using namespace tc;

constexpr int UN = 32;                     // columns per half-unit (genes for ROW, cells for GENE)
constexpr int NG = 4;                      // epilogue groups (4 warps each); group g owns the half-units j = g (mod 4)
constexpr int CTRL_WARPS = 4;              // x TMA, small-operand TMA, MMA1 issue, MMA2 issue
constexpr int THREADS = 32 * (CTRL_WARPS + 4 * NG);
constexpr int X_BYTES = 16384;
constexpr int SLAB_BYTES = 8192;           // [64 rows x 128 B] of T: 32 columns of H
constexpr int IMG_BYTES = 4 * SLAB_BYTES;  // the T tile of a unit: 64 rows x 128 columns of H (SWIZZLE_128B_BASE32B)
constexpr uint32_t LT_BASE32B = 1;         // descriptor layout type
#ifndef F11_NW
#define F11_NW 4                           // T-tile ring depth (measured: 4 stages + 4 x slots beat 3 + 8)
#endif
#ifndef F11_NX
#define F11_NX 4                           // x ring depth: a multiple of NG (see launch_f11)
#endif
// compile-time ring depths: every ring slot and mbarrier address is the shared base plus an immediate
constexpr int NXC = F11_NX, NWC = F11_NW;
constexpr int NBS = NXC;                   // bias ring (ROW pass): one slot per x slot
static_assert(2 * NXC + 2 * NWC + 5 * 4 + 4 + 1 <= 64, "the barrier block is 64 words");
static_assert(NXC % 4 == 0, "the x ring depth must be a multiple of the number of epilogue groups");
constexpr int TILE_FLOATS = 128 * 128;
constexpr uint32_t COL_ACC2 = 0, COL_ACC1 = 128, COL_A2 = 256, COL_R = 384;

struct F11Args {
  int B, D, H, HN;
  int batch, rtiles, ktiles;        // arms, 128-row blocks of R per arm, units (pairs of 32-column halves) per tile
  int x_batched;
  int nx, nw;                       // ring depths: x tiles, T tiles
  int want_grad;
  float gscale;
  const float* R; int64_t r_arm_stride; int r_rows;     // resident operand: [arm][r_rows][H]
  const float* bias; int64_t bias_arm_stride;           // fc11.bias
  float* x_rec; int64_t xrec_arm_stride;                // ROW: optional materialised reconstruction
  double* recon_acc;                                    // ROW: loss sums
  float* part;                                          // partial tiles [slot][128][128]
  float* db_part;                                       // GENE: d fc11.bias partials [slot][NG][128]
};

__host__ __device__ inline int64_t cta_of_unit(int64_t u, int64_t U, int64_t G) { return ((u + 1) * G - 1) / U; }

__device__ __forceinline__ uint64_t make_desc_lt(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
#ifdef F11_DEBUG
__device__ int g_f11_dead = 0;
__device__ __forceinline__ void dbg_wait(uint64_t* bar, uint32_t parity, int tag, int idx, int gene) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*(volatile int*)&g_f11_dead) return;
    if (clock64() - t0 > 400000000LL) {
      if ((threadIdx.x & 31) == 0)
        printf("F11 timeout gene %d blk %d warp %d tag %d idx %d parity %u\n", gene, blockIdx.x, threadIdx.x >> 5, tag, idx, parity);
      *(volatile int*)&g_f11_dead = 1;
      return;
    }
  }
}
#define WAIT(bar, par, tag, idx) dbg_wait(bar, par, tag, idx, (int)GENE)
#else
#define WAIT(bar, par, tag, idx) mbar_wait(bar, par)
#endif
#ifdef F11_STAMPS
// development only (-DF11_STAMPS, built into a separate library by profiles/tools/f11_stamps.py): clock64 stamps of the
// protocol events of CTA 0 and CTA 74, first 128 (half-)units, per pass
__device__ long long g_f11_stamps[2][2][10][128];
#define STAMP(ev, idx)                                                                                       \
  do {                                                                                                       \
    if ((blockIdx.x == 0 || blockIdx.x == 74) && (idx) < 128 && (threadIdx.x & 31) == 0)                     \
      g_f11_stamps[GENE ? 1 : 0][blockIdx.x ? 1 : 0][ev][idx] = clock64();                                   \
  } while (0)
#else
#define STAMP(ev, idx) do { } while (0)
#endif
__device__ __forceinline__ void mbar_wait3(uint64_t* a, uint32_t pa, uint64_t* b, uint32_t pb, uint64_t* c, uint32_t pc) {
  const bool ra = mbar_try_wait(a, pa), rb = mbar_try_wait(b, pb), rc = mbar_try_wait(c, pc);
  if (!ra) mbar_wait(a, pa);
  if (!rb) mbar_wait(b, pb);
  if (!rc) mbar_wait(c, pc);
}

// TRAIN = true: the training-step instance (gradients wanted, no materialised reconstruction): the flags are compile-time
// so the epilogue carries no per-element branches for them.
// Index conventions: a CTA owns the units [u0, u1) of the (tile, unit) sequence; inside the CTA, unit i consists of the
// half-units j = 2 i and 2 i + 1; KT = 2 * ktiles is the number of half-units per tile (the last one may lie entirely
// beyond the matrix: TMA fills zeros, the epilogue guards its stores).
template <bool GENE, bool TRAIN>
__global__ void __launch_bounds__(THREADS, 1)
fc11_ts_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmT, const F11Args a) {
  const bool want_grad = TRAIN || a.want_grad;
  float* const x_rec = TRAIN ? nullptr : a.x_rec;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer stays derived from smem_raw, so the compiler keeps the
  // shared address space (LDS / direct mbarrier addresses instead of generic loads and 64-bit window arithmetic)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto xs = [&](int s) { return smem + (size_t)s * X_BYTES; };
  auto ts = [&](int s) { return smem + (size_t)NXC * X_BYTES + (size_t)s * IMG_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NXC * X_BYTES + (size_t)NWC * IMG_BYTES);
  uint64_t* x_full = bars;                 uint64_t* x_empty = x_full + NXC;
  uint64_t* w_full = x_empty + NXC;       uint64_t* w_empty = w_full + NWC;
  uint64_t* acc1_full = w_empty + NWC;    uint64_t* acc1_empty = acc1_full + NG;
  uint64_t* a2_full = acc1_empty + NG;     uint64_t* a2_empty = a2_full + NG;
  uint64_t* r_full = a2_empty + NG;        // R of the current segment is in TMEM (and acc2 of the previous one drained)
  // segment s complete: barrier s % NG.  An epilogue group can be up to NG half-units -- hence several segments when a
  // CTA's share of a tile is a single unit -- ahead of the tensor pipe; one parity bit cannot tell those apart.
  uint64_t* acc2_full = r_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_full + NG);
  // ROW: the 32 bias values of a half-unit travel with its x tile (1-D bulk copy on the same mbarrier) into a slot of their
  // own per x slot; the x slot is handed back after the last use of both
  float* bias_s = reinterpret_cast<float*>(bars + 64);          // [NBS][32], 512 bytes behind the (1 KB-aligned) barrier block

  const int KP = a.ktiles, KT = 2 * KP;
  const int64_t U = (int64_t)a.batch * a.rtiles * KP, G = gridDim.x;
  const int64_t u0 = (int64_t)blockIdx.x * U / G, u1 = ((int64_t)blockIdx.x + 1) * U / G;
  const int np = (int)(u1 - u0), nu = 2 * np;          // units and half-units of this CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < NXC; ++s) { mbar_init(x_full + s, 1); mbar_init(x_empty + s, 4); }
    for (int s = 0; s < NWC; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, want_grad ? 2 : 1); }
    for (int s = 0; s < NG; ++s) {
      mbar_init(acc1_full + s, 1); mbar_init(acc1_empty + s, 4);
      mbar_init(a2_full + s, 4);   mbar_init(a2_empty + s, 1);
    }
    mbar_init(r_full, 8);
    for (int s = 0; s < NG; ++s) mbar_init(acc2_full + s, 1);
    fence_barrier_init();
  }
  if (warp == CTRL_WARPS) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the set-up above and the first ring fills of the step's constants -- x, and for the row pass the fc11.weight
  // tiles -- overlap the tail of the previous kernel; h10 (R of the row pass, T of the gene pass) and every store wait
  pdl_trigger();
  if (!(warp == 0 || (!GENE && warp == 1))) pdl_wait();

  const int t_first = (int)(u0 / KP), kp_first = (int)(u0 - (int64_t)t_first * KP);
  const int kt_first = 2 * kp_first;
  const int rb_first = t_first / a.batch, arm_first = t_first - rb_first * a.batch;

  if (warp == 0) {
    // ===== TMA producer: raw x tiles, one per half-unit =====
    int kt = kt_first, rb = rb_first, arm = arm_first, sx = 0;
    uint32_t phx = 1;
    for (int j = 0; j < nu; ++j) {
      const int xb = a.x_batched ? arm : 0;
      WAIT(x_empty + sx, phx, 1, j);
      STAMP(0, j);
      if (elect_one()) {
        if (!GENE) {
          const int g0 = kt * UN;
          const int nb = max(0, min(UN, a.D - g0)) * 4;                              // bias bytes of this half-unit (D % 4 == 0)
          mbar_expect_tx(x_full + sx, X_BYTES + nb);
          tma_load_3d(&tmX, x_full + sx, xs(sx), g0, rb * 128, xb);                  // [128 cells][32 genes], SW128
          if (nb > 0) bulk_load_1d(bias_s + (j % NBS) * UN, a.bias + (int64_t)arm * a.bias_arm_stride + g0, nb, x_full + sx);
        } else {
          mbar_expect_tx(x_full + sx, X_BYTES);
          tma_load_3d(&tmX, x_full + sx, xs(sx), rb * 128, kt * UN, xb);             // [32 cells][128 genes], linear
        }
      }
      __syncwarp();
      if (++sx == NXC) { sx = 0; phx ^= 1; }
      if (++kt == KT) { kt = 0; if (++arm == a.batch) { arm = 0; ++rb; } }
    }
  } else if (warp == 1) {
    // ===== TMA producer: the T tile of a unit (64 rows x H), four slabs of 32 columns =====
    int kp = kp_first, arm = arm_first, sw = 0;
    uint32_t phw = 1;
    for (int i = 0; i < np; ++i) {
      WAIT(w_empty + sw, phw, 2, i);
      STAMP(1, i);
      if (elect_one()) {
        mbar_expect_tx(w_full + sw, IMG_BYTES);
#pragma unroll
        for (int q = 0; q < 4; ++q) tma_load_3d(&tmT, w_full + sw, ts(sw) + q * SLAB_BYTES, 32 * q, kp * 2 * UN, arm);
      }
      __syncwarp();
      if (++sw == NWC) { sw = 0; phw ^= 1; }
      if (++kp == KP) { kp = 0; if (++arm == a.batch) arm = 0; }
    }
  } else if (warp == 2) {
    // ===== MMA1 issuer: acc1[unit] = R . T^T, N = 64 (uniform loop, one elected lane issues).  MMA1 and MMA2 are issued
    // by two warps: one warp doing both spent ~1900 cycles per 32 columns in its own serial latencies (four satisfied
    // mbarrier waits at ~170 cycles each, commits, fences) and was the bottleneck of the kernel, not the tensor pipe.
    const uint32_t idesc1 = make_idesc(128, 2 * UN, false, false);
    const int ksteps1 = (a.H + 7) / 8;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    int kp = kp_first, sw = 0, seg = 0;
    uint32_t phw = 0;
    for (int i = 0; i < np; ++i) {
      const int g0 = 2 * (i & 1);                  // the two epilogue groups of this unit: g0, g0 + 1
      if ((i == 0) || (kp == 0)) {                 // new segment: R (and the drained acc2) must be in place
        WAIT(r_full, seg & 1, 3, i);
        ++seg;
      }
      const uint32_t pe = (uint32_t)(((i >> 1) & 1) ^ 1);
#ifdef F11_DEBUG
      WAIT(w_full + sw, phw, 4, i); WAIT(acc1_empty + g0, pe, 5, i); WAIT(acc1_empty + g0 + 1, pe, 6, i);
#else
      mbar_wait3(w_full + sw, phw, acc1_empty + g0, pe, acc1_empty + g0 + 1, pe);
#endif
      STAMP(2, i);
      tc_fence_after();
      const uint32_t ta = smem_u32(ts(sw));
      if (elect_one()) {
        // K-major read of the BASE32B image: row groups are its 4-row atoms (SBO = 512); k-step ks = 32 bytes inside
        // slab ks / 4.  Fully unrolled with one descriptor per unit: the k-step offsets are immediates.
        const uint64_t kd0 = make_desc_lt(ta, 0, 512, LT_BASE32B);
        const uint32_t dacc = tb + COL_ACC1 + (uint32_t)g0 * 32u, aR = tb + COL_R;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          if (ks < ksteps1)
            umma_tf32_ts(dacc, aR + ks * 8, kd0 + (uint64_t)(((ks >> 2) * SLAB_BYTES + (ks & 3) * 32) >> 4), idesc1, ks > 0 ? 1u : 0u);
        umma_commit(acc1_full + g0);
        umma_commit(acc1_full + g0 + 1);
        umma_commit(w_empty + sw);
        if (!want_grad && ((kp + 1 == KP) || (i == np - 1))) umma_commit(acc2_full + ((seg - 1) & (NG - 1)));   // loss-only: segment end marker
      }
      __syncwarp();
      if (++sw == NWC) { sw = 0; phw ^= 1; }
      if (++kp == KP) kp = 0;
    }
  } else if (warp == 3) {
    // ===== MMA2 issuer: acc2 += dY[g] . T (rows 32 s .. 32 s + 31 of the unit's tile, MN-major; dY was written to TMEM
    // by epilogue group g) =====
    if (want_grad) {
      const uint32_t idesc2 = make_idesc(128, a.HN, false, true);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      int kt = kt_first, sw = 0, seg = 0;
      uint32_t phw = 0, acc2 = 0;
      for (int j = 0; j < nu; ++j) {
        const int g = j & (NG - 1), half = j & 1;
        const bool last = (kt + 1 == KT) || (j == nu - 1);       // the segment ends with this half-unit
#ifdef F11_DEBUG
        WAIT(w_full + sw, phw, 7, j); WAIT(a2_full + g, (j / NG) & 1, 8, j);
#else
        mbar_wait2(w_full + sw, phw, a2_full + g, (j / NG) & 1);
#endif
        STAMP(3, j);
        tc_fence_after();
        const uint32_t ta = smem_u32(ts(sw)) + (uint32_t)half * 4096u;
        if (elect_one()) {
          const uint64_t md0 = make_desc_lt(ta, SLAB_BYTES, 512, LT_BASE32B);
          const uint32_t a2 = tb + COL_A2 + (uint32_t)g * 32u;
#pragma unroll
          for (int ks = 0; ks < UN / 8; ++ks)
            umma_tf32_ts(tb + COL_ACC2, a2 + ks * 8, md0 + (uint64_t)((ks * 1024) >> 4), idesc2, (acc2 | (uint32_t)ks) ? 1u : 0u);
          umma_commit(a2_empty + g);
          if (half) umma_commit(w_empty + sw);
          if (last) umma_commit(acc2_full + (seg & (NG - 1)));
        }
        __syncwarp();
        acc2 = last ? 0u : 1u;
        if (last) ++seg;
        if (half && ++sw == NWC) { sw = 0; phw ^= 1; }
        if (++kt == KT) kt = 0;
      }
    }
  } else {
    // ===== epilogue groups =====
    const int quad = warp & 3, grp = (warp - CTRL_WARPS) >> 2;
    const int hs = grp & 1;                                // which half of its units this group owns
    const int r = quad * 32 + lane;                        // TMEM lane = row of R within the block
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const uint32_t tacc1 = tmem_base + lane_bits + COL_ACC1 + (uint32_t)grp * 32u;
    const uint32_t ta2 = tmem_base + lane_bits + COL_A2 + (uint32_t)grp * 32u;
    const uint32_t tR = tmem_base + lane_bits + COL_R;
    // the two groups of a segment's last unit share the segment change: k-steps of R and 16-column chunks of acc2

    // this group's k-steps of the R block of tile (rb, arm): global -> registers (all loads in flight together) ...
    auto fetch_R = [&](uint32_t (&rv)[8][8], int rb, int arm) {
      const int ksteps1 = (a.H + 7) / 8, kc0 = hs ? (ksteps1 + 1) / 2 : 0, kc1 = hs ? ksteps1 : (ksteps1 + 1) / 2;
      const int row = rb * 128 + r;
      const float* src = a.R + (int64_t)arm * a.r_arm_stride + (int64_t)row * a.H;
      const bool ok = row < a.r_rows;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int kc = kc0 + c;
        float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
        if (ok && kc < kc1) {
          f0 = __ldg(reinterpret_cast<const float4*>(src + 8 * kc));
          if (8 * kc + 4 < a.H) f1 = __ldg(reinterpret_cast<const float4*>(src + 8 * kc + 4));
        }
        rv[c][0] = __float_as_uint(f0.x); rv[c][1] = __float_as_uint(f0.y); rv[c][2] = __float_as_uint(f0.z); rv[c][3] = __float_as_uint(f0.w);
        rv[c][4] = __float_as_uint(f1.x); rv[c][5] = __float_as_uint(f1.y); rv[c][6] = __float_as_uint(f1.z); rv[c][7] = __float_as_uint(f1.w);
      }
    };
    // ... -> TMEM (rows beyond r_rows and columns beyond H are zero)
    auto store_R = [&](const uint32_t (&rv)[8][8]) {
      const int ksteps1 = (a.H + 7) / 8, kc0 = hs ? (ksteps1 + 1) / 2 : 0, kc1 = hs ? ksteps1 : (ksteps1 + 1) / 2;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (kc0 + c < kc1) tmem_st8(tR + 8 * (kc0 + c), rv[c]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(r_full);
    };

    int kt = kt_first + grp, t = t_first, rb = rb_first, arm = arm_first;
    while (kt >= KT) { kt -= KT; ++t; if (++arm == a.batch) { arm = 0; ++rb; } }
    int sx = grp % NXC;
    uint32_t phx = (uint32_t)((grp / NXC) & 1), ph1 = 0, pha = 1;
    if (grp < 2) {
      uint32_t rv[8][8];
      fetch_R(rv, rb_first, arm_first);
      store_R(rv);
    }      // (np >= 1: the grid never exceeds the unit count)
    double sse = 0.0, mism = 0.0;      // ROW: loss partial sums of the current tile
    float dbsum = 0.f;                 // GENE: d fc11.bias partial of the current tile
    float bj = 0.f;                    // GENE: bias of this thread's gene
    int t_cur = -1;
    const float* __restrict__ bias_arm = a.bias;
    bool row_ok = false;
    for (int i = grp; i < nu; i += NG) {
      if (t != t_cur) {                 // this group's first half-unit of a tile
        t_cur = t;
        bias_arm = a.bias + (int64_t)arm * a.bias_arm_stride;
        row_ok = rb * 128 + r < a.r_rows;
        if (GENE) bj = row_ok ? __ldg(bias_arm + rb * 128 + r) : 0.f;
      }
      const int c0 = kt * UN;           // first gene (ROW) / cell (GENE) of the half-unit
      if (quad == 0) STAMP(9, i);
#ifdef F11_DEBUG
      WAIT(x_full + sx, phx, 9, i);
      const bool acc1_ready = false;
#else
      // both barriers are polled together (a satisfied try_wait still costs its ~170 cycles of latency)
      const bool x_ready = mbar_try_wait(x_full + sx, phx), acc1_ready = mbar_try_wait(acc1_full + grp, ph1);
      if (!x_ready) mbar_wait(x_full + sx, phx);
#endif
      if (quad == 0) STAMP(4, i);
      const uint8_t* tile = xs(sx);
      // (x is read from the slot chunk by chunk inside the math below: holding all 32 values next to the 32 accumulator
      // columns spilled the loop state to local memory, which does not fit the small L1 beside 192 KB of shared memory --
      // the stamps showed 730 cycles of "bookkeeping" per half-unit)
      if (!acc1_ready) WAIT(acc1_full + grp, ph1, 10, i);
      if (quad == 0) STAMP(5, i);
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld16(tacc1, acc);
      tmem_ld16(tacc1 + 16u, acc + 16);
      tmem_ld_wait();
      if (quad == 0) STAMP(6, i);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc1_empty + grp);        // the accumulator is in registers (tcgen05.wait::ld)
      bool a2_waited = false;
      float fs = 0.f, fm = 0.f;
      const float rscale = row_ok ? a.gscale : 0.f;        // rows beyond the batch contribute no gradient
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t dy[16];
        if (!GENE) {
          const int g0 = c0 + 16 * half;
          const float* bsl = bias_s + (i % NBS) * UN + 16 * half;      // delivered with the x tile (broadcast reads)
          const bool tail = g0 + 16 > a.D;                              // beyond the matrix the bias slot holds stale values
          float xh[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = *reinterpret_cast<const float4*>(bsl + 4 * q);
            const float4 x4 = *reinterpret_cast<const float4*>(tile + r * 128 + (((4 * half + q) ^ (r & 7)) << 4));
            float bq[4] = {b4.x, b4.y, b4.z, b4.w};
            const float xq[4] = {x4.x, x4.y, x4.z, x4.w};
            if (tail) {
#pragma unroll
              for (int e = 0; e < 4; ++e) bq[e] = g0 + 4 * q + e < a.D ? bq[e] : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * q + e;
              const float xin = xq[e];
              xh[j] = fmaxf(__uint_as_float(acc[16 * half + j]) + bq[e], 0.f);
              const float d = xh[j] - xin;
              fs = fmaf(d, d, fs);
              fm += ((xh[j] > 0.1f) != (xin > 0.1f)) ? 1.f : 0.f;
              dy[j] = __float_as_uint(xh[j] > 0.f ? rscale * d : 0.f);
            }
          }
          if (x_rec && row_ok) {
            float* xr = x_rec + (int64_t)arm * a.xrec_arm_stride + (int64_t)(rb * 128 + r) * a.D + g0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (g0 + j < a.D) xr[j] = xh[j];
          }
        } else {
          const int cell0 = c0 + 16 * half;
          const bool all = cell0 + 16 <= a.B;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float xh = fmaxf(__uint_as_float(acc[16 * half + j]) + bj, 0.f);
            const float xin = *reinterpret_cast<const float*>(tile + (16 * half + j) * 512 + r * 4);
            float v = xh > 0.f ? a.gscale * (xh - xin) : 0.f;
            if (!all && cell0 + j >= a.B) v = 0.f;
            dbsum += v;
            dy[j] = __float_as_uint(v);
          }
        }
        if (want_grad) {
          if (!a2_waited) {
            WAIT(a2_empty + grp, pha, 11, i);
            if (quad == 0) STAMP(7, i);
            tc_fence_after();
            a2_waited = true;
          }
          tmem_st8(ta2 + 16u * half, dy);
          tmem_st8(ta2 + 16u * half + 8u, dy + 8);
        }
      }
      // The raw slot goes back only here, after the last USE of the values read from it.  Handing it back right behind the
      // loads raced (round 1), and so did an arrive made dependent on an OR over all 32 loaded registers (round 2: run-to-run
      // differences in test_large_batch_unaligned_rows_run_to_run_identical) -- measured, not explained.
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty + sx);
      if (!GENE && row_ok) { sse += (double)fs; mism += (double)fm; }
      if (want_grad) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full + grp);
      }
      if (quad == 0) STAMP(8, i);
      ph1 ^= 1;
      pha ^= 1;
      // this half-unit belongs to the last unit of the CTA's share of tile t (both of that unit's groups see it)
      const bool seg_end = (kt >= KT - 2) || (i >= nu - 2);
      // position of this group's next half-unit
      int kt_n = kt + NG, t_n = t, rb_n = rb, arm_n = arm;
      while (kt_n >= KT) { kt_n -= KT; ++t_n; if (++arm_n == a.batch) { arm_n = 0; ++rb_n; } }
      const bool leaving = (t_n != t) || (i + NG >= nu);
      if (leaving) {
        // ---- this group's loss / bias-gradient partials of tile t
        if (!GENE) {
          const double s1 = warp_sum(sse), s2 = warp_sum(mism);
          if (lane == 0 && a.recon_acc) {
            atomicAdd(a.recon_acc + accl_recon(arm), s1);
            atomicAdd(a.recon_acc + accl_recon(arm) + 1, s2);
          }
          sse = 0.0; mism = 0.0;
        } else if (a.db_part) {
          a.db_part[(((int64_t)blockIdx.x + t) * NG + grp) * 128 + r] = dbsum;
          dbsum = 0.f;
        }
      }
      if (seg_end) {
        // ---- the R block of the next tile is fetched while the tensor pipe finishes the segment; then this group's
        // share of acc2 (the CTA's partial of tile t) is drained and its share of R stored
        const bool more = i < nu - 2;
        WAIT(acc2_full + ((t - t_first) & (NG - 1)), ((t - t_first) / NG) & 1, 12, i);
        tc_fence_after();
        if (want_grad) {
          float* prt = a.part + ((int64_t)blockIdx.x + t) * TILE_FLOATS + r * 128;
          const int nch = a.HN / 16, ch0 = hs ? (nch + 1) / 2 : 0, ch1 = hs ? nch : (nch + 1) / 2;
          for (int j = ch0; j < ch1; ++j) {
            uint32_t rr[16];
            tmem_ld16(tmem_base + lane_bits + COL_ACC2 + (uint32_t)(j * 16), rr);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 4; ++e)
              reinterpret_cast<float4*>(prt + j * 16)[e] = make_float4(__uint_as_float(rr[4 * e]), __uint_as_float(rr[4 * e + 1]),
                                                                       __uint_as_float(rr[4 * e + 2]), __uint_as_float(rr[4 * e + 3]));
          }
          tc_fence_before();
        }
        if (more) {
          uint32_t rv[8][8];
          int arm2 = arm + 1, rb2 = rb;
          if (arm2 == a.batch) { arm2 = 0; ++rb2; }
          fetch_R(rv, rb2, arm2);
          store_R(rv);
        }
      }
      sx += NG;
      if (sx >= NXC) { sx -= NXC; phx ^= 1; }
      kt = kt_n; t = t_n; rb = rb_n; arm = arm_n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CTRL_WARPS) tmem_dealloc(tmem_base, 512);
}

// Fix-up of the stream-K partials.  One launch after the gene pass serves both passes (the row pass keeps its partial
// tiles in a region of their own until then; nothing between the two passes needs d h10):
//   rows:  d h10[arm][cell][h]      = sum over the CTAs that touched tile (cell / 128, arm) of their partial, in CTA order
//   genes: d fc11.weight[arm][gene][h] likewise;  d fc11.bias[arm][gene] = sum over CTAs and epilogue groups
struct F11FixOne {               // one stream-K pass
  const float* part; int batch, ktiles; int64_t U, G;
  float* out; int64_t out_arm_stride; int ld, rows, cols;
  int nblk;                      // blocks of this part: ceil(rows / 128) * 8 * batch
};
struct F11FixArgs {
  F11FixOne row, gene;
  const float* db_part; float* db_out; int64_t db_arm_stride; int D;
  int nblk_db;                   // ceil(D / 256) * batch blocks (256 genes each)
  int b0;                        // first block of this launch in the (rows | genes | bias) sequence
};

// block = 16 rows of one 128-row tile; thread = (row mod 8, float4 column): 2 rows each (cols % 4 == 0 on this path)
__device__ __forceinline__ void f11_fix_tile(const F11FixOne& p, int bx, int arm, int* cc) {
  const int tid = threadIdx.x, r8 = tid >> 5, c4 = tid & 31;
  const int row0 = (bx >> 3) * 128, sub0 = (bx & 7) * 16;
  const int64_t t = (int64_t)(bx >> 3) * p.batch + arm;
  if (tid == 0) {
    cc[0] = (int)cta_of_unit(t * p.ktiles, p.U, p.G);
    cc[1] = (int)cta_of_unit(t * p.ktiles + p.ktiles - 1, p.U, p.G);
  }
  __syncthreads();
  if (4 * c4 >= p.cols) return;
  const int c0 = cc[0], c1 = cc[1];
  const float* base = p.part + (int64_t)(c0 + t) * TILE_FLOATS + 4 * c4;
  float* o = p.out + (int64_t)arm * p.out_arm_stride + 4 * c4;
  for (int k = 0; k < 2; ++k) {
    const int rl = sub0 + r8 + 8 * k, row = row0 + rl;
    if (row >= p.rows) break;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c0; c <= c1; ++c) {
      const float4 q = *reinterpret_cast<const float4*>(base + (int64_t)(c - c0) * TILE_FLOATS + rl * 128);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    *reinterpret_cast<float4*>(o + (int64_t)row * p.ld) = v;
  }
}

// a (CTA, group) pair that processed no half-unit of the tile wrote nothing: group g owns the half-units j = g (mod NG)
// counted from the CTA's first one.  (ktiles, U: units = pairs of half-units, as in the main kernel)
__device__ __forceinline__ void f11_fix_bias(const F11FixArgs& p, int gene, int arm) {
  if (gene >= p.D) return;
  const F11FixOne& q = p.gene;
  const int64_t t = (int64_t)(gene >> 7) * q.batch + arm;
  const int64_t ua = t * q.ktiles, ub = ua + q.ktiles;        // units of the tile
  const int64_t c0 = cta_of_unit(ua, q.U, q.G), c1 = cta_of_unit(ub - 1, q.U, q.G);
  float v = 0.f;
  for (int64_t c = c0; c <= c1; ++c) {
    const int64_t s0 = c * q.U / q.G, s1 = (c + 1) * q.U / q.G;     // units of CTA c
    const int64_t lo = 2 * (ua > s0 ? ua : s0), hi = 2 * (ub < s1 ? ub : s1);   // half-units of the tile in CTA c
    for (int g = 0; g < NG; ++g) {
      const int64_t first = lo + ((g - (lo - 2 * s0)) % NG + NG) % NG;           // first half-unit >= lo of group g
      if (first < hi) v += p.db_part[((c + t) * NG + g) * 128 + (gene & 127)];
    }
  }
  p.db_out[(int64_t)arm * p.db_arm_stride + gene] = v;
}

__global__ void __launch_bounds__(256) f11_fixup_kernel(const F11FixArgs p) {
  __shared__ int cc[2];
  int b = blockIdx.x + p.b0;
  pdl_trigger();
  pdl_wait();
  if (b < p.row.nblk) {
    f11_fix_tile(p.row, b / p.row.batch, b % p.row.batch, cc);
  } else if ((b -= p.row.nblk) < p.gene.nblk) {
    f11_fix_tile(p.gene, b / p.gene.batch, b % p.gene.batch, cc);
  } else {
    b -= p.gene.nblk;
    f11_fix_bias(p, (b / p.gene.batch) * 256 + threadIdx.x, b % p.gene.batch);
  }
}

static F11FixOne fix_one(const float* part, int batch, int ktiles, int64_t U, int64_t G, float* out, int64_t out_arm_stride, int ld,
                         int rows, int cols) {
  F11FixOne f;
  f.part = part; f.batch = batch; f.ktiles = ktiles; f.U = U; f.G = G;
  f.out = out; f.out_arm_stride = out_arm_stride; f.ld = ld; f.rows = rows; f.cols = cols;
  f.nblk = (rows + 127) / 128 * 8 * batch;
  return f;
}

int sm_count2() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0 || n > 160) n = 148;
  }
  return n;
}

template <bool GENE, bool TRAIN>
int launch_f11(const CUtensorMap& tmX, const CUtensorMap& tmT, F11Args& a, int64_t* U_out, int64_t* G_out, cudaStream_t s) {
  const int64_t U = (int64_t)a.batch * a.rtiles * a.ktiles;
  int64_t G = sm_count2();
  if (G > U) G = U;
  // a T tile lives from its load until MMA2 of its second half: one being loaded, MMA1 up to two units ahead of the
  // epilogue, MMA2 behind it -> 4 stages.  The x ring depth must be a multiple of NG, so that a slot is always refilled
  // for the group that emptied it: with any other depth a slot alternates between two groups, and a group can reach
  // "its" fill k while fill k - 1 (the other group's; TMA completions are not ordered) is still in flight -- the parity
  // wait for fill k then succeeds on the completed fill k - 2 (stale tile, early release, two fills pending on one
  // barrier: intermittent launch failures when x rows are not 128-byte aligned, D = 5032 with 6 slots).
  a.nw = NWC;
  a.nx = NXC;
  static_assert((size_t)NXC * X_BYTES + (size_t)NWC * IMG_BYTES + 64 * 8 + (size_t)NBS * UN * 4 + 1024 <= 227 * 1024, "rings exceed the shared memory of an SM");
  const size_t smem = (size_t)a.nx * X_BYTES + (size_t)a.nw * IMG_BYTES + 64 * 8 + (size_t)NBS * UN * 4 + 1024;   // rings | barriers | bias ring | alignment
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr[64] = {};
  if (dev >= 0 && dev < 64 && !attr[dev]) {
    MVAE_CUDA(cudaFuncSetAttribute(fc11_ts_kernel<GENE, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr[dev] = true;
  }
  launch_pdl(fc11_ts_kernel<GENE, TRAIN>, dim3((unsigned)G), dim3(THREADS), smem, s, tmX, tmT, a);
  {
    ::mvae::g_launches++;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      ::mvae::set_error("fc11 %s pass: kernel launch failed: %s (%s:%d)", GENE ? "gene" : "row", cudaGetErrorString(e), __FILE__, __LINE__);
      return (int)e;
    }
  }
  *U_out = U; *G_out = G;
  return 0;
}

}  // namespace

// x_hat / loss sums / d h10 in one pass over x (x_rec optional: materialised reconstruction)
static int fc11_rows(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale, int want_grad,
                     float* x_rec, double* recon_acc, F11FixOne* fix, cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  float* work = st.work;
  F11Args a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.D = D; a.H = H; a.HN = (H + 15) / 16 * 16;
  a.batch = A; a.rtiles = (B + 127) / 128; a.ktiles = (D + 2 * UN - 1) / (2 * UN);
  a.x_batched = in.x_arm_stride > 0;
  a.want_grad = want_grad; a.gscale = gscale;
  a.R = work + w.d[4]; a.r_arm_stride = (int64_t)B * H; a.r_rows = B;
  a.bias = st.params + L.offset[FC11_B]; a.bias_arm_stride = L.arm_stride;
  a.x_rec = x_rec; a.xrec_arm_stride = (int64_t)B * D;
  a.recon_acc = recon_acc;
  a.part = work + w.fc1_part;
  CUtensorMap tmX, tmT;
  int rc = make_map_ex(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, 32, 128, 1);
  if (rc) return rc;
  rc = make_map_ex(&tmT, st.params + L.offset[FC11_W], H, D, H, A, L.arm_stride, 32, 2 * UN, 2);
  if (rc) return rc;
  int64_t U, G;
  rc = (want_grad && !x_rec) ? launch_f11<false, true>(tmX, tmT, a, &U, &G, s)
                             : launch_f11<false, false>(tmX, tmT, a, &U, &G, s);
  if (rc || !want_grad) return rc;
  if (fix) {           // the gene pass that follows sums these partials in its own fix-up launch
    *fix = fix_one(a.part, A, a.ktiles, U, G, work + w.g_d10, (int64_t)B * H, H, B, H);
    return 0;
  }
  F11FixArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.row = fix_one(a.part, A, a.ktiles, U, G, work + w.g_d10, (int64_t)B * H, H, B, H);
  launch_pdl(f11_fixup_kernel, dim3(fa.row.nblk), dim3(256), 0, s, fa);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// d fc11.weight and d fc11.bias by the gene pass (x_hat recomputed)
static thread_local F11FixArgs tl_gene_fix;
static thread_local bool tl_gene_fix_pending = false;

static int fc11_genes(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale,
                      const F11FixOne* row_fix, cudaStream_t s, bool defer_gene_fix) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  float* work = st.work;
  F11Args a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.D = D; a.H = H; a.HN = (H + 15) / 16 * 16;
  a.batch = A; a.rtiles = (D + 127) / 128; a.ktiles = (B + 2 * UN - 1) / (2 * UN);
  a.x_batched = in.x_arm_stride > 0;
  a.want_grad = 1; a.gscale = gscale;
  a.R = st.params + L.offset[FC11_W]; a.r_arm_stride = L.arm_stride; a.r_rows = D;
  a.bias = st.params + L.offset[FC11_B]; a.bias_arm_stride = L.arm_stride;
  // partial tiles behind those of the row pass (slot = CTA + tile < tiles + #SM): layout.cu sizes the region for both
  a.part = work + w.fc1_part + (int64_t)f11_gene_slot0(A, B) * TILE_FLOATS;
  a.db_part = work + w.db_part;
  CUtensorMap tmX, tmT;
  int rc = make_map_ex(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, 128, 32, 0);
  if (rc) return rc;
  rc = make_map_ex(&tmT, work + w.d[4], H, B, H, A, (int64_t)B * H, 32, 2 * UN, 2);
  if (rc) return rc;
  int64_t U, G;
  rc = launch_f11<true, true>(tmX, tmT, a, &U, &G, s);
  if (rc) return rc;
  F11FixArgs fa;
  memset(&fa, 0, sizeof(fa));
  if (row_fix) fa.row = *row_fix;
  fa.gene = fix_one(a.part, A, a.ktiles, U, G, st.grads + L.offset[FC11_W], L.arm_stride, H, D, H);
  fa.db_part = a.db_part; fa.db_out = st.grads + L.offset[FC11_B]; fa.db_arm_stride = L.arm_stride; fa.D = D;
  fa.nblk_db = (D + 255) / 256 * A;
  if (defer_gene_fix && row_fix) {
    launch_pdl(f11_fixup_kernel, dim3(fa.row.nblk), dim3(256), 0, s, fa);            // d h10: the backward chain waits for it
    MVAE_LAUNCH_CHECK();
    tl_gene_fix = fa;
    tl_gene_fix.b0 = fa.row.nblk;
    tl_gene_fix_pending = true;
    return 0;
  }
  launch_pdl(f11_fixup_kernel, dim3(fa.row.nblk + fa.gene.nblk + fa.nblk_db), dim3(256), 0, s, fa);
  MVAE_LAUNCH_CHECK();
  return 0;
}

int ts_fc11_gene_fixup(cudaStream_t s) {
  if (!tl_gene_fix_pending) return 1;
  tl_gene_fix_pending = false;
  const F11FixArgs& fa = tl_gene_fix;
  const int main_pdl = tl_pdl;
  tl_pdl = 0;                          // an ordinary launch (side stream)
  launch_pdl(f11_fixup_kernel, dim3(fa.gene.nblk + fa.nblk_db), dim3(256), 0, s, fa);
  tl_pdl = main_pdl;
  MVAE_LAUNCH_CHECK();
  return 0;
}

// x_hat / loss sums (/ d h10) alone: the evaluation paths and mvae_forward's materialised reconstruction
int ts_fc11_rows(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale, int want_grad,
                 float* x_rec, double* recon_acc, cudaStream_t s) {
  return fc11_rows(d, st, in, w, gscale, want_grad, x_rec, recon_acc, nullptr, s);
}

// both passes of the training step: row owner (loss sums, d h10), gene owner (d fc11.weight, d fc11.bias), one fix-up
int ts_fc11_loss_grad(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale,
                      double* recon_acc, cudaStream_t s, bool defer_gene_fix) {
  F11FixOne row_fix;
  int rc = fc11_rows(d, st, in, w, gscale, 1, nullptr, recon_acc, &row_fix, s);
  if (rc) return rc;
  return fc11_genes(d, st, in, w, gscale, &row_fix, s, defer_gene_fix);
}

}  // namespace mvae

#ifdef F11_STAMPS
extern "C" int mvae_debug_f11_stamps(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, mvae::g_f11_stamps, sizeof(mvae::g_f11_stamps));
}
#endif
