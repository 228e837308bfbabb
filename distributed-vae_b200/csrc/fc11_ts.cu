// fc11_ts.cu — the last decoder layer fused with the reconstruction loss and its own backward, second generation.
//
// Reference ops (mmidas/nn_model.py): x_hat = relu(fc11(h10)) :287; 0.5*mse_sum/B + 0.5*BCE(bin(x_hat), bin(x))
// :542-546; autograd of both (dY = max(A-1,1)/B * (x_hat - x) * [x_hat > 0]; the BCE half acts on constants).
// Two passes over x, nothing [B,D]-sized is ever written:
//
//   ROW  pass (thread = cell):  X[128 cells x 32 genes] = h10 . W11^T ; loss sums ; dY ; d h10 += dY . W11
//   GENE pass (thread = gene):  X^T[128 genes x 32 cells] = W11 . h10^T ; dY^T ; d fc11.weight += dY^T . h10 ; d fc11.bias
//
// Same skeleton for both (roles of h10 and W11 swapped through the tensor maps):
//   * the resident operand R (128 rows x H: h10 block / W11 block) lives in TENSOR MEMORY and feeds MMA1 in the TS form;
//   * streamed per unit (32 columns): the raw x tile (own deep ring: it comes from HBM) and the small operand tile in
//     two images (K-major for MMA1, MN-major for MMA2; separate rings, L2-resident);
//   * MMA1 -> acc1[g] (TMEM) -> 4 epilogue groups (unit i belongs to group i % 4; lane = row of R) compute x_hat, the
//     loss terms and dY in registers and store dY as the A operand of MMA2 straight into TMEM (tcgen05.st);
//   * MMA2 (TS form) accumulates into acc2 (TMEM) over the whole segment;
//   * stream-K over (tile, unit): one CTA per SM, one wave; partial tiles are summed in a fixed order by a fix-up kernel.
// TMEM columns: acc2 [0,128) | acc1 4 x 32 [128,256) | dY stages 4 x 32 [256,384) | R [384,512).
#include "gemm_tc.h"
#include "tc_common.cuh"

namespace mvae {

namespace {
using namespace tc;

constexpr int UN = 32;                     // columns per unit (genes for ROW, cells for GENE)
constexpr int NG = 4;                      // epilogue groups (4 warps each)
constexpr int CTRL_WARPS = 4;              // x TMA, small-operand TMA, MMA1 issue, MMA2 issue
constexpr int THREADS = 32 * (CTRL_WARPS + 4 * NG);
constexpr int X_BYTES = 16384;
constexpr int IMG_BYTES = 16384;           // one image of the small operand tile: 4 slabs of [32 rows x 128 B]
constexpr int TILE_FLOATS = 128 * 128;
constexpr uint32_t COL_ACC2 = 0, COL_ACC1 = 128, COL_A2 = 256, COL_R = 384;

struct F11Args {
  int B, D, H, HN;
  int batch, rtiles, ktiles;        // arms, 128-row blocks of R per arm, units per tile
  int x_batched;
  int nx, nk, nm;                   // ring depths: x, K-image, MN-image
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

// TRAIN = true: the training-step instance (gradients wanted, no materialised reconstruction): the flags are compile-time
// so the epilogue carries no per-element branches for them.
template <bool GENE, bool TRAIN>
__global__ void __launch_bounds__(THREADS, 1)
fc11_ts_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmTk,
               const __grid_constant__ CUtensorMap tmTm, const F11Args a) {
  const bool want_grad = TRAIN || a.want_grad;
  float* const x_rec = TRAIN ? nullptr : a.x_rec;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto xs = [&](int s) { return smem + (size_t)s * X_BYTES; };
  auto tk = [&](int s) { return smem + (size_t)a.nx * X_BYTES + (size_t)s * IMG_BYTES; };
  auto tm = [&](int s) { return smem + (size_t)a.nx * X_BYTES + (size_t)(a.nk + s) * IMG_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.nx * X_BYTES + (size_t)(a.nk + a.nm) * IMG_BYTES);
  uint64_t* x_full = bars;                 uint64_t* x_empty = x_full + a.nx;
  uint64_t* k_full = x_empty + a.nx;       uint64_t* k_empty = k_full + a.nk;
  uint64_t* m_full = k_empty + a.nk;       uint64_t* m_empty = m_full + a.nm;
  uint64_t* acc1_full = m_empty + a.nm;    uint64_t* acc1_empty = acc1_full + NG;
  uint64_t* a2_full = acc1_empty + NG;     uint64_t* a2_empty = a2_full + NG;
  uint64_t* r_full = a2_empty + NG;        // R of the current segment is in TMEM (and acc2 of the previous one drained)
  // segment s complete: barrier s % NG.  An epilogue group can be up to NG units -- hence NG segments when a CTA's share
  // of a tile is a single unit -- ahead of the tensor pipe; one parity bit cannot tell those apart, NG barriers can.
  uint64_t* acc2_full = r_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_full + NG);

  const int KT = a.ktiles;
  const int64_t U = (int64_t)a.batch * a.rtiles * KT, G = gridDim.x;
  const int64_t u0 = (int64_t)blockIdx.x * U / G, u1 = ((int64_t)blockIdx.x + 1) * U / G;
  const int nu = (int)(u1 - u0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nx; ++s) { mbar_init(x_full + s, 1); mbar_init(x_empty + s, 4); }
    for (int s = 0; s < a.nk; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int s = 0; s < a.nm; ++s) { mbar_init(m_full + s, 1); mbar_init(m_empty + s, 1); }
    for (int s = 0; s < NG; ++s) {
      mbar_init(acc1_full + s, 1); mbar_init(acc1_empty + s, 4);
      mbar_init(a2_full + s, 4);   mbar_init(a2_empty + s, 1);
    }
    mbar_init(r_full, 4);
    for (int s = 0; s < NG; ++s) mbar_init(acc2_full + s, 1);
    fence_barrier_init();
  }
  if (warp == CTRL_WARPS) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int t_first = (int)(u0 / KT), kt_first = (int)(u0 - (int64_t)t_first * KT);
  const int rb_first = t_first / a.batch, arm_first = t_first - rb_first * a.batch;

  if (warp == 0) {
    // ===== TMA producer: raw x tiles =====
    int kt = kt_first, rb = rb_first, arm = arm_first, sx = 0;
    uint32_t phx = 1;
    for (int i = 0; i < nu; ++i) {
      const int xb = a.x_batched ? arm : 0;
      mbar_wait(x_empty + sx, phx);
      if (elect_one()) {
        mbar_expect_tx(x_full + sx, X_BYTES);
        if (!GENE) tma_load_3d(&tmX, x_full + sx, xs(sx), kt * UN, rb * 128, xb);    // [128 cells][32 genes], SW128
        else tma_load_3d(&tmX, x_full + sx, xs(sx), rb * 128, kt * UN, xb);          // [32 cells][128 genes], linear
      }
      __syncwarp();
      if (++sx == a.nx) { sx = 0; phx ^= 1; }
      if (++kt == KT) { kt = 0; if (++arm == a.batch) { arm = 0; ++rb; } }
    }
  } else if (warp == 1) {
    // ===== TMA producer: the small operand tile (32 rows x H) in its two images =====
    int kt = kt_first, arm = arm_first, sk = 0, sm = 0;
    uint32_t phk = 1, phm = 1;
    for (int i = 0; i < nu; ++i) {
      mbar_wait(k_empty + sk, phk);
      if (elect_one()) {
        mbar_expect_tx(k_full + sk, IMG_BYTES);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_3d(&tmTk, k_full + sk, tk(sk) + j * 4096, 32 * j, kt * UN, arm);
      }
      __syncwarp();
      if (want_grad) {
        mbar_wait(m_empty + sm, phm);
        if (elect_one()) {
          mbar_expect_tx(m_full + sm, IMG_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(&tmTm, m_full + sm, tm(sm) + j * 4096, 32 * j, kt * UN, arm);
        }
        __syncwarp();
        if (++sm == a.nm) { sm = 0; phm ^= 1; }
      }
      if (++sk == a.nk) { sk = 0; phk ^= 1; }
      if (++kt == KT) { kt = 0; if (++arm == a.batch) arm = 0; }
    }
  } else if (warp == 2) {
    // ===== MMA1 issuer: acc1[g] = R . T_K^T (uniform loop, one elected lane issues).  MMA1 and MMA2 are issued by two
    // warps: one warp doing both spent ~1900 cycles per unit in its own serial latencies (four satisfied mbarrier
    // waits at ~170 cycles each, commits, fences) and was the bottleneck of the kernel, not the tensor pipe.
    const uint32_t idesc1 = make_idesc(128, UN, false, false);
    const int ksteps1 = (a.H + 7) / 8;
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    int kt = kt_first, sk = 0, seg = 0;
    uint32_t phk = 0;
    for (int i = 0; i < nu; ++i) {
      const int g = i & (NG - 1);
      if ((i == 0) || (kt == 0)) {                 // new segment: R (and the drained acc2) must be in place
        mbar_wait(r_full, seg & 1);
        ++seg;
      }
      mbar_wait2(k_full + sk, phk, acc1_empty + g, ((i / NG) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tka = smem_u32(tk(sk));
      if (elect_one()) {
        // fully unrolled with one descriptor per unit: the k-step offsets are immediates
        const uint64_t kd0 = make_smem_desc(tka, 0, 1024, false);
        const uint32_t dacc = tb + COL_ACC1 + (uint32_t)g * 32u, aR = tb + COL_R;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          if (ks < ksteps1)
            umma_tf32_ts(dacc, aR + ks * 8, kd0 + (uint64_t)(((ks >> 2) * 4096 + (ks & 3) * 32) >> 4), idesc1, ks > 0 ? 1u : 0u);
        umma_commit(acc1_full + g);
        umma_commit(k_empty + sk);
        if (!want_grad && ((kt + 1 == KT) || (i == nu - 1))) umma_commit(acc2_full + ((seg - 1) & (NG - 1)));   // loss-only: segment end marker
      }
      __syncwarp();
      if (++sk == a.nk) { sk = 0; phk ^= 1; }
      if (++kt == KT) kt = 0;
    }
  } else if (warp == 3) {
    // ===== MMA2 issuer: acc2 += dY[g] . T_MN (dY was written to TMEM by epilogue group g) =====
    if (want_grad) {
      const uint32_t idesc2 = make_idesc(128, a.HN, false, true);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      int kt = kt_first, sm = 0, seg = 0;
      uint32_t phm = 0, acc2 = 0;
      for (int j = 0; j < nu; ++j) {
        const int g = j & (NG - 1);
        const bool last = (kt + 1 == KT) || (j == nu - 1);       // the segment ends with this unit
        mbar_wait2(m_full + sm, phm, a2_full + g, (j / NG) & 1);
        tc_fence_after();
        const uint32_t tma = smem_u32(tm(sm));
        if (elect_one()) {
          const uint64_t md0 = make_smem_desc(tma, 4096, 512, true);
          const uint32_t a2 = tb + COL_A2 + (uint32_t)g * 32u;
#pragma unroll
          for (int ks = 0; ks < UN / 8; ++ks)
            umma_tf32_ts(tb + COL_ACC2, a2 + ks * 8, md0 + (uint64_t)((ks * 1024) >> 4), idesc2, (acc2 | (uint32_t)ks) ? 1u : 0u);
          umma_commit(a2_empty + g);
          umma_commit(m_empty + sm);
          if (last) umma_commit(acc2_full + (seg & (NG - 1)));
        }
        __syncwarp();
        acc2 = last ? 0u : 1u;
        if (last) ++seg;
        if (++sm == a.nm) { sm = 0; phm ^= 1; }
        if (++kt == KT) kt = 0;
      }
    }
  } else {
    // ===== epilogue groups =====
    const int quad = warp & 3, grp = (warp - CTRL_WARPS) >> 2;
    const int r = quad * 32 + lane;                        // TMEM lane = row of R within the block
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const uint32_t tacc1 = tmem_base + lane_bits + COL_ACC1 + (uint32_t)grp * 32u;
    const uint32_t ta2 = tmem_base + lane_bits + COL_A2 + (uint32_t)grp * 32u;
    const uint32_t tR = tmem_base + lane_bits + COL_R;
    const int ksteps1 = (a.H + 7) / 8;

    // R block of tile (rb, arm) -> TMEM (rows beyond r_rows and columns beyond H are zero)
    auto load_R = [&](int rb, int arm) {
      const int row = rb * 128 + r;
      const float* src = a.R + (int64_t)arm * a.r_arm_stride + (int64_t)row * a.H;
      const bool ok = row < a.r_rows;
      for (int c = 0; c < ksteps1; ++c) {
        uint32_t v[8];
        float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
        if (ok) {
          f0 = __ldg(reinterpret_cast<const float4*>(src + 8 * c));
          if (8 * c + 4 < a.H) f1 = __ldg(reinterpret_cast<const float4*>(src + 8 * c + 4));
        }
        v[0] = __float_as_uint(f0.x); v[1] = __float_as_uint(f0.y); v[2] = __float_as_uint(f0.z); v[3] = __float_as_uint(f0.w);
        v[4] = __float_as_uint(f1.x); v[5] = __float_as_uint(f1.y); v[6] = __float_as_uint(f1.z); v[7] = __float_as_uint(f1.w);
        tmem_st8(tR + 8 * c, v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(r_full);
    };

    int kt = kt_first + grp, t = t_first, rb = rb_first, arm = arm_first;
    while (kt >= KT) { kt -= KT; ++t; if (++arm == a.batch) { arm = 0; ++rb; } }
    int sx = grp % a.nx;
    uint32_t phx = (uint32_t)((grp / a.nx) & 1), ph1 = 0, pha = 1;
    if (grp == 0 && nu > 0) load_R(rb_first, arm_first);
    double sse = 0.0, mism = 0.0;      // ROW: loss partial sums of the current tile
    float dbsum = 0.f;                 // GENE: d fc11.bias partial of the current tile
    float bj = 0.f;                    // GENE: bias of this thread's gene
    int t_cur = -1;
    const float* __restrict__ bias_arm = a.bias;
    bool row_ok = false;
    for (int i = grp; i < nu; i += NG) {
      if (t != t_cur) {                 // this group's first unit of a tile
        t_cur = t;
        bias_arm = a.bias + (int64_t)arm * a.bias_arm_stride;
        row_ok = rb * 128 + r < a.r_rows;
        if (GENE) bj = row_ok ? __ldg(bias_arm + rb * 128 + r) : 0.f;
      }
      const int c0 = kt * UN;           // first gene (ROW) / cell (GENE) of the unit
      mbar_wait(x_full + sx, phx);
      const uint8_t* tile = xs(sx);
      // the x tile does not depend on MMA1: read it while the accumulator is still being produced
      float xv[32];
      if (!GENE) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t4 = *reinterpret_cast<const float4*>(tile + r * 128 + ((q ^ (r & 7)) << 4));
          xv[4 * q] = t4.x; xv[4 * q + 1] = t4.y; xv[4 * q + 2] = t4.z; xv[4 * q + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) xv[j] = *reinterpret_cast<const float*>(tile + j * 512 + r * 4);
      }
      mbar_wait(acc1_full + grp, ph1);
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld16(tacc1, acc);
      tmem_ld16(tacc1 + 16u, acc + 16);
      tmem_ld_wait();
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
          float bb[16];
          const int g0 = c0 + 16 * half;
          if (g0 + 16 <= a.D) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_arm + g0 + 4 * q));
              bb[4 * q] = b4.x; bb[4 * q + 1] = b4.y; bb[4 * q + 2] = b4.z; bb[4 * q + 3] = b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) bb[j] = g0 + j < a.D ? __ldg(bias_arm + g0 + j) : 0.f;
          }
          float xh[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float xin = xv[16 * half + j];
            xh[j] = fmaxf(__uint_as_float(acc[16 * half + j]) + bb[j], 0.f);
            const float d = xh[j] - xin;
            fs = fmaf(d, d, fs);
            fm += ((xh[j] > 0.1f) != (xin > 0.1f)) ? 1.f : 0.f;
            dy[j] = __float_as_uint(xh[j] > 0.f ? rscale * d : 0.f);
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
            float v = xh > 0.f ? a.gscale * (xh - xv[16 * half + j]) : 0.f;
            if (!all && cell0 + j >= a.B) v = 0.f;
            dbsum += v;
            dy[j] = __float_as_uint(v);
          }
        }
        if (want_grad) {
          if (!a2_waited) {
            mbar_wait(a2_empty + grp, pha);
            tc_fence_after();
            a2_waited = true;
          }
          tmem_st8(ta2 + 16u * half, dy);
          tmem_st8(ta2 + 16u * half + 8u, dy + 8);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(x_empty + sx);            // every value read from the x tile has been consumed by now
      if (!GENE && row_ok) { sse += (double)fs; mism += (double)fm; }
      if (want_grad) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a2_full + grp);
      }
      ph1 ^= 1;
      pha ^= 1;
      const bool seg_end = (kt == KT - 1) || (i == nu - 1);
      // position of this group's next unit
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
        // ---- drain acc2 (this CTA's share of tile t), then bring in R of the next tile
        mbar_wait(acc2_full + ((t - t_first) & (NG - 1)), ((t - t_first) / NG) & 1);
        tc_fence_after();
        if (want_grad) {
          float* prt = a.part + ((int64_t)blockIdx.x + t) * TILE_FLOATS + r * 128;
          for (int j = 0; j < a.HN / 16; ++j) {
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
        if (i != nu - 1) {
          int arm2 = arm + 1, rb2 = rb;
          if (arm2 == a.batch) { arm2 = 0; ++rb2; }
          load_R(rb2, arm2);
        }
      }
      // GENE: groups that did not see the tile at all must still define their db partial (zero) -- handled by the fix-up
      sx += NG;
      if (sx >= a.nx) { sx -= a.nx; phx ^= 1; }
      kt = kt_n; t = t_n; rb = rb_n; arm = arm_n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CTRL_WARPS) tmem_dealloc(tmem_base, 512);
}

// out[arm][row][col] = sum over the CTAs that touched tile (row / 128, arm) of their partial, in CTA order
__global__ void __launch_bounds__(256) f11_fixup_kernel(const float* part, int batch, int ktiles, int64_t U, int64_t G, float* out,
                                                        int64_t out_arm_stride, int ld, int rows, int cols) {
  // block = one 128-row tile; thread = (row mod 8, float4 column): 16 rows each (cols % 4 == 0 on this path)
  __shared__ int cc[2];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, r8 = tid >> 5, c4 = tid & 31;
  const int row0 = (blockIdx.x >> 3) * 128, sub0 = (blockIdx.x & 7) * 16;
  const int64_t t = (int64_t)(blockIdx.x >> 3) * batch + arm;
  if (tid == 0) {
    cc[0] = (int)cta_of_unit(t * ktiles, U, G);
    cc[1] = (int)cta_of_unit(t * ktiles + ktiles - 1, U, G);
  }
  __syncthreads();
  if (4 * c4 >= cols) return;
  const int c0 = cc[0], c1 = cc[1];
  const float* base = part + (int64_t)(c0 + t) * TILE_FLOATS + 4 * c4;
  float* o = out + (int64_t)arm * out_arm_stride + 4 * c4;
  for (int k = 0; k < 2; ++k) {
    const int rl = sub0 + r8 + 8 * k, row = row0 + rl;
    if (row >= rows) break;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c0; c <= c1; ++c) {
      const float4 q = *reinterpret_cast<const float4*>(base + (int64_t)(c - c0) * TILE_FLOATS + rl * 128);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    *reinterpret_cast<float4*>(o + (int64_t)row * ld) = v;
  }
}

// d fc11.bias[arm][gene] = sum over CTAs and epilogue groups of the partials; a (CTA, group) pair that processed no unit
// of the tile wrote nothing: its units are i = g (mod NG) counted from the CTA's first unit.
__global__ void __launch_bounds__(128) f11_db_fixup_kernel(const float* db_part, int batch, int ktiles, int64_t U, int64_t G,
                                                           float* out, int64_t out_arm_stride, int D) {
  const int arm = blockIdx.y;
  const int gene = blockIdx.x * 128 + threadIdx.x;
  if (gene >= D) return;
  const int64_t t = (int64_t)blockIdx.x * batch + arm;
  const int64_t ua = t * ktiles, ub = ua + ktiles;            // units of the tile
  const int64_t c0 = cta_of_unit(ua, U, G), c1 = cta_of_unit(ub - 1, U, G);
  float v = 0.f;
  for (int64_t c = c0; c <= c1; ++c) {
    const int64_t s0 = c * U / G, s1 = (c + 1) * U / G;       // units of CTA c
    const int64_t lo = ua > s0 ? ua : s0, hi = ub < s1 ? ub : s1;
    for (int g = 0; g < NG; ++g) {
      // first unit >= lo with (u - s0) % NG == g
      int64_t first = lo + ((g - (lo - s0)) % NG + NG) % NG;
      if (first < hi) v += db_part[((c + t) * NG + g) * 128 + threadIdx.x];
    }
  }
  out[(int64_t)arm * out_arm_stride + gene] = v;
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
int launch_f11(const CUtensorMap& tmX, const CUtensorMap& tmTk, const CUtensorMap& tmTm, F11Args& a, int64_t* U_out,
               int64_t* G_out, cudaStream_t s) {
  const int64_t U = (int64_t)a.batch * a.rtiles * a.ktiles;
  int64_t G = sm_count2();
  if (G > U) G = U;
  a.nk = 4; a.nm = 4;
  a.nx = (227 * 1024 - 2048 - (a.nk + a.nm) * IMG_BYTES) / X_BYTES;
  if (a.nx > 8) a.nx = 8;
  const size_t smem = (size_t)a.nx * X_BYTES + (size_t)(a.nk + a.nm) * IMG_BYTES + (2 * a.nx + 2 * a.nk + 2 * a.nm + 5 * NG + 4) * 8 + 1024;
  static bool attr = false;
  if (!attr) {
    MVAE_CUDA(cudaFuncSetAttribute(fc11_ts_kernel<GENE, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  fc11_ts_kernel<GENE, TRAIN><<<dim3((unsigned)G), THREADS, smem, s>>>(tmX, tmTk, tmTm, a);
  MVAE_LAUNCH_CHECK();
  *U_out = U; *G_out = G;
  return 0;
}

}  // namespace

// x_hat / loss sums / d h10 in one pass over x (x_rec optional: materialised reconstruction)
int ts_fc11_rows(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale, int want_grad,
                 float* x_rec, double* recon_acc, cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  float* work = st.work;
  F11Args a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.D = D; a.H = H; a.HN = (H + 15) / 16 * 16;
  a.batch = A; a.rtiles = (B + 127) / 128; a.ktiles = (D + UN - 1) / UN;
  a.x_batched = in.x_arm_stride > 0;
  a.want_grad = want_grad; a.gscale = gscale;
  a.R = work + w.d[4]; a.r_arm_stride = (int64_t)B * H; a.r_rows = B;
  a.bias = st.params + L.offset[FC11_B]; a.bias_arm_stride = L.arm_stride;
  a.x_rec = x_rec; a.xrec_arm_stride = (int64_t)B * D;
  a.recon_acc = recon_acc;
  a.part = work + w.fc1_part;
  CUtensorMap tmX, tmTk, tmTm;
  int rc = make_map_ex(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, 32, 128, 1);
  if (rc) return rc;
  rc = make_map_ex(&tmTk, st.params + L.offset[FC11_W], H, D, H, A, L.arm_stride, 32, UN, 1);
  if (rc) return rc;
  rc = make_map_ex(&tmTm, st.params + L.offset[FC11_W], H, D, H, A, L.arm_stride, 32, UN, 2);
  if (rc) return rc;
  int64_t U, G;
  rc = (want_grad && !x_rec) ? launch_f11<false, true>(tmX, tmTk, tmTm, a, &U, &G, s)
                             : launch_f11<false, false>(tmX, tmTk, tmTm, a, &U, &G, s);
  if (rc || !want_grad) return rc;
  f11_fixup_kernel<<<dim3((B + 127) / 128 * 8, A), 256, 0, s>>>(a.part, A, a.ktiles, U, G, work + w.g_d10, (int64_t)B * H, H, B, H);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// d fc11.weight and d fc11.bias by the gene pass (x_hat recomputed)
int ts_fc11_genes(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale, cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  float* work = st.work;
  F11Args a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.D = D; a.H = H; a.HN = (H + 15) / 16 * 16;
  a.batch = A; a.rtiles = (D + 127) / 128; a.ktiles = (B + UN - 1) / UN;
  a.x_batched = in.x_arm_stride > 0;
  a.want_grad = 1; a.gscale = gscale;
  a.R = st.params + L.offset[FC11_W]; a.r_arm_stride = L.arm_stride; a.r_rows = D;
  a.bias = st.params + L.offset[FC11_B]; a.bias_arm_stride = L.arm_stride;
  a.part = work + w.fc1_part;
  a.db_part = work + w.db_part;
  CUtensorMap tmX, tmTk, tmTm;
  int rc = make_map_ex(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, 128, 32, 0);
  if (rc) return rc;
  rc = make_map_ex(&tmTk, work + w.d[4], H, B, H, A, (int64_t)B * H, 32, UN, 1);
  if (rc) return rc;
  rc = make_map_ex(&tmTm, work + w.d[4], H, B, H, A, (int64_t)B * H, 32, UN, 2);
  if (rc) return rc;
  int64_t U, G;
  rc = launch_f11<true, true>(tmX, tmTk, tmTm, a, &U, &G, s);
  if (rc) return rc;
  f11_fixup_kernel<<<dim3((D + 127) / 128 * 8, A), 256, 0, s>>>(a.part, A, a.ktiles, U, G, st.grads + L.offset[FC11_W], L.arm_stride,
                                                           H, D, H);
  MVAE_LAUNCH_CHECK();
  f11_db_fixup_kernel<<<dim3((D + 127) / 128, A), 128, 0, s>>>(a.db_part, A, a.ktiles, U, G, st.grads + L.offset[FC11_B],
                                                             L.arm_stride, D);
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae
