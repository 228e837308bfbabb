// ts_gemm.cu — the two passes over the gene matrix x that belong to fc1 (mmidas/nn_model.py:264):
//
//   FWD    a1_pre[cell][h]   = sum_gene dropout(x)[cell][gene] * W1[h][gene]      (3xTF32 or TF32)
//   WGRAD  dW1^T[gene][h]    = sum_cell dropout(x)[cell][gene] * delta1[cell][h]  (TF32)
//
// Both are "x-streaming" GEMMs with a 128-wide output: HBM-bound if the per-element work on x (dropout
// mask, TF32 hi/lo split) keeps up.  Design (sm_100a):
//   * x tiles arrive by TMA in a deep ring of raw 16 KB tiles (they come from HBM: long latency);
//     the small operand (W1 / delta1, L2-resident) has its own ring.
//   * 16 transform warps read the raw tile ONCE from shared memory, apply the dropout mask and the hi/lo
//     split in registers and write the MMA's A operand straight into TENSOR MEMORY (tcgen05.st); the
//     MMAs are issued in the TS form (A from TMEM, B from shared memory).  No transformed tile is ever
//     written back to shared memory, the raw slot is released as soon as it has been read, and for WGRAD the
//     transposition of x (gene-major A operand) happens for free in the register -> TMEM step.
//   * stream-K: the (tile, k-tile) units are cut into gridDim.x equal contiguous ranges (one CTA per SM,
//     one wave, no tail); a CTA writes one partial tile per output tile it touches and a fix-up kernel
//     sums the partials of each tile in a fixed order (deterministic) and applies the layer epilogue.
//   * W1_lo = W1 - tf32(W1) is computed once per step by a tiny kernel, so the W operand needs no transform.
//   * The kernels are bound by L2 -> SM bytes (x tile + small-operand tile per unit, ~6 TB/s on the chip), so an
//     output tile is 256 rows = two 128-row blocks that share every small-operand stage (two accumulators).
#include "gemm_tc.h"
#include "tc_common.cuh"

namespace mvae {

namespace {
using namespace tc;

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int NG = 4;                      // transform groups: group g owns the units i = g (mod NG) of its CTA
constexpr int NTW = 4 * NG;                // transform / drain warps: 4 per group (one per TMEM lane quadrant)
constexpr int CTRL_WARPS = 3;              // warp 0: TMA of x, warp 1: TMA of the small operand, warp 2: MMA issue
constexpr int THREADS = 32 * (CTRL_WARPS + NTW);
constexpr int NA = NG;                     // A-operand stages in tensor memory (64 columns each: hi | lo), one per group
constexpr int X_BYTES = 16384;
constexpr int TILE_FLOATS = 128 * 128;     // one partial tile
constexpr int NB = 2;                      // 128-row blocks per output tile (they share the small-operand stages)
constexpr uint32_t ACC_COLS = 128 * NB;    // one accumulator per block: [0,128), [128,256); the A stages follow

struct TsArgs {
  int BN;                       // UMMA N (multiple of 16, <= 128)
  int batch, mtiles, ktiles;    // arms, 256-row tiles per arm, 32-deep k tiles
  int x_batched;                // x has an arm coordinate
  int nx, nw;                   // ring depths
  int w_tile_bytes;             // bytes of one W tile image
  int split3;                   // 3xTF32 (FWD only)
  float* part;                  // partial tiles [slot][128][128]
  DropSpec drop;
};

__host__ __device__ inline int64_t cta_of_unit(int64_t u, int64_t U, int64_t G) { return ((u + 1) * G - 1) / U; }

// NXT / NWT: compile-time ring depths (0: run-time values of a.nx / a.nw) -- with constants every ring slot and mbarrier
// address is the shared base plus an immediate
template <bool WGRAD, int NXT, int NWT>
__global__ void __launch_bounds__(THREADS, 1)
ts_gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmWlo, const TsArgs a) {
  const int NX = NXT ? NXT : a.nx, NW = NWT ? NWT : a.nw;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer stays derived from smem_raw, so the compiler keeps the
  // shared address space (LDS / direct mbarrier addresses instead of generic loads and 64-bit window arithmetic)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w_stage_bytes = a.w_tile_bytes * (a.split3 ? 2 : 1);
  auto xs = [&](int s) { return smem + (size_t)s * X_BYTES; };
  auto ws = [&](int s) { return smem + (size_t)NX * X_BYTES + (size_t)s * w_stage_bytes; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NX * X_BYTES + (size_t)NW * w_stage_bytes);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + NX;
  uint64_t* w_full = x_empty + NX;
  uint64_t* w_empty = w_full + NW;
  uint64_t* a_full = w_empty + NW;
  uint64_t* a_empty = a_full + NA;
  // MMAs of segment s complete (both blocks): barrier s % NA.  A transform group can be up to NA units -- hence up to
  // NA segments when a CTA's share of a tile is a single unit -- ahead of the tensor pipe, which one parity bit
  // cannot tell apart; NA barriers can.
  uint64_t* acc_full = a_empty + NA;
  uint64_t* acc_empty = acc_full + NA;    // both accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int KT = a.ktiles;
  const int64_t U = (int64_t)a.batch * a.mtiles * KT, G = gridDim.x;
  const int64_t u0 = (int64_t)blockIdx.x * U / G, u1 = ((int64_t)blockIdx.x + 1) * U / G;
  const int nu = (int)(u1 - u0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NX; ++s) { mbar_init(x_full + s, 1); mbar_init(x_empty + s, 4); }
    for (int s = 0; s < NW; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
    for (int s = 0; s < NA; ++s) { mbar_init(a_full + s, 4); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < NA; ++s) mbar_init(acc_full + s, 1);
    mbar_init(acc_empty, 4 * NB);
    fence_barrier_init();
  }
  if (warp == CTRL_WARPS) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the set-up above and the first fills of the x ring (x is an input of the step) overlap the tail of the previous
  // kernel; everything else -- the small operand, the generator keys, the partial tiles -- waits for it
  pdl_trigger();
  if (warp != 0) pdl_wait();

  // first unit of this CTA; every role then advances (kt, arm, mt) and its ring slots incrementally (no divisions
  // inside the loops: the transform warps are instruction-issue bound)
  const int t_first = (int)(u0 / KT), kt_first = (int)(u0 - (int64_t)t_first * KT);
  const int mt_first = t_first / a.batch, arm_first = t_first - mt_first * a.batch;

  if (warp == 0) {
    // ===== TMA producer of the raw x tiles (HBM): runs ahead by the whole x ring, independent of the W ring =====
    // (the whole warp runs the uniform loop; one elected lane issues, so the TMA operands stay in uniform registers)
    int kt = kt_first, mt = mt_first, arm = arm_first, sx = 0;
    uint32_t phx = 1;
    for (int i = 0; i < nu; ++i) {
      const int xb = a.x_batched ? arm : 0;
#pragma unroll
      for (int b = 0; b < NB; ++b) {                 // x-unit j = NB * i + b: block b of the tile
        const int m0 = (NB * mt + b) * BM;
        mbar_wait(x_empty + sx, phx);
        if (elect_one()) {
          mbar_expect_tx(x_full + sx, X_BYTES);
          if (!WGRAD) tma_load_3d(&tmX, x_full + sx, xs(sx), kt * BK, m0, xb);         // [128 cells][32 genes], SW128
          else tma_load_3d(&tmX, x_full + sx, xs(sx), m0, kt * BK, xb);                // [32 cells][128 genes], linear
        }
        __syncwarp();
        if (++sx == NX) { sx = 0; phx ^= 1; }
      }
      if (++kt == KT) { kt = 0; if (++arm == a.batch) { arm = 0; ++mt; } }
    }
  } else if (warp == 1) {
    // ===== TMA producer of the small operand (W1 hi/lo or delta1; L2-resident) =====
    int kt = kt_first, arm = arm_first, sw = 0;
    uint32_t phw = 1;
    for (int i = 0; i < nu; ++i) {
      mbar_wait(w_empty + sw, phw);
      if (elect_one()) {
        mbar_expect_tx(w_full + sw, w_stage_bytes);
        if (!WGRAD) {
          tma_load_3d(&tmW, w_full + sw, ws(sw), kt * BK, 0, arm);
          if (a.split3) tma_load_3d(&tmWlo, w_full + sw, ws(sw) + a.w_tile_bytes, kt * BK, 0, arm);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(&tmW, w_full + sw, ws(sw) + j * 4096, 32 * j, kt * BK, arm);
        }
      }
      __syncwarp();
      if (++sw == NW) { sw = 0; phw ^= 1; }
      if (++kt == KT) { kt = 0; if (++arm == a.batch) arm = 0; }
    }
  } else if (warp == 2) {
    // ===== MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues =====
    const uint32_t idesc = make_idesc(BM, a.BN, false, WGRAD);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t acc = 0;
    int kt = kt_first, sw = 0, seg = 0;
    uint32_t phw = 0;
    for (int i = 0; i < nu; ++i) {
      if (kt == 0 || i == 0) {                       // a new segment: the accumulators must have been drained
        acc = 0;
        mbar_wait(acc_empty, (seg & 1) ^ 1);
      }
      // the three barriers of a unit are waited for together (a satisfied try_wait still costs ~150 cycles of latency)
      {
        const int j0 = NB * i, j1 = NB * i + 1;
        const bool r0 = mbar_try_wait(w_full + sw, phw);
        const bool r1 = mbar_try_wait(a_full + (j0 & (NA - 1)), (j0 / NA) & 1);
        const bool r2 = mbar_try_wait(a_full + (j1 & (NA - 1)), (j1 / NA) & 1);
        if (!r0) mbar_wait(w_full + sw, phw);
        if (!r1) mbar_wait(a_full + (j0 & (NA - 1)), (j0 / NA) & 1);
        if (!r2) mbar_wait(a_full + (j1 & (NA - 1)), (j1 / NA) & 1);
      }
      tc_fence_after();
      const uint32_t wb = smem_u32(ws(sw));
      if (elect_one()) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int j = NB * i + b, sa = j & (NA - 1);
          const uint32_t dcol = tb + (uint32_t)b * 128u;
          const uint32_t a_hi = tb + ACC_COLS + (uint32_t)sa * 64u, a_lo = a_hi + 32u;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint32_t accf = (acc | (uint32_t)ks) ? 1u : 0u;
            if (!WGRAD) {
              const uint64_t bh = make_smem_desc(wb + ks * 32, 0, 1024, false);
              if (a.split3) {
                umma_tf32_ts(dcol, a_lo + ks * 8, bh, idesc, accf);
                umma_tf32_ts(dcol, a_hi + ks * 8, make_smem_desc(wb + a.w_tile_bytes + ks * 32, 0, 1024, false), idesc, 1u);
                umma_tf32_ts(dcol, a_hi + ks * 8, bh, idesc, 1u);
              } else {
                umma_tf32_ts(dcol, a_hi + ks * 8, bh, idesc, accf);
              }
            } else {
              umma_tf32_ts(dcol, a_hi + ks * 8, make_smem_desc(wb + ks * 1024, 4096, 512, true), idesc, accf);
            }
          }
          umma_commit(a_empty + sa);
        }
      }
      __syncwarp();
      acc = 1;
      const bool seg_end = (kt + 1 == KT) || (i == nu - 1);
      if (elect_one()) {
        umma_commit(w_empty + sw);
        if (seg_end) umma_commit(acc_full + (seg & (NA - 1)));          // this CTA's share of the tile is complete
      }
      __syncwarp();
      if (++sw == NW) { sw = 0; phw ^= 1; }
      if (++kt == KT) kt = 0;
      if (seg_end) ++seg;
    }
  } else {
    // ===== transform warps: raw x tile (smem) -> masked / split A operand (TMEM); the group that transforms the last
    // unit of a segment also drains its accumulator.  Group g works on units g, g + NG, ...: NG units are in flight,
    // which hides the per-unit latency chain (barrier -> LDS -> hash -> STTM -> fence -> barrier).
    const int quad = warp & 3, grp = (warp - CTRL_WARPS) >> 2;
    const int r = quad * 32 + lane;                       // TMEM lane: cell (FWD) or gene (WGRAD) within the tile
    const uint32_t lane_bits = (uint32_t)(quad * 32) << 16;
    const DropSpec& dp = a.drop;
    const uint64_t Dq = (uint64_t)dp.D >> 2;
    const uint32_t thr = dp.thresh16;
    const uint32_t sh = 8u * (uint32_t)(lane & 3);
    const uint32_t acol = tmem_base + lane_bits + ACC_COLS + (uint32_t)grp * 64u;      // this group's A stage
    // group g owns the x-units j = g (mod NG), j = NB * i + b: block b = g % NB of the units i = g / NB (mod NG / NB)
    const int blk = grp % NB;
    constexpr int ISTEP = NG / NB;
    int kt = kt_first + grp / NB, t = t_first, mt = mt_first, arm = arm_first;
    while (kt >= KT) { kt -= KT; ++t; if (++arm == a.batch) { arm = 0; ++mt; } }
    int sx = grp % NX;
    uint32_t phx = (uint32_t)((grp / NX) & 1);
    uint32_t pha = 1;                                     // parity for a_empty (first use passes)
    uint64_t dkey = 0;                                    // generator key of (seed, step, global arm): re-read only when the arm changes
    int key_arm = -1;                                     // (a load per half-unit sat in the middle of the group's latency chain)
    for (int i = grp / NB; i < nu; i += ISTEP) {
      const int m0 = (NB * mt + blk) * BM;                // first row (FWD: cell, WGRAD: gene) of this block
      if (dp.mode == 2 && arm != key_arm) {
        dkey = dp.keys[arm];
        key_arm = arm;
      }
      mbar_wait(x_full + sx, phx);
      const uint8_t* tile = xs(sx);
      bool waited = false;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[16];
        if (!WGRAD) {
          // thread = cell r, genes kt*32 + 16*half .. +15: 16-byte chunks (SWIZZLE_128B: chunk c lives at c ^ (r & 7))
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 t4 = *reinterpret_cast<const uint4*>(tile + r * 128 + (((4 * half + q) ^ (r & 7)) << 4));
            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
          }
        } else {
          // thread = gene r, cells kt*32 + 16*half .. +15: column reads of the linear [32 cells][128 genes] tile
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const uint32_t*>(tile + (16 * half + j) * 512 + r * 4);
        }
        // ---- dropout: zero the dropped elements (the 1/(1-p) scale is applied by the fix-up kernel)
        if (dp.mode == 2) {
          if (!WGRAD) {
            const uint64_t chunk = (uint64_t)(m0 + r) * Dq + (uint64_t)(kt * 8 + 4 * half);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t b = drop_bits4(dkey, chunk + q);
#pragma unroll
              for (int j = 0; j < 4; ++j) v[4 * q + j] = ((b >> (8 * j)) & 0xFFu) >= thr ? v[4 * q + j] : 0u;
            }
          } else {
            // the warp's 16 cells x 32 genes are 128 generator chunks: 4 per lane, shared by shuffles
            const uint64_t cell0 = (uint64_t)(kt * BK + 16 * half + (lane >> 3));
            const uint64_t ccol = (uint64_t)((m0 + quad * 32) >> 2) + (uint64_t)(lane & 7);
            uint32_t hq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) hq[q] = drop_bits4(dkey, (cell0 + 4 * q) * Dq + ccol);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t h = __shfl_sync(0xffffffffu, hq[j >> 2], (j & 3) * 8 + (lane >> 2));
              v[j] = ((h >> sh) & 0xFFu) >= thr ? v[j] : 0u;
            }
          }
        } else if (dp.mode == 1) {
          if (!WGRAD) {
            const int64_t xrow = m0 + r, xcol = (int64_t)kt * BK + 16 * half;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t kw = 0;
              if (xrow < dp.rows && xcol + 4 * q < dp.D)
                kw = *reinterpret_cast<const uint32_t*>(dp.keep + (int64_t)arm * dp.keep_arm_stride + xrow * dp.D + xcol + 4 * q);
#pragma unroll
              for (int j = 0; j < 4; ++j) v[4 * q + j] = ((kw >> (8 * j)) & 0xFFu) ? v[4 * q + j] : 0u;
            }
          } else {
            const int64_t gene = m0 + r;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int64_t cell = (int64_t)kt * BK + 16 * half + j;
              uint8_t kb = 0;
              if (gene < dp.D && cell < dp.rows) kb = dp.keep[(int64_t)arm * dp.keep_arm_stride + cell * dp.D + gene];
              v[j] = kb ? v[j] : 0u;
            }
          }
        }
        if (half == 1 && dp.mode == 2) {
          // every value read from the raw slot has been consumed by the mask selects above (register dependencies), so
          // the shared-memory reads have been performed: hand the slot back before the TMEM stores
          __syncwarp();
          if (lane == 0) mbar_arrive(x_empty + sx);
        }
        if (!waited) {
          mbar_wait(a_empty + grp, pha);                   // the MMAs that read this TMEM stage have completed
          tc_fence_after();
          waited = true;
        }
        tmem_st8(acol + 16u * half, v);                    // "hi": the tensor core truncates to TF32 itself
        tmem_st8(acol + 16u * half + 8u, v + 8);
        if (!WGRAD && a.split3) {
          uint32_t lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) lo[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(v[j] & 0xFFFFE000u));
          tmem_st8(acol + 32u + 16u * half, lo);
          tmem_st8(acol + 32u + 16u * half + 8u, lo + 8);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      // the raw slot goes back only now: every value read from it has been consumed (a register dependency), so the
      // shared-memory reads have certainly been performed before the TMA may overwrite the slot
      if (lane == 0) {
        if (dp.mode != 2) mbar_arrive(x_empty + sx);       // (no consuming instruction before the stores in these modes)
        mbar_arrive(a_full + grp);
      }
      pha ^= 1;
      if (kt == KT - 1 || i == nu - 1) {
        // ---- drain this CTA's share of tile t into its partial slot (slot = cta + tile: unique, monotone)
        const int seg = t - t_first;
        mbar_wait(acc_full + (seg & (NA - 1)), (seg / NA) & 1);
        tc_fence_after();
        float* prt = a.part + (((int64_t)blockIdx.x + t) * NB + blk) * TILE_FLOATS;
        const uint32_t dcol = tmem_base + lane_bits + (uint32_t)blk * 128u;
        for (int j = 0; j < a.BN / 16; ++j) {
          uint32_t rr[16];
          tmem_ld16(dcol + (uint32_t)(j * 16), rr);
          tmem_ld_wait();
          if (!WGRAD) {
            float4* dst = reinterpret_cast<float4*>(prt + r * 128 + j * 16);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              dst[e] = make_float4(__uint_as_float(rr[4 * e]), __uint_as_float(rr[4 * e + 1]), __uint_as_float(rr[4 * e + 2]),
                                   __uint_as_float(rr[4 * e + 3]));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) prt[(j * 16 + e) * 128 + r] = __uint_as_float(rr[e]);   // [h][gene]: coalesced
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
      // advance to x-unit j + NG (unit i + NG / NB)
      sx += NG;
      if (sx >= NX) { sx -= NX; phx ^= 1; }
      kt += ISTEP;
      while (kt >= KT) { kt -= KT; ++t; if (++arm == a.batch) { arm = 0; ++mt; } }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CTRL_WARPS) tmem_dealloc(tmem_base, 512);
}

// W1_lo = W1 - tf32(W1) for every arm (the part of fc1.weight the tensor core drops when it truncates to TF32)
__global__ void __launch_bounds__(256) w_lo_kernel(const float* __restrict__ w, int64_t w_arm_stride, float* __restrict__ lo,
                                                   int64_t lo_arm_stride, int64_t n4) {
  const int arm = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(w + (int64_t)arm * w_arm_stride);
  float4* dst = reinterpret_cast<float4*>(lo + (int64_t)arm * lo_arm_stride);
  pdl_trigger();
  pdl_wait();      // (the scratch buffer W1_lo lives in is read by the previous step's last kernels)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}

// fc1 fix-up: a1 = relu(sum of the tile's partials (fixed order) + b1), fp64 column sums for batch_l1.
struct Fc1FixArgs {
  const float* part; int batch, ktiles; int64_t U, G;
  float scale;                                   // 1/(1-p) of the input dropout (the GEMM ran on the unscaled mask)
  const float* params; int64_t p_arm_stride, offB;
  float* out; double* stats_out; int B, H;
};
__global__ void __launch_bounds__(256) fc1_fixup_kernel(const Fc1FixArgs p) {
  // block = 16 rows of a 128-row block of an output tile; thread = (row mod 8, float4 column): 2 rows each, 512-byte rows
  __shared__ double red[8][2][128];
  __shared__ int cc[2];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, r8 = tid >> 5, c4 = tid & 31;
  const int row0 = (blockIdx.x >> 3) * 128, sub0 = (blockIdx.x & 7) * 16;
  const int64_t t = (int64_t)(row0 >> 8) * p.batch + arm;
  pdl_trigger();
  pdl_wait();
  if (tid == 0) {
    cc[0] = (int)cta_of_unit(t * p.ktiles, p.U, p.G);
    cc[1] = (int)cta_of_unit(t * p.ktiles + p.ktiles - 1, p.U, p.G);
  }
  __syncthreads();
  const int c0 = cc[0], c1 = cc[1];
  const bool col_ok = 4 * c4 < p.H;                       // H % 4 == 0 on this path
  float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col_ok) bj = *reinterpret_cast<const float4*>(p.params + (int64_t)arm * p.p_arm_stride + p.offB + 4 * c4);
  float* out = p.out + (int64_t)arm * p.B * p.H;
  const float* base = p.part + ((int64_t)(c0 + t) * NB + ((row0 >> 7) & 1)) * TILE_FLOATS + 4 * c4;
  double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  if (col_ok) {
    for (int k = 0; k < 2; ++k) {
      const int rl = sub0 + r8 + 8 * k, row = row0 + rl;
      if (row >= p.B) break;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = c0; c <= c1; ++c) {
        const float4 q = *reinterpret_cast<const float4*>(base + (int64_t)(c - c0) * NB * TILE_FLOATS + rl * 128);
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      v.x = fmaxf(fmaf(v.x, p.scale, bj.x), 0.f); v.y = fmaxf(fmaf(v.y, p.scale, bj.y), 0.f);
      v.z = fmaxf(fmaf(v.z, p.scale, bj.z), 0.f); v.w = fmaxf(fmaf(v.w, p.scale, bj.w), 0.f);
      *reinterpret_cast<float4*>(out + (int64_t)row * p.H + 4 * c4) = v;
      s1[0] += (double)v.x; s1[1] += (double)v.y; s1[2] += (double)v.z; s1[3] += (double)v.w;
      s2[0] += (double)v.x * (double)v.x; s2[1] += (double)v.y * (double)v.y;
      s2[2] += (double)v.z * (double)v.z; s2[3] += (double)v.w * (double)v.w;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    red[r8][0][4 * c4 + e] = s1[e];
    red[r8][1][4 * c4 + e] = s2[e];
  }
  __syncthreads();
  {
    const int which = tid >> 7, j = tid & 127;
    if (j < p.H) {
      double sacc = 0.0;
      for (int w = 0; w < 8; ++w) sacc += red[w][which][j];
      atomicAdd(p.stats_out + (int64_t)arm * 256 + which * 128 + j, sacc);
    }
  }
}

// d fc1.weight fix-up: dW1[arm][h][gene] = sum of the partials [slot][h][gene in tile] in a fixed order
__global__ void __launch_bounds__(256) wgrad_fixup_kernel(const float* part, int batch, int ktiles, int64_t U, int64_t G,
                                                          float scale, float* grads, int64_t g_arm_stride, int D, int H) {
  // block = one 128-gene block (partials are [h][gene]); thread = (h mod 8, float4 of genes)
  __shared__ int cc[2];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, h8 = tid >> 5, g4 = tid & 31;
  const int gb = blockIdx.x >> 2, hq = blockIdx.x & 3;     // 128-gene block, quarter of the h rows
  const int gene = gb * 128 + 4 * g4;
  const int64_t t = (int64_t)(gb >> 1) * batch + arm;
  pdl_trigger();
  pdl_wait();
  if (tid == 0) {
    cc[0] = (int)cta_of_unit(t * ktiles, U, G);
    cc[1] = (int)cta_of_unit(t * ktiles + ktiles - 1, U, G);
  }
  __syncthreads();
  if (gene >= D) return;                                  // D % 4 == 0 on this path
  const int c0 = cc[0], c1 = cc[1];
  const float* base = part + ((int64_t)(c0 + t) * NB + (gb & 1)) * TILE_FLOATS + 4 * g4;
  float* out = grads + (int64_t)arm * g_arm_stride + gene;
  for (int h = h8 + 8 * hq; h < H; h += 32) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c0; c <= c1; ++c) {
      const float4 q = *reinterpret_cast<const float4*>(base + (int64_t)(c - c0) * NB * TILE_FLOATS + h * 128);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    *reinterpret_cast<float4*>(out + (int64_t)h * D) = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <bool WGRAD>
int launch_ts(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmWlo, TsArgs& a, int64_t* U_out,
              int64_t* G_out, cudaStream_t s) {
  const int64_t U = (int64_t)a.batch * a.mtiles * a.ktiles;
  int64_t G = sm_count();
  if (G > U) G = U;
  const int w_stage = a.w_tile_bytes * (a.split3 ? 2 : 1);
  // rings: the small operand gets 4-5 stages, the raw x ring the rest of the 227 KB
  // The x ring depth must be a multiple of NG: then a slot is always refilled for the group that emptied it.  With any
  // other depth a slot alternates between two groups, and a group can reach "its" fill k while fill k - 1 -- the other
  // group's, TMA completions are not ordered -- is still in flight: the parity wait for fill k then succeeds on the
  // completed fill k - 2 and the protocol falls apart (stale tile, early release, two fills pending on one barrier:
  // seen as intermittent launch failures when x rows are not 128-byte aligned).
  a.nw = a.split3 ? 4 : 5;
  auto nx_for = [&](int nw) { return (227 * 1024 - 1024 - 512 - nw * w_stage) / X_BYTES; };
#ifdef TS_NX
  a.nx = TS_NX;
#else
  if (nx_for(a.nw) < 2 * NG && nx_for(a.nw - 1) >= 2 * NG) --a.nw;
  a.nx = nx_for(a.nw) >= 2 * NG ? 2 * NG : NG;
#endif
  const size_t smem = (size_t)a.nx * X_BYTES + (size_t)a.nw * w_stage + (2 * a.nx + 2 * a.nw + 3 * NA + 6) * 8 + 1024;
  static bool attr[64] = {};
  if (first_on_device(attr)) {
    MVAE_CUDA(cudaFuncSetAttribute(ts_gemm_kernel<WGRAD, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MVAE_CUDA(cudaFuncSetAttribute(ts_gemm_kernel<WGRAD, 8, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MVAE_CUDA(cudaFuncSetAttribute(ts_gemm_kernel<WGRAD, 8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  // (the depths of the reference's fc_dim = 100: 8 + 3 for the 3xTF32 forward, 8 + 5 otherwise)
  if (a.nx == 8 && a.nw == 5) launch_pdl(ts_gemm_kernel<WGRAD, 8, 5>, dim3((unsigned)G), dim3(THREADS), smem, s, tmX, tmW, tmWlo, a);
  else if (a.nx == 8 && a.nw == 3) launch_pdl(ts_gemm_kernel<WGRAD, 8, 3>, dim3((unsigned)G), dim3(THREADS), smem, s, tmX, tmW, tmWlo, a);
  else launch_pdl(ts_gemm_kernel<WGRAD, 0, 0>, dim3((unsigned)G), dim3(THREADS), smem, s, tmX, tmW, tmWlo, a);
  MVAE_LAUNCH_CHECK();
  *U_out = U; *G_out = G;
  return 0;
}

}  // namespace

int64_t ts_part_floats(int A, int Bpad, int Dpad) {
  const int64_t t1 = (int64_t)A * (Bpad / 128), t2 = (int64_t)A * (Dpad / 128);
  return ((t1 > t2 ? t1 : t2) + 2 * A + 320) * TILE_FLOATS;
}

// fc1 forward: a1 = relu(dropout(x) . W1^T + b1) and the batch_l1 column sums (replaces GEMM + epilogue kernel)
int ts_fc1_forward(const mvae_dims& d, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                   const DropSpec& drop, const Work& w, float* a1_out, double* stats_out, Fc1Deferred* defer, cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  const int split3 = hp.precision != 2;
  float* wlo = st.work + w.w11_t;                       // [A][128 * Dpad] scratch, pitch D
  const int64_t wlo_stride = (int64_t)128 * w.Dpad;
  if (split3) {
    const int64_t n4 = (int64_t)H * D / 4;
    launch_pdl(w_lo_kernel, dim3(148, A), dim3(256), 0, s, (const float*)(st.params + L.offset[FC1_W]), L.arm_stride, wlo, wlo_stride, n4);
    MVAE_LAUNCH_CHECK();
  }
  TsArgs a;
  memset(&a, 0, sizeof(a));
  a.BN = (H + 15) / 16 * 16;
  a.batch = A; a.mtiles = (B + NB * BM - 1) / (NB * BM); a.ktiles = (D + BK - 1) / BK;
  a.x_batched = in.x_arm_stride > 0;
  a.w_tile_bytes = a.BN * 128;
  a.split3 = split3;
  a.part = st.work + w.fc1_part;
  a.drop = drop;
  CUtensorMap tmX, tmW, tmWlo;
  int rc = make_map_ex(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, 32, 128, 1);
  if (rc) return rc;
  rc = make_map_ex(&tmW, st.params + L.offset[FC1_W], D, H, D, A, L.arm_stride, 32, a.BN, 1);
  if (rc) return rc;
  rc = make_map_ex(&tmWlo, split3 ? wlo : st.params + L.offset[FC1_W], D, H, D, A, split3 ? wlo_stride : L.arm_stride, 32,
                   a.BN, 1);
  if (rc) return rc;
  int64_t U, G;
  rc = launch_ts<false>(tmX, tmW, tmWlo, a, &U, &G, s);
  if (rc) return rc;
  if (defer) {          // the encoder chain sums the partials itself (kernels_chain.cu)
    defer->part = a.part; defer->batch = A; defer->ktiles = a.ktiles; defer->U = U; defer->G = G;
    defer->scale = drop.mode ? drop.scale : 1.f; defer->offB = L.offset[FC1_B]; defer->valid = 1;
    return 0;
  }
  Fc1FixArgs f;
  memset(&f, 0, sizeof(f));
  f.part = a.part; f.batch = A; f.ktiles = a.ktiles; f.U = U; f.G = G;
  f.scale = drop.mode ? drop.scale : 1.f;
  f.params = st.params; f.p_arm_stride = L.arm_stride; f.offB = L.offset[FC1_B];
  f.out = a1_out; f.stats_out = stats_out; f.B = B; f.H = H;
  launch_pdl(fc1_fixup_kernel, dim3((B + 127) / 128 * 8, A), dim3(256), 0, s, f);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// d fc1.weight = delta1^T . dropout(x), TF32
int ts_fc1_wgrad(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const DropSpec& drop, const Work& w,
                 cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  TsArgs a;
  memset(&a, 0, sizeof(a));
  a.BN = (H + 15) / 16 * 16;
  a.batch = A; a.mtiles = (D + NB * BM - 1) / (NB * BM); a.ktiles = (B + BK - 1) / BK;
  a.x_batched = in.x_arm_stride > 0;
  a.w_tile_bytes = 16384;
  a.split3 = 0;
  a.part = st.work + w.fc1_part;
  a.drop = drop;
  CUtensorMap tmX, tmW;
  int rc = make_map_ex(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, 128, 32, 0);
  if (rc) return rc;
  rc = make_map_ex(&tmW, st.work + w.delta_enc[0], H, B, H, A, (int64_t)B * H, 32, 32, 2);
  if (rc) return rc;
  int64_t U, G;
  rc = launch_ts<true>(tmX, tmW, tmW, a, &U, &G, s);
  if (rc) return rc;
  launch_pdl(wgrad_fixup_kernel, dim3((D + 127) / 128 * 4, A), dim3(256), 0, s, (const float*)a.part, A, a.ktiles, U, G,
             drop.mode ? drop.scale : 1.f, st.grads + L.offset[FC1_W], L.arm_stride, D, H);
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae
