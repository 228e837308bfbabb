// gemm_tc.cu — the gene-dimension GEMMs of the cpl-mixVAE step on the 5th-generation tensor cores:
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma kind::tf32 with the
// accumulator in TMEM -> tcgen05.ld epilogue.  sm_100a only.
//
//   G1  fc1 forward        part[s] = dropout(x)[B,D] . W1[H,D]^T       (3xTF32, split-K)   nn_model.py:264
//   G2  fc11 forward       pre     = h10[B,H] . W11[D,H]^T                                 nn_model.py:287
//   G3  d h10              part[s] = dY[B,D] . W11[D,H]                (split-K)            autograd of :287
//   G4  d fc11.weight      part[s] = dY[B,D]^T . h10[B,H]              (split-K)
//   G5  d fc1.weight       dW1     = delta1[B,H]^T . dropout(x)[B,D]
//
// Every operand is read in its natural row-major layout: an operand whose reduction index is the
// contiguous one is "K-major", the others ("MN-major") use the transposing shared-memory descriptor
// of tcgen05 — no transposed copies are materialised.  Warp roles: warp 0 TMA producer, warp 1 MMA
// issuer (one elected lane), warps 2-5 epilogue (TMEM lane quadrants 2,3,0,1), warps 6-9 operand
// transform (dropout mask on the x operand and the hi/lo split of the error-compensated 3xTF32 scheme:
// a = a_hi + a_lo with a_hi = a truncated to TF32 by the tensor core itself, a_lo = a - trunc(a);
// a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, fp32 accumulate).
#include <stdlib.h>

#include "gemm_tc.h"
#include "tc_common.cuh"

namespace mvae {

namespace {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 32;           // floats per pipeline stage along K (= one 128-byte swizzle row)
constexpr int UK = 8;            // K of one tcgen05.mma kind::tf32
constexpr int TILE_BYTES = 16384;  // one operand tile: 128 x 128 B (K-major) or 4 slabs of 32 x 128 B (MN-major)
constexpr int NUM_THREADS = 448;      // TMA, MMA, 4 epilogue warps, 8 transform warps
constexpr int TRANSFORM_WARP0 = 6;
constexpr int TRANSFORM_THREADS = 256;

enum : int { F_SPLIT_A = 1, F_SPLIT_B = 2, F_DROP_A = 4, F_DROP_B = 8 };

struct TcArgs {
  int M, N, K;               // logical GEMM shape (per batch entry)
  int BN;                    // UMMA N (multiple of 16, <= 128)
  int stages;
  int nsplit;                // split-K factor; grid.z = batch * nsplit
  int ktiles_per_split;      // K tiles (of BK) per split
  int a_batched, b_batched;  // operand has a batch (arm) coordinate
  int flags;
  float* C; int64_t ldc, c_batch_stride, c_split_stride;
  DropSpec drop;
  // optional fused epilogue (nsplit == 1): C = act(acc * ep_scale[col] + ep_shift[col]); act 0 none, 1 relu, 2 elu, 3 sigmoid
  const float* ep_scale; const float* ep_shift; int ep_act;
};

__device__ __forceinline__ float ep_apply(float v, float sc, float sh, int act) {
  v = fmaf(v, sc, sh);
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : expm1f(v);
  if (act == 3) return 1.f / (1.f + expf(-v));
  return v;
}

using namespace tc;

// One operand tile in shared memory, in place: apply the dropout mask (x operand) and/or write the
// low part of the TF32 split to `lo`.  The tile is the TMA image: rows of 128 bytes; K-major tiles
// (SWIZZLE_128B): 16-byte chunk p of row r holds logical chunk p ^ (r & 7); MN-major tiles
// (SWIZZLE_128B_ATOM_32B): 32-byte chunk P of row r holds logical 32-byte chunk P ^ (r & 3).
//   K-major tile : row r = tile row (M/N index), logical chunk c4 -> K offset 4*c4
//   MN-major tile: slab j (32 MN elements), row r = K offset, logical chunk c4 -> MN offset 32*j + 4*c4
template <bool MN>
__device__ __forceinline__ void transform_tile(float* hi, float* lo, bool do_split, bool do_drop, const DropSpec& drop,
                                               int arm, int mn0, int k0, int tid, int nthr) {
#pragma unroll 4
  for (int q = tid; q < TILE_BYTES / 16; q += nthr) {
    float4 v = reinterpret_cast<float4*>(hi)[q];
    if (do_drop) {
      const int r = q >> 3, p = q & 7;
      int64_t xrow, xcol;
      if (!MN) {
        xrow = mn0 + r;
        xcol = k0 + 4 * (p ^ (r & 7));
      } else {
        const int slab = r >> 5, rr = r & 31;
        xrow = k0 + rr;
        xcol = mn0 + 32 * slab + 4 * (((((p >> 1) ^ (rr & 3)) << 1)) | (p & 1));   // 32-byte-chunk swizzle
      }
      float m[4];
      if (drop.mode == 1) {
        uint32_t kb = 0;
        if (xcol < drop.D && xrow < drop.rows)   // overhanging rows/cols hold zeros from the TMA fill
          kb = *reinterpret_cast<const uint32_t*>(drop.keep + (int64_t)arm * drop.keep_arm_stride + xrow * drop.D + xcol);
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = ((kb >> (8 * i)) & 0xFF) ? drop.scale : 0.f;
      } else {
        // xcol % 4 == 0 and D % 4 == 0: the 4 elements are exactly one generator chunk
        const uint32_t bits = drop_bits4(drop.keys[arm], ((uint64_t)xrow * (uint64_t)drop.D + (uint64_t)xcol) >> 2);
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = ((bits >> (8 * i)) & 0xFFu) >= drop.thresh16 ? drop.scale : 0.f;
      }
      v.x *= m[0]; v.y *= m[1]; v.z *= m[2]; v.w *= m[3];
      reinterpret_cast<float4*>(hi)[q] = v;
    }
    if (do_split) {
      float4 l;
      l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
      reinterpret_cast<float4*>(lo)[q] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment of every tile is required by SWIZZLE_128B (descriptor base_offset = 0)
  // 1 KB alignment by an OFFSET in the shared window: the pointer stays derived from smem_raw, so the compiler keeps the
  // shared address space (LDS / direct mbarrier addresses instead of generic loads and 64-bit window arithmetic)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool splitA = args.flags & F_SPLIT_A, splitB = args.flags & F_SPLIT_B;
  const bool dropA = args.flags & F_DROP_A, dropB = args.flags & F_DROP_B;
  const bool need_transform = args.flags != 0;
  const int tiles_per_stage = 2 + (splitA ? 1 : 0) + (splitB ? 1 : 0);
  const int stage_bytes = tiles_per_stage * TILE_BYTES;
  const int S = args.stages;
  auto tileA = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto tileB = [&](int s) { return smem + (size_t)s * stage_bytes + TILE_BYTES; };
  auto tileAlo = [&](int s) { return smem + (size_t)s * stage_bytes + 2 * TILE_BYTES; };
  auto tileBlo = [&](int s) { return smem + (size_t)s * stage_bytes + (2 + (splitA ? 1 : 0)) * TILE_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* full = bars;            // TMA bytes landed
  uint64_t* ready = bars + S;       // transform done (only when need_transform)
  uint64_t* empty = bars + 2 * S;   // MMAs that read the stage have completed
  uint64_t* tmem_full = bars + 3 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);

  const int n0 = blockIdx.x * args.BN;
  const int m0 = blockIdx.y * BM;
  const int batch = blockIdx.z / args.nsplit;
  const int split = blockIdx.z - batch * args.nsplit;
  const int ktiles_total = (args.K + BK - 1) / BK;
  const int kt0 = split * args.ktiles_per_split;
  const int kt1 = min(ktiles_total, kt0 + args.ktiles_per_split);
  const int nkt = max(kt1 - kt0, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(ready + s, TRANSFORM_THREADS);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int ab = args.a_batched ? batch : 0, bb = args.b_batched ? batch : 0;
      for (int i = 0; i < nkt; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, 2 * TILE_BYTES);
        const int k0 = (kt0 + i) * BK;
        if (!A_MN) {
          tma_load_3d(&tmA, full + s, tileA(s), k0, m0, ab);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(&tmA, full + s, tileA(s) + j * 4096, m0 + 32 * j, k0, ab);
        }
        if (!B_MN) {
          tma_load_3d(&tmB, full + s, tileB(s), k0, n0, bb);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(&tmB, full + s, tileB(s) + j * 4096, n0 + 32 * j, k0, bb);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, args.BN, A_MN, B_MN);
      const int kround = (args.K + UK - 1) / UK * UK;
      uint32_t acc = 0;
      for (int i = 0; i < nkt; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(need_transform ? ready + s : full + s, ph);
        tc_fence_after();
        const int k0 = (kt0 + i) * BK;
        const uint32_t a_hi = smem_u32(tileA(s)), b_hi = smem_u32(tileB(s));
        const uint32_t a_lo = smem_u32(tileAlo(s)), b_lo = smem_u32(tileBlo(s));
#pragma unroll
        for (int ks = 0; ks < BK / UK; ++ks) {
          if (k0 + ks * UK >= kround) break;
          const uint32_t aoff = A_MN ? ks * 1024 : ks * 32;
          const uint32_t boff = B_MN ? ks * 1024 : ks * 32;
          const uint32_t albo = A_MN ? 4096 : 0, blbo = B_MN ? 4096 : 0;
          const uint32_t asbo = A_MN ? 512 : 1024, bsbo = B_MN ? 512 : 1024;
          const uint64_t dah = make_smem_desc(a_hi + aoff, albo, asbo, A_MN);
          const uint64_t dbh = make_smem_desc(b_hi + boff, blbo, bsbo, B_MN);
          if (splitA) { umma_tf32(tmem_base, make_smem_desc(a_lo + aoff, albo, asbo, A_MN), dbh, idesc, acc); acc = 1; }
          if (splitB) { umma_tf32(tmem_base, dah, make_smem_desc(b_lo + boff, blbo, bsbo, B_MN), idesc, acc); acc = 1; }
          umma_tf32(tmem_base, dah, dbh, idesc, acc);
          acc = 1;
        }
        umma_commit(empty + s);       // frees the stage once the MMAs above have read it
      }
      umma_commit(tmem_full);         // accumulator complete
    }
  } else if (warp >= TRANSFORM_WARP0) {
    // ===== operand transform (dropout mask, TF32 hi/lo split) =====
    if (need_transform) {
      const int tid = threadIdx.x - TRANSFORM_WARP0 * 32;
      for (int i = 0; i < nkt; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(full + s, ph);
        const int k0 = (kt0 + i) * BK;
        const bool doA = splitA || dropA, doB = splitB || dropB;
        if (doA && doB) {             // 4 warps per operand
          if (tid < 128)
            transform_tile<A_MN>(reinterpret_cast<float*>(tileA(s)), reinterpret_cast<float*>(tileAlo(s)), splitA, dropA,
                                 args.drop, batch, m0, k0, tid, 128);
          else
            transform_tile<B_MN>(reinterpret_cast<float*>(tileB(s)), reinterpret_cast<float*>(tileBlo(s)), splitB, dropB,
                                 args.drop, batch, n0, k0, tid - 128, 128);
        } else if (doA) {
          transform_tile<A_MN>(reinterpret_cast<float*>(tileA(s)), reinterpret_cast<float*>(tileAlo(s)), splitA, dropA,
                               args.drop, batch, m0, k0, tid, TRANSFORM_THREADS);
        } else if (doB) {
          transform_tile<B_MN>(reinterpret_cast<float*>(tileB(s)), reinterpret_cast<float*>(tileBlo(s)), splitB, dropB,
                               args.drop, batch, n0, k0, tid, TRANSFORM_THREADS);
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(ready + s);
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access
    const int row = m0 + quad * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* crow = args.C + (int64_t)batch * args.c_batch_stride + (int64_t)split * args.c_split_stride +
                  (int64_t)row * args.ldc;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
    for (int c0 = 0; c0 < args.BN; c0 += 16) {
      uint32_t r[16];
      if (nkt > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = 0u;
      }
      if (row < args.M) {
        const int col = n0 + c0;
        if (args.ep_scale) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (col + i < args.N)
              r[i] = __float_as_uint(ep_apply(__uint_as_float(r[i]), args.ep_scale[col + i], args.ep_shift[col + i], args.ep_act));
        }
        if (vec_ok && col + 16 <= args.N) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(crow + col)[i] =
                make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                            __uint_as_float(r[4 * i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (col + i < args.N) crow[col + i] = __uint_as_float(r[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------
// Grouped weight-gradient GEMM of the narrow layers (fc2..fc4, fc8..fc10: delta^T . bn(input), both operands MN-major):
// the same pipeline as tc_gemm_kernel<true, true>, one (problem, arm, K split) per CTA.  The warp-level mma.sync path
// (kernels_mma.cu) spends its issue slots on LDS.32 fragment loads and per-chunk staging (ncu: 11 M warp instructions for
// these six problems, HMMA pipe < 20 % busy); here the tensor core reads both operands from shared memory itself.
// The transform warps normalise the input tile in place ((v - mean) * rstd, batch_l1..l4), write a column of ones
// behind the last input column (row j of that accumulator column is the bias gradient sum_b delta[b][j]) and, in the
// 3xTF32 mode, the low halves.  Partials go to the layout wgrad_reduce2_kernel sums.
// ---------------------------------------------------------------------------------------------
constexpr int WGTC_MAX = 8;
struct WgTcProblem {
  int M, N;                  // nout, nin (the ones column is column N)
  int bn_layer;              // -1: raw input
  int pad;
  int64_t offW, offB;        // offsets of the partial weight / bias gradient inside one (split, arm) block
};
struct WgTcParams {
  CUtensorMap tmA[WGTC_MAX], tmB[WGTC_MAX];
  WgTcProblem prob[WGTC_MAX];
  int A, K, stages, nsplit, ktiles_per_split, split3;
  float* part; int64_t part_split_stride, part_arm_stride;
  const float* bn_mean; const float* bn_rstd;
};

// B tile (MN-major image, see transform_tile): normalise, ones column, optional low half
__device__ __forceinline__ void wg_transform_b(float* hi, float* lo, bool do_split, bool do_norm, const float* bm,
                                               const float* br, int N, int k0, int K, int tid, int nthr) {
#pragma unroll 4
  for (int q = tid; q < TILE_BYTES / 16; q += nthr) {
    const int r = q >> 3, p = q & 7;
    const int slab = r >> 5, rr = r & 31;
    const int col = 32 * slab + 4 * (((((p >> 1) ^ (rr & 3)) << 1)) | (p & 1));
    if (col > N) {                               // beyond the ones column: zeros from the TMA fill stay zeros
      if (do_split) reinterpret_cast<float4*>(lo)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    float4 v = reinterpret_cast<float4*>(hi)[q];
    if (do_norm) {
      const float4 m = *reinterpret_cast<const float4*>(bm + col), rs = *reinterpret_cast<const float4*>(br + col);
      v.x = (v.x - m.x) * rs.x; v.y = (v.y - m.y) * rs.y; v.z = (v.z - m.z) * rs.z; v.w = (v.w - m.w) * rs.w;
    }
    if (col + 4 > N) {                           // the chunk that holds column N
      const float one = (k0 + rr < K) ? 1.f : 0.f;
      const int e = N - col;
      if (e == 0) { v.x = one; v.y = 0.f; v.z = 0.f; v.w = 0.f; }
      else if (e == 1) { v.y = one; v.z = 0.f; v.w = 0.f; }
      else if (e == 2) { v.z = one; v.w = 0.f; }
      else v.w = one;
    }
    reinterpret_cast<float4*>(hi)[q] = v;
    if (do_split) {
      float4 l;
      l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
      reinterpret_cast<float4*>(lo)[q] = l;
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) wg_tc_kernel(const __grid_constant__ WgTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer stays derived from smem_raw, so the compiler keeps the
  // shared address space (LDS / direct mbarrier addresses instead of generic loads and 64-bit window arithmetic)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, pi = blockIdx.y, arm = blockIdx.z;
  const WgTcProblem& pr = P.prob[pi];
  const bool split3 = P.split3 != 0;
  const int tiles_per_stage = split3 ? 4 : 2;
  const int stage_bytes = tiles_per_stage * TILE_BYTES;
  const int S = P.stages;
  auto tileA = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto tileB = [&](int s) { return smem + (size_t)s * stage_bytes + TILE_BYTES; };
  auto tileAlo = [&](int s) { return smem + (size_t)s * stage_bytes + 2 * TILE_BYTES; };
  auto tileBlo = [&](int s) { return smem + (size_t)s * stage_bytes + 3 * TILE_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* ready = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* tmem_full = bars + 3 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);
  uint8_t* bm_raw = reinterpret_cast<uint8_t*>(bars + 3 * S + 2);
  float* bm = reinterpret_cast<float*>(bm_raw + ((16u - (smem_u32(bm_raw) & 15u)) & 15u));   // [128] | [128], 16-byte aligned
  float* br = bm + 128;

  const int BNp = (pr.N + 1 + 15) / 16 * 16;                 // UMMA N: inputs + the ones column
  const int ktiles_total = (P.K + BK - 1) / BK;
  const int kt0 = split * P.ktiles_per_split;
  const int kt1 = min(ktiles_total, kt0 + P.ktiles_per_split);
  const int nkt = max(kt1 - kt0, 0);
  const bool do_norm = pr.bn_layer >= 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(ready + s, TRANSFORM_THREADS);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 128);
  pdl_trigger();
  pdl_wait();          // PDL: barriers and tensor memory are set up while the backward chain drains
  if (threadIdx.x >= 64 && threadIdx.x < 192) {
    const int i = threadIdx.x - 64;
    const bool have = do_norm && i < pr.N;
    bm[i] = have ? P.bn_mean[(pr.bn_layer * P.A + arm) * 128 + i] : 0.f;
    br[i] = have ? P.bn_rstd[(pr.bn_layer * P.A + arm) * 128 + i] : 1.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const CUtensorMap* tmA = &P.tmA[pi];
      const CUtensorMap* tmB = &P.tmB[pi];
      for (int i = 0; i < nkt; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, 2 * TILE_BYTES);
        const int k0 = (kt0 + i) * BK;
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_3d(tmA, full + s, tileA(s) + j * 4096, 32 * j, k0, arm);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_3d(tmB, full + s, tileB(s) + j * 4096, 32 * j, k0, arm);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BNp, true, true);
      const int kround = (P.K + UK - 1) / UK * UK;
      uint32_t acc = 0;
      for (int i = 0; i < nkt; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(ready + s, ph);
        tc_fence_after();
        const int k0 = (kt0 + i) * BK;
        const uint32_t a_hi = smem_u32(tileA(s)), b_hi = smem_u32(tileB(s));
        const uint32_t a_lo = smem_u32(tileAlo(s)), b_lo = smem_u32(tileBlo(s));
#pragma unroll
        for (int ks = 0; ks < BK / UK; ++ks) {
          if (k0 + ks * UK >= kround) break;
          const uint32_t off = ks * 1024;
          const uint64_t dah = make_smem_desc(a_hi + off, 4096, 512, true);
          const uint64_t dbh = make_smem_desc(b_hi + off, 4096, 512, true);
          if (split3) {
            umma_tf32(tmem_base, make_smem_desc(a_lo + off, 4096, 512, true), dbh, idesc, acc);
            acc = 1;
            umma_tf32(tmem_base, dah, make_smem_desc(b_lo + off, 4096, 512, true), idesc, acc);
          }
          umma_tf32(tmem_base, dah, dbh, idesc, acc);
          acc = 1;
        }
        umma_commit(empty + s);
      }
      umma_commit(tmem_full);
    }
  } else if (warp >= TRANSFORM_WARP0) {
    // ===== operand transform: 4 warps on the input tile (normalise, ones column, low half), 4 on delta (low half) =====
    const int tid = threadIdx.x - TRANSFORM_WARP0 * 32;
    for (int i = 0; i < nkt; ++i) {
      const int s = i % S;
      const uint32_t ph = (i / S) & 1;
      mbar_wait(full + s, ph);
      const int k0 = (kt0 + i) * BK;
      if (split3) {
        if (tid < 128) {
          DropSpec nodrop;
          nodrop.mode = 0;
          transform_tile<true>(reinterpret_cast<float*>(tileA(s)), reinterpret_cast<float*>(tileAlo(s)), true, false, nodrop,
                               arm, 0, k0, tid, 128);
        } else {
          wg_transform_b(reinterpret_cast<float*>(tileB(s)), reinterpret_cast<float*>(tileBlo(s)), true, do_norm, bm, br,
                         pr.N, k0, P.K, tid - 128, 128);
        }
      } else {
        wg_transform_b(reinterpret_cast<float*>(tileB(s)), nullptr, false, do_norm, bm, br, pr.N, k0, P.K, tid,
                       TRANSFORM_THREADS);
      }
      fence_proxy_async();
      mbar_arrive(ready + s);
    }
  } else {
    // ===== epilogue: row j of the accumulator = output unit j; columns = inputs, then the bias gradient =====
    const int quad = warp & 3;
    const int j = quad * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* blk = P.part + (int64_t)split * P.part_split_stride + (int64_t)arm * P.part_arm_stride;
    float* wrow = blk + pr.offW + (int64_t)j * pr.N;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(wrow) & 15) == 0);
    for (int c0 = 0; c0 < BNp; c0 += 16) {
      uint32_t r[16];
      if (nkt > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = 0u;
      }
      if (j < pr.M) {
        if (vec_ok && c0 + 16 <= pr.N) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(wrow + c0)[i] =
                make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                            __uint_as_float(r[4 * i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (c0 + i < pr.N) wrow[c0 + i] = __uint_as_float(r[i]);
            else if (c0 + i == pr.N) blk[pr.offB + j] = __uint_as_float(r[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 128);
}

// ---------------------------------------------------------------------------------------------
// Persistent Linear + affine + activation (the augmenter layers): y = act((x . W^T) * scale + shift), both
// operands K-major.  Same stage pipeline as tc_gemm_kernel, but a CTA walks over output tiles (tile = blockIdx.x + i *
// gridDim.x, m fastest so that neighbouring CTAs share the weight tile in L2) with TWO accumulators in tensor memory: the
// epilogue warps drain tile i (tcgen05.ld, affine, activation, stores) while the TMA / MMA warps are already in the
// main loop of tile i + 1, and the barriers, the TMEM allocation and the tensor-map fetch are paid once per CTA instead
// of once per tile.  Measured on the first (one tile per CTA) version: fc11 of the augmenter spent ~1/3 of every CTA's
// life outside the main loop.
// ---------------------------------------------------------------------------------------------
struct LinArgs {
  int M, N, K, BN, stages, split3, tiles_m, tiles_n;
  float* C; int64_t ldc;
  const float* scale; const float* shift; int act;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const LinArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1 KB alignment by an OFFSET in the shared window: the pointer stays derived from smem_raw, so the compiler keeps the
  // shared address space (LDS / direct mbarrier addresses instead of generic loads and 64-bit window arithmetic)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split3 = args.split3 != 0;
  const bool wide = args.BN > 128;                      // 256-column tiles (single-pass TF32 only): B tile = two TMA boxes
  // 3xTF32, TS form (args.split3 == 2): the hi / lo halves of the activation tile go to TENSOR MEMORY (64 columns per
  // stage behind the two accumulators) and the MMAs take their A operand from there -- measured issue cost of
  // tcgen05.mma kind::tf32 M128 N128 K8: 76 cycles from TMEM vs 109 from shared memory (profiles/r1_mma_issue_cost.txt).
  const bool ts = args.split3 == 2;
  const int tiles_per_stage = split3 ? (ts ? 3 : 4) : (wide ? 3 : 2);
  const int stage_bytes = tiles_per_stage * TILE_BYTES;
  const uint32_t acc_cols = wide ? 256u : 128u;
  const uint32_t tmem_cols = ts ? 512u : 2 * acc_cols;
  constexpr uint32_t COL_A = 256;                       // TS form: A stages at [256 + 64 s, +64): hi 32 | lo 32
  const int S = args.stages;
  auto tileA = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto tileB = [&](int s) { return smem + (size_t)s * stage_bytes + TILE_BYTES; };
  auto tileAlo = [&](int s) { return smem + (size_t)s * stage_bytes + 2 * TILE_BYTES; };
  auto tileBlo = [&](int s) { return smem + (size_t)s * stage_bytes + (ts ? 2 : 3) * TILE_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* ready = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* tfull = bars + 3 * S;          // [2] accumulator complete
  uint64_t* tempty = bars + 3 * S + 2;     // [2] accumulator drained by the 4 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 4);

  const int ntiles = args.tiles_m * args.tiles_n;
  const int nkt = (args.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(ready + s, TRANSFORM_THREADS);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t % args.tiles_m) * BM, n0 = (t / args.tiles_m) * args.BN;
        for (int i = 0; i < nkt; ++i, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, (wide ? 3 : 2) * TILE_BYTES);
          tma_load_3d(&tmA, full + s, tileA(s), i * BK, m0, 0);
          tma_load_3d(&tmB, full + s, tileB(s), i * BK, n0, 0);
          if (wide) tma_load_3d(&tmB, full + s, tileB(s) + TILE_BYTES, i * BK, n0 + 128, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, args.BN, false, false);
      const int kround = (args.K + UK - 1) / UK * UK;
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
        const int buf = j & 1;
        mbar_wait(tempty + buf, ((j >> 1) & 1) ^ 1);        // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t dacc = tmem_base + (uint32_t)buf * acc_cols;
        uint32_t acc = 0;
        for (int i = 0; i < nkt; ++i, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(split3 ? ready + s : full + s, ph);
          tc_fence_after();
          const int k0 = i * BK;
          const uint32_t a_hi = smem_u32(tileA(s)), b_hi = smem_u32(tileB(s));
          const uint32_t a_lo = smem_u32(tileAlo(s)), b_lo = smem_u32(tileBlo(s));
#pragma unroll
          for (int ks = 0; ks < BK / UK; ++ks) {
            if (k0 + ks * UK >= kround) break;
            const uint32_t off = ks * 32;
            const uint64_t dah = make_smem_desc(a_hi + off, 0, 1024, false);
            const uint64_t dbh = make_smem_desc(b_hi + off, 0, 1024, false);
            if (ts) {
              const uint32_t ta = tmem_base + COL_A + (uint32_t)s * 64u + (uint32_t)ks * 8u;     // hi; lo at +32
              umma_tf32_ts(dacc, ta + 32u, dbh, idesc, acc);
              umma_tf32_ts(dacc, ta, make_smem_desc(b_lo + off, 0, 1024, false), idesc, 1u);
              umma_tf32_ts(dacc, ta, dbh, idesc, 1u);
              acc = 1;
              continue;
            }
            if (split3) {
              umma_tf32(dacc, make_smem_desc(a_lo + off, 0, 1024, false), dbh, idesc, acc);
              acc = 1;
              umma_tf32(dacc, dah, make_smem_desc(b_lo + off, 0, 1024, false), idesc, acc);
            }
            umma_tf32(dacc, dah, dbh, idesc, acc);
            acc = 1;
          }
          umma_commit(empty + s);
        }
        umma_commit(tfull + buf);
      }
    }
  } else if (warp >= TRANSFORM_WARP0) {
    // ===== TF32 low halves (3xTF32 mode): 4 warps per operand =====
    if (split3) {
      const int tid = threadIdx.x - TRANSFORM_WARP0 * 32;
      DropSpec nodrop;
      nodrop.mode = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int i = 0; i < nkt; ++i, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full + s, ph);
          if (tid < 128 && ts) {
            // thread = row of the activation tile (TMEM lane quad * 32 + lane): 32 K values -> hi | lo columns of stage s
            const int quad = warp & 3, r = quad * 32 + lane;
            const uint8_t* rowp = tileA(s) + r * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(rowp + ((q ^ (r & 7)) << 4));
              hi[4 * q] = __float_as_uint(v.x); hi[4 * q + 1] = __float_as_uint(v.y);
              hi[4 * q + 2] = __float_as_uint(v.z); hi[4 * q + 3] = __float_as_uint(v.w);
              lo[4 * q] = __float_as_uint(tf32_lo(v.x)); lo[4 * q + 1] = __float_as_uint(tf32_lo(v.y));
              lo[4 * q + 2] = __float_as_uint(tf32_lo(v.z)); lo[4 * q + 3] = __float_as_uint(tf32_lo(v.w));
            }
            const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + COL_A + (uint32_t)s * 64u;
            tmem_st16(ta, hi);
            tmem_st16(ta + 16u, hi + 16);
            tmem_st16(ta + 32u, lo);
            tmem_st16(ta + 48u, lo + 16);
            tmem_st_wait();
            tc_fence_before();
          } else if (tid < 128)
            transform_tile<false>(reinterpret_cast<float*>(tileA(s)), reinterpret_cast<float*>(tileAlo(s)), true, false, nodrop, 0,
                                  0, 0, tid, 128);
          else
            transform_tile<false>(reinterpret_cast<float*>(tileB(s)), reinterpret_cast<float*>(tileBlo(s)), true, false, nodrop, 0,
                                  0, 0, tid - 128, 128);
          fence_proxy_async();
          mbar_arrive(ready + s);
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> affine + activation -> global, one accumulator behind the MMA warp =====
    const int quad = warp & 3;
    int j = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
      const int buf = j & 1;
      const int m0 = (t % args.tiles_m) * BM, n0 = (t / args.tiles_m) * args.BN;
      const int row = m0 + quad * 32 + lane;
      mbar_wait(tfull + buf, (j >> 1) & 1);
      tc_fence_after();
      float* crow = args.C + (int64_t)row * args.ldc;
      const bool vec_ok = ((reinterpret_cast<uintptr_t>(crow) & 15) == 0);
      const uint32_t tsrc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * acc_cols;
      for (int c0 = 0; c0 < args.BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tsrc + (uint32_t)c0, r);
        tmem_ld_wait();
        if (c0 + 16 >= args.BN) {                        // last read of this accumulator: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty + buf);
        }
        const int col = n0 + c0;
        if (row < args.M && col < args.N) {
          if (col + 16 <= args.N) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              r[i] = __float_as_uint(ep_apply(__uint_as_float(r[i]), __ldg(args.scale + col + i), __ldg(args.shift + col + i), args.act));
            if (vec_ok) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4*>(crow + col)[i] =
                    make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                __uint_as_float(r[4 * i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) crow[col + i] = __uint_as_float(r[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (col + i < args.N)
                crow[col + i] = ep_apply(__uint_as_float(r[i]), __ldg(args.scale + col + i), __ldg(args.shift + col + i), args.act);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Operand {
  const float* base;       // matrix stored row-major [outer][inner]
  int64_t pitch;           // floats between rows
  int64_t batch_stride;    // floats between batch entries (0: shared)
  bool mn_major;           // true: inner index is the M/N index (K is the row index)
};

// C[batch][split] (M x N, ldc) = A . B  over K; split-K partials are summed by the caller.
int run_tc_gemm(const Operand& A, const Operand& B, int M, int N, int K, int BN, int batch, int nsplit, int flags,
                const DropSpec& drop, float* C, int64_t ldc, int64_t c_batch_stride, int64_t c_split_stride,
                cudaStream_t s, const float* ep_scale = nullptr, const float* ep_shift = nullptr, int ep_act = 0) {
  CUtensorMap tmA, tmB;
  // K-major: inner = K, outer = M (box rows = 128 / BN); MN-major: inner = M/N, outer = K (box rows = BK)
  int rc = A.mn_major ? make_map(&tmA, A.base, M, K, A.pitch, batch, A.batch_stride, BK, true)
                      : make_map(&tmA, A.base, K, M, A.pitch, batch, A.batch_stride, BM, false);
  if (rc) return rc;
  rc = B.mn_major ? make_map(&tmB, B.base, N, K, B.pitch, batch, B.batch_stride, BK, true)
                  : make_map(&tmB, B.base, K, N, B.pitch, batch, B.batch_stride, 128, false);
  if (rc) return rc;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.M = M; a.N = N; a.K = K; a.BN = BN;
  const int tiles = 2 + ((flags & F_SPLIT_A) ? 1 : 0) + ((flags & F_SPLIT_B) ? 1 : 0);
  a.stages = tiles == 2 ? 6 : (tiles == 3 ? 4 : 3);
  const int ktiles = (K + BK - 1) / BK;
  if (nsplit > ktiles) nsplit = ktiles;
  a.nsplit = nsplit;
  a.ktiles_per_split = (ktiles + nsplit - 1) / nsplit;
  a.a_batched = A.batch_stride > 0; a.b_batched = B.batch_stride > 0;
  a.flags = flags;
  a.C = C; a.ldc = ldc; a.c_batch_stride = c_batch_stride; a.c_split_stride = c_split_stride;
  a.drop = drop;
  a.ep_scale = ep_scale; a.ep_shift = ep_shift; a.ep_act = ep_act;
  MVAE_CHECK_ARG(ep_scale == nullptr || (a.nsplit == 1 && ep_shift != nullptr), "the fused epilogue needs nsplit == 1");
  const size_t smem = (size_t)a.stages * tiles * TILE_BYTES + (3 * a.stages + 2) * 8 + 1024;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch * nsplit);
#define TC_LAUNCH(AM, BMJ)                                                                                         \
  do {                                                                                                             \
    static bool attr[64] = {};                                                                                      \
    if (first_on_device(attr)) {                                                                                                   \
      MVAE_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<AM, BMJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    }                                                                                                              \
    tc_gemm_kernel<AM, BMJ><<<grid, NUM_THREADS, smem, s>>>(tmA, tmB, a);                                          \
  } while (0)
  if (!A.mn_major && !B.mn_major) TC_LAUNCH(false, false);
  else if (!A.mn_major && B.mn_major) TC_LAUNCH(false, true);
  else if (A.mn_major && !B.mn_major) TC_LAUNCH(true, false);
  else TC_LAUNCH(true, true);
#undef TC_LAUNCH
  MVAE_LAUNCH_CHECK();
  return 0;
}

// out[batch][m][n] = sum_s part[s][batch][m][n]
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* part, int64_t split_stride, int64_t batch_stride,
                                                          int64_t ld, int nsplit, float* out, int64_t out_batch_stride,
                                                          int64_t out_ld, int M, int N) {
  const int batch = blockIdx.z;
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int m = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (m >= M || n >= N) return;
  const float* p = part + (int64_t)batch * batch_stride + (int64_t)m * ld + n;
  float v = 0.f;
  for (int s = 0; s < nsplit; ++s) v += p[(int64_t)s * split_stride];
  out[(int64_t)batch * out_batch_stride + (int64_t)m * out_ld + n] = v;
}

int round16(int x) { return (x + 15) / 16 * 16; }

int mma_sm_count_tc() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int choose_split(int tiles_mn, int ktiles, int max_split) {
  // one CTA per SM (the pipeline takes most of the shared memory): pick the split whose grid fills
  // whole waves of 148 CTAs best; ties go to the smaller split (less partial traffic)
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_split && s <= ktiles; ++s) {
    const int ctas = tiles_mn * s;
    const int waves = (ctas + 147) / 148;
    const double eff = (double)ctas / (waves * 148.0);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

}  // namespace

// Weight gradients of the wide narrow-layer problems on tcgen05 (see wg_tc_kernel).  `idx` lists the problems of `a`
// to run (all with TMA-compatible operands, see tc_narrow_wgrad_ok); their split count is written to a.prob[].nsplit.
bool tc_narrow_wgrad_ok(const WgArgs& a, const WgProblem& q) {
  if (get_encode() == nullptr) return false;
  const uintptr_t d = reinterpret_cast<uintptr_t>(a.work + q.delta_off), i = reinterpret_cast<uintptr_t>(a.work + q.in_off);
  return q.nin > 0 && q.nout % 4 == 0 && q.nout <= 128 && q.in_ld % 4 == 0 && q.nin + 1 <= 128 && (d & 15) == 0 &&
         (i & 15) == 0 && q.delta_arm_stride % 4 == 0 && q.in_arm_stride % 4 == 0;
}

int tc_narrow_wgrad(WgArgs& a, const int* idx, int n, int split3, cudaStream_t s, bool share_sm) {
  MVAE_CHECK_ARG(n >= 1 && n <= WGTC_MAX, "tc_narrow_wgrad: %d problems", n);
  WgTcParams P;
  memset(&P, 0, sizeof(P));
  for (int t = 0; t < n; ++t) {
    const WgProblem& q = a.prob[idx[t]];
    // MN-major operands: inner = the M/N index, outer = K (cells), box = 32 x BK
    int rc = make_map(&P.tmA[t], a.work + q.delta_off, q.nout, a.B, q.nout, a.A, q.delta_arm_stride, BK, true);
    if (rc) return rc;
    rc = make_map(&P.tmB[t], a.work + q.in_off, q.nin, a.B, q.in_ld, a.A, q.in_arm_stride, BK, true);
    if (rc) return rc;
    P.prob[t].M = q.nout; P.prob[t].N = q.nin; P.prob[t].bn_layer = q.bn_layer;
    P.prob[t].offW = q.poffW - a.base_off; P.prob[t].offB = q.poffB - a.base_off;
  }
  const int ktiles = (a.B + BK - 1) / BK;
  int nsplit = mma_sm_count_tc() / (n * a.A);
  if (nsplit > kWgMaxSplit) nsplit = kWgMaxSplit;
  if (nsplit > ktiles) nsplit = ktiles;
  if (nsplit < 1) nsplit = 1;
  P.A = a.A; P.K = a.B; P.split3 = split3 ? 1 : 0;
  // share_sm (fused step): 3 stages = 98 KB, so that a CTA of wgrad2 (105 KB), launched on the side branch, fits on the same SM
  P.stages = (split3 || share_sm) ? 3 : 6;
  P.ktiles_per_split = (ktiles + nsplit - 1) / nsplit;
  nsplit = (ktiles + P.ktiles_per_split - 1) / P.ktiles_per_split;
  P.nsplit = nsplit;
  P.part = a.part; P.part_split_stride = a.part_split_stride; P.part_arm_stride = a.part_arm_stride;
  P.bn_mean = a.bn_mean; P.bn_rstd = a.bn_rstd;
  for (int t = 0; t < n; ++t) a.prob[idx[t]].nsplit = nsplit;
  const size_t smem = (size_t)P.stages * (split3 ? 4 : 2) * TILE_BYTES + (3 * P.stages + 2) * 8 + 16 + 256 * 4 + 1024;
  static bool attr[64] = {};
  if (first_on_device(attr)) {
    MVAE_CUDA(cudaFuncSetAttribute(wg_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  launch_pdl(wg_tc_kernel, dim3(nsplit, n, a.A), dim3(NUM_THREADS), smem, s, P);
  MVAE_LAUNCH_CHECK();
  return 0;
}

bool gemm_tc_supported(int B, int D, int H) {
  // TMA needs 16-byte aligned row pitches: D % 4 == 0 (x, W1 rows) and H % 4 == 0 (h10, W11, delta1 rows)
  return D % 4 == 0 && H % 4 == 0 && H <= 128 && get_encode() != nullptr;
}

int tc_part_floats(int A, int Bpad, int Dpad) {
  int64_t a = (int64_t)4 * A * Bpad * 128, b = (int64_t)2 * A * Dpad * 128;
  return (int)(a > b ? a : b);
}

// every gene GEMM as 3xTF32 (precision == 1): separate GEMMs around a materialised x_hat / dY
static int tc_fc11_loss_grad_unfused(const mvae_dims& d, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                      const Work& w, float gscale, int want_grad, cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  float* work = st.work;
  double* acc_loss = reinterpret_cast<double*>(work + w.acc_loss);
  const int split3 = hp.precision == 1 ? (F_SPLIT_A | F_SPLIT_B) : 0;
  DropSpec nodrop;
  memset(&nodrop, 0, sizeof(nodrop));
  // G2: pre = h10 . W11^T  -> big [A][B][D]
  Operand h10{work + w.d[4], H, (int64_t)B * H, false};
  Operand W11{st.params + L.offset[FC11_W], H, L.arm_stride, false};
  int rc = run_tc_gemm(h10, W11, B, D, H, 128, A, 1, split3, nodrop, work + w.big, D, (int64_t)B * D, 0, s);
  if (rc) return rc;
  ReconElemArgs r;
  memset(&r, 0, sizeof(r));
  r.pre = work + w.big; r.x = in.x; r.x_arm_stride = in.x_arm_stride; r.x_row_stride = in.x_row_stride;
  r.params = st.params; r.p_arm_stride = L.arm_stride; r.offB = L.offset[FC11_B];
  r.x_rec = nullptr; r.recon_acc = acc_loss; r.B = B; r.D = D; r.gscale = gscale; r.want_grad = want_grad;
  rc = launch_recon_elem(r, A, s);
  if (rc || !want_grad) return rc;
  float* part = work + w.fc1_part;
  // G3: d h10 = dY . W11   (A = dY K-major over genes; B(k=gene, n=h) = W11[gene][h]: MN-major)
  {
    Operand dY{work + w.big, D, (int64_t)B * D, false};
    Operand W11t{st.params + L.offset[FC11_W], H, L.arm_stride, true};
    const int mt = (B + BM - 1) / BM;
    const int nsplit = choose_split(mt * A, (D + BK - 1) / BK, 4);
    const int64_t bs = (int64_t)w.Bpad * 128, ss = (int64_t)A * bs;
    rc = run_tc_gemm(dY, W11t, B, H, D, round16(H), A, nsplit, split3, nodrop, part, 128, bs, ss, s);
    if (rc) return rc;
    partial_sum_kernel<<<dim3((H + 31) / 32, (B + 7) / 8, A), 256, 0, s>>>(part, ss, bs, 128, nsplit, work + w.g_d10,
                                                                            (int64_t)B * H, H, B, H);
    MVAE_LAUNCH_CHECK();
  }
  // G4: d W11 = dY^T . h10   (A(m=gene,k=row) = dY[row][gene]: MN-major; B(k=row,n=h) = h10[row][h]: MN-major)
  {
    Operand dYt{work + w.big, D, (int64_t)B * D, true};
    Operand h10t{work + w.d[4], H, (int64_t)B * H, true};
    const int mt = (D + BM - 1) / BM;
    const int nsplit = choose_split(mt * A, (B + BK - 1) / BK, 2);
    const int64_t bs = (int64_t)w.Dpad * 128, ss = (int64_t)A * bs;
    rc = run_tc_gemm(dYt, h10t, D, H, B, round16(H), A, nsplit, split3, nodrop, part, 128, bs, ss, s);
    if (rc) return rc;
    partial_sum_kernel<<<dim3((H + 31) / 32, (D + 7) / 8, A), 256, 0, s>>>(part, ss, bs, 128, nsplit,
                                                                            st.grads + L.offset[FC11_W], L.arm_stride, H,
                                                                            D, H);
    MVAE_LAUNCH_CHECK();
  }
  return launch_colsum(work + w.big, (int64_t)B * D, st.grads + L.offset[FC11_B], L.arm_stride, B, D, A, s);
}

int tc_fc11_loss_grad(const mvae_dims& d, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                      const Work& w, float gscale, int want_grad, cudaStream_t s, bool defer_gene_fix) {
  double* acc_loss = reinterpret_cast<double*>(st.work + w.acc_loss);
  if (hp.precision == 1) return tc_fc11_loss_grad_unfused(d, hp, st, in, w, gscale, want_grad, s);
  // fused passes: row owner (x_hat, loss sums, d h10), then gene owner (d fc11.weight, d fc11.bias)
  if (!want_grad) return ts_fc11_rows(d, st, in, w, gscale, 0, nullptr, acc_loss, s);
  return ts_fc11_loss_grad(d, st, in, w, gscale, acc_loss, s, defer_gene_fix);
}

int tc_fc1_wgrad(const mvae_dims& d, const mvae_hparams& hp, const mvae_state& st, const mvae_inputs& in,
                 const DropSpec& drop, const Work& w, cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  // G5: dW1[h][gene] = sum_row delta1[row][h] * xd[row][gene]: both operands MN-major; split over cells
  Operand d1{st.work + w.delta_enc[0], H, (int64_t)B * H, true};
  Operand x{in.x, in.x_row_stride, in.x_arm_stride, true};
  int flags = hp.precision == 1 ? (F_SPLIT_A | F_SPLIT_B) : 0;
  if (drop.mode) flags |= F_DROP_B;
  const int nt = (D + 127) / 128;
  const int nsplit = choose_split(nt * A, (B + BK - 1) / BK, 8);
  if (nsplit == 1)
    return run_tc_gemm(d1, x, H, D, B, 128, A, 1, flags, drop, st.grads + L.offset[FC1_W], D, L.arm_stride, 0, s);
  float* part = st.work + w.fc1_part;                 // [split][A][128][Dpad]
  const int64_t bs = (int64_t)128 * w.Dpad, ss = (int64_t)A * bs;
  int rc = run_tc_gemm(d1, x, H, D, B, 128, A, nsplit, flags, drop, part, w.Dpad, bs, ss, s);
  if (rc) return rc;
  partial_sum_kernel<<<dim3((D + 31) / 32, (H + 7) / 8, A), 256, 0, s>>>(part, ss, bs, w.Dpad, nsplit,
                                                                          st.grads + L.offset[FC1_W], L.arm_stride, D, H, D);
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae

extern "C" int mvae_debug_tc_gemm(const float* A, int a_mn, int64_t a_pitch, const float* B, int b_mn, int64_t b_pitch,
                                  int M, int N, int K, int BN, int nsplit, int flags, float* C, int64_t ldc,
                                  int64_t c_split_stride, void* stream) {
  using namespace mvae;
  MVAE_CHECK_ARG(A && B && C, "null argument");
  MVAE_CHECK_ARG(BN % 16 == 0 && BN >= 16 && BN <= 128, "BN must be a multiple of 16 in [16,128]");
  MVAE_CHECK_ARG((flags & ~3) == 0, "only the TF32-split flags (1: A, 2: B) are accepted here");
  Operand a{A, a_pitch, 0, a_mn != 0};
  Operand b{B, b_pitch, 0, b_mn != 0};
  DropSpec nodrop;
  memset(&nodrop, 0, sizeof(nodrop));
  return run_tc_gemm(a, b, M, N, K, BN, 1, nsplit, flags, nodrop, C, ldc, 0, c_split_stride, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Augmenter forward (mmidas/augmentation/udagan.py:217-329, eval mode): a chain of Linear layers whose
// BatchNorm (running statistics) / bias / activation collapse into a per-column affine + activation epilogue.
// ---------------------------------------------------------------------------------------------
namespace mvae {
namespace {
// scale = gamma / sqrt(var + eps), shift = (bias - mean) * scale + beta   (null mean/var: plain bias)
__global__ void __launch_bounds__(256) fold_affine_kernel(const float* bias, const float* mean, const float* var,
                                                          const float* gamma, const float* beta, float eps, int n,
                                                          float* scale, float* shift) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float sc = 1.f, sh = bias ? bias[i] : 0.f;
  if (var) {
    sc = (gamma ? gamma[i] : 1.f) / sqrtf(var[i] + eps);
    sh = (sh - mean[i]) * sc + (beta ? beta[i] : 0.f);
  }
  scale[i] = sc;
  shift[i] = sh;
}
// out[r][c] = a[r][c] * b[r][c] + c0[r][c] on [rows x n] views with their own pitches (reparameterisation s = eps * sigma + mu)
__global__ void __launch_bounds__(256) fma_rows_kernel(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c0,
                                                       int64_t ldc0, float* out, int64_t ldo, int64_t rows, int n, float a_scale) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * n) return;
  const int64_t r = idx / n;
  const int c = (int)(idx - r * n);
  const float bv = b ? b[r * ldb + c] : 1.f, cv = c0 ? c0[r * ldc0 + c] : 0.f;
  out[r * ldo + c] = fmaf(a_scale * a[r * lda + c], bv, cv);
}
}  // namespace
}  // namespace mvae

namespace mvae {
int launch_fold_affine(const float* bias, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                       int n, float* scale, float* shift, cudaStream_t s) {
  fold_affine_kernel<<<(n + 255) / 256, 256, 0, s>>>(bias, mean, var, gamma, beta, eps, n, scale, shift);
  MVAE_LAUNCH_CHECK();
  return 0;
}
int launch_fma_rows(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, float* out, int64_t ldo,
                    int64_t rows, int n, float a_scale, cudaStream_t s) {
  if (rows == 0) return 0;
  fma_rows_kernel<<<(unsigned)((rows * n + 255) / 256), 256, 0, s>>>(a, lda, b, ldb, c, ldc, out, ldo, rows, n, a_scale);
  MVAE_LAUNCH_CHECK();
  return 0;
}
// y[rows][n_out] (pitch y_pitch) = act((x[rows][k] . w[n_out][k]^T) * scale + shift)
int tc_linear_act(const float* x, int64_t x_pitch, const float* w, int64_t w_pitch, float* y, int64_t y_pitch, int64_t rows,
                  int n_out, int k, const float* scale, const float* shift, int act, int split3, cudaStream_t s) {
  int BN = n_out >= 128 ? 128 : round16(n_out);
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, x, k, rows, x_pitch, 1, 0, BM, false);
  if (rc) return rc;
  rc = make_map(&tmB, w, k, n_out, w_pitch, 1, 0, 128, false);
  if (rc) return rc;
  // single-pass TF32 is bound by L2 -> shared-memory operand traffic: 256-column tiles move 24 KB instead of 32 KB per
  // 128 x 128 x 32 block of products (3xTF32 is MMA-issue-bound and its four tiles per stage leave no room for them)
  if (!split3 && n_out >= 384 && k >= 768) BN = 256;   // short-K layers are epilogue-bound: more, smaller tiles win there
  LinArgs a;
  memset(&a, 0, sizeof(a));
  a.M = (int)rows; a.N = n_out; a.K = k; a.BN = BN;
  a.split3 = split3 ? 2 : 0;                    // 2: 3xTF32 with the activation halves in tensor memory (TS-form MMAs)
  a.stages = split3 ? 4 : (BN > 128 ? 4 : 6);
  a.tiles_m = (int)((rows + BM - 1) / BM); a.tiles_n = (n_out + BN - 1) / BN;
  a.C = y; a.ldc = y_pitch; a.scale = scale; a.shift = shift; a.act = act;
  const size_t smem = (size_t)a.stages * (split3 ? 3 : (BN > 128 ? 3 : 2)) * TILE_BYTES + (3 * a.stages + 6) * 8 + 1024;
  static bool attr[64] = {};
  if (first_on_device(attr)) {
    MVAE_CUDA(cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int ntiles = a.tiles_m * a.tiles_n;
  const int grid = ntiles < mma_sm_count_tc() ? ntiles : mma_sm_count_tc();
  tc_linear_kernel<<<grid, NUM_THREADS, smem, s>>>(tmA, tmB, a);
  MVAE_LAUNCH_CHECK();
  return 0;
}
}  // namespace mvae
