// kernels_chain.cu — the decoder stack fc7..fc10 (nn_model.py:281-284) as ONE kernel per direction.
//
// There is no batch coupling between these four layers (no BatchNorm), so a CTA keeps its 64 cells in
// shared memory from h6 to h10 (forward) / from d h10 to d h6 (backward) and only streams the weights:
// the next layer's weight matrix is fetched with cp.async while the current layer is multiplied.  Same
// warp-level 3xTF32 MMA as kernels_mma.cu; weights are used in their natural [out][in] layout in both
// directions (forward: B(k,n) = W[n][k] with pitch = 4 mod 8; backward: B(k,n) = W[k][n], pitch = 8 mod 32).
#include "common.cuh"
#include "kernels.h"

namespace mvae {

namespace {

// A CTA owns 16*MT cells (MT row groups x 2 column halves = 2*MT warps).  MT = 5 when that makes the
// grid fit one wave of 148 CTAs (B = 5000, A = 2: 126 CTAs), else 4.  Pitches depend on H (runtime):
//   XP  = round8(H) + 4   activation tiles and forward weights ([out][in]): pitch % 8 == 4
//   WPB = round8(H) (+8..) backward weights read as [k][n]: pitch % 32 in {8, 24}

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v);
  lo = __float_as_uint(v - __uint_as_float(hi & 0xFFFFE000u));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// copy a [rows x cols] fp32 matrix (row pitch ld) into smem (row pitch sp); 16-byte cp.async when every
// row start is 16-byte aligned, scalar otherwise.  Rows >= rows_valid are zero-filled by the caller.
__device__ __forceinline__ void async_tile(float* dst, int sp, const float* src, int64_t ld, int rows_valid, int cols,
                                           int tid, int nthr) {
  if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (ld % 4 == 0) && (cols % 4 == 0)) {
    const int cpr = cols / 4;
    for (int idx = tid; idx < rows_valid * cpr; idx += nthr) {
      const int r = idx / cpr, c = (idx - r * cpr) * 4;
      cp_async16(dst + r * sp + c, src + (int64_t)r * ld + c);
    }
  } else {
    for (int idx = tid; idx < rows_valid * cols; idx += nthr) {
      const int r = idx / cols, c = idx - r * cols;
      dst[r * sp + c] = src[(int64_t)r * ld + c];
    }
  }
}

// acc[NTW][4] += A(16 x 8*ksteps) . B ; A(m,k) = As[m*ap + k];  B(k,n) = BT ? Bs[n*bp + k] : Bs[k*bp + n]
template <int NTW, bool BT, bool SPLIT>
__device__ __forceinline__ void warp_gemm2(const float* __restrict__ As, int ap, const float* __restrict__ Bs, int bp,
                                           int ksteps, int nt_used, float (&acc)[NTW][4], int lane) {
  const int g = lane >> 2, tig = lane & 3;
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k0 = ks * 8;
    const float av[4] = {As[g * ap + k0 + tig], As[(g + 8) * ap + k0 + tig], As[g * ap + k0 + tig + 4],
                         As[(g + 8) * ap + k0 + tig + 4]};
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt < nt_used) {
        float b0, b1;
        if (BT) {
          b0 = Bs[(nt * 8 + g) * bp + k0 + tig];
          b1 = Bs[(nt * 8 + g) * bp + k0 + tig + 4];
        } else {
          b0 = Bs[(k0 + tig) * bp + nt * 8 + g];
          b1 = Bs[(k0 + tig + 4) * bp + nt * 8 + g];
        }
        uint32_t bh[2], bl[2];
        if (SPLIT) {
          split_tf32(b0, bh[0], bl[0]);
          split_tf32(b1, bh[1], bl[1]);
          mma_tf32(acc[nt], al, bh);
          mma_tf32(acc[nt], ah, bl);
        } else {
          bh[0] = __float_as_uint(b0);
          bh[1] = __float_as_uint(b1);
        }
        mma_tf32(acc[nt], ah, bh);
      }
    }
  }
}

#ifdef CHAIN_STAMPS
// development only (-DCHAIN_STAMPS, profiles/tools/chain_stamps.py): clock64 at the phase boundaries of CTA (0, 0)
__device__ long long g_chain_stamps[2][64];
#define CSTAMP(dir, k)                                                                          \
  do {                                                                                          \
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_chain_stamps[dir][k] = clock64(); \
  } while (0)
#else
#define CSTAMP(dir, k) do { } while (0)
#endif

// Two buffers of a ping-pong pair as "base + index * stride": a run-time index into an ARRAY of pointers makes the compiler
// lose the shared address space (generic LD / ST with 64-bit addresses in the MMA loops of the multi-tile kernels)
struct SmemPair {
  float* base; int stride;
  __device__ __forceinline__ float* operator[](int i) const { return base + i * stride; }
};

struct ChainArgs {
  int XP, WPB, Hp;               // pitches (floats) and round8(H)
  const float* params; int64_t p_arm_stride;
  int64_t offW[4], offB[4];      // fc7..fc10
  int B, H, L;
  // forward: in = h6 [A][B][L]; out[l] = h7..h10 [A][B][H]
  const float* h6; float* hout[4];
  // backward: g10 = d loss / d h10 [A][B][H]; act[l] = h7..h10; delta[l] out; g6 out [A][B][L]
  const float* g10; const float* act[4]; float* delta[4]; float* g6;
};

// =============================================================================================
// forward chain
// =============================================================================================
template <int MT, bool SPLIT, int HC, int LC>
__global__ void __launch_bounds__(128 * MT) dec_chain_fwd_kernel(const ChainArgs p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CR = 16 * MT, CT = 128 * MT;     // MT row groups x 4 column groups of warps (latency-bound: more warps)
  constexpr bool FAST = HC > 0;
  const int XP = FAST ? (((HC + 7) & ~7) + 4) : p.XP, WPF = XP, Hp = FAST ? ((HC + 7) & ~7) : p.Hp;
  float* Ws0 = smem;                       // [Hp][WPF]
  float* Ws1 = Ws0 + Hp * WPF;
  float* Xs0 = Ws1 + Hp * WPF;             // [CR][XP]
  float* Xs1 = Xs0 + CR * XP;
  float* bias = Xs1 + CR * XP;             // [4][128]
  const int arm = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int warp = tid >> 5, wr = warp % MT, wc = warp / MT;
  const int g = lane >> 2, tig = lane & 3;
  const int H = FAST ? HC : p.H, L = FAST ? LC : p.L, B = p.B;
  const int row0 = blockIdx.x * CR;
  const int rows_valid = min(CR, B - row0);
  const float* par = p.params + (int64_t)arm * p.p_arm_stride;

  pdl_trigger();
  for (int idx = tid; idx < 2 * Hp * WPF + 2 * CR * XP; idx += CT) smem[idx] = 0.f;
  for (int idx = tid; idx < 4 * 128; idx += CT) {
    const int l = idx >> 7, j = idx & 127;
    bias[idx] = j < H ? par[p.offB[l] + j] : 0.f;
  }
  __syncthreads();
  // layer 0 operands (fc7: [H][L], rows of L floats: scalar path) + prefetch of fc8
  async_tile(Ws0, WPF, par + p.offW[0], L, H, L, tid, CT);
  pdl_wait();          // PDL: everything above touches parameters only and overlaps the tail of the head kernel
  async_tile(Xs0, XP, p.h6 + ((int64_t)arm * B + row0) * L, L, rows_valid, L, tid, CT);
  cp_async_commit();
  async_tile(Ws1, WPF, par + p.offW[1], H, H, H, tid, CT);
  cp_async_commit();

  const SmemPair Ws{Ws0, (int)(Ws1 - Ws0)};
  const SmemPair Xs{Xs0, (int)(Xs1 - Xs0)};
  constexpr int NTW = 4;
  constexpr int UNR = FAST ? 4 : 1;
#pragma unroll UNR
  for (int l = 0; l < 4; ++l) {
    const int K = l == 0 ? L : H;
    const int ksteps = (K + 7) / 8;
    if (l < 3) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    float acc[NTW][4];
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const int nt_used = max(0, min(NTW, (H + 7) / 8 - wc * NTW));
    {
      const float* Aw = Xs[l & 1] + wr * 16 * XP;
      const float* Bw = Ws[l & 1] + wc * NTW * 8 * WPF;
      if (FAST) {
        const int nt_total = (H + 7) / 8, full = nt_total / NTW, rem = nt_total % NTW;      // constants
        if (wc < full) warp_gemm2<NTW, true, SPLIT>(Aw, XP, Bw, WPF, ksteps, NTW, acc, lane);
        else if (wc == full && rem > 0) warp_gemm2<NTW, true, SPLIT>(Aw, XP, Bw, WPF, ksteps, rem, acc, lane);
      } else {
        warp_gemm2<NTW, true, SPLIT>(Aw, XP, Bw, WPF, ksteps, nt_used, acc, lane);
      }
    }
    // epilogue: bias + ReLU -> global h_{7+l} and the next layer's operand tile
    float* out = p.hout[l] + ((int64_t)arm * B + row0) * H;
    float* Xn = Xs[(l + 1) & 1];
    const int ra = wr * 16 + g, rb = ra + 8;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt < nt_used) {
        const int c = (wc * NTW + nt) * 8 + 2 * tig;
        const float b0 = bias[l * 128 + c], b1 = bias[l * 128 + c + 1];
        const float v00 = fmaxf(acc[nt][0] + b0, 0.f), v01 = fmaxf(acc[nt][1] + b1, 0.f);
        const float v10 = fmaxf(acc[nt][2] + b0, 0.f), v11 = fmaxf(acc[nt][3] + b1, 0.f);
        if (c < H) {      // H is even in practice; handle the odd tail element-wise
          if (ra < rows_valid) { out[(int64_t)ra * H + c] = v00; if (c + 1 < H) out[(int64_t)ra * H + c + 1] = v01; }
          if (rb < rows_valid) { out[(int64_t)rb * H + c] = v10; if (c + 1 < H) out[(int64_t)rb * H + c + 1] = v11; }
          Xn[ra * XP + c] = v00; Xn[rb * XP + c] = v10;
          if (c + 1 < H) { Xn[ra * XP + c + 1] = v01; Xn[rb * XP + c + 1] = v11; }
        }
      }
    }
    __syncthreads();                       // everyone is done with Ws[l&1] / Xs[l&1]
    if (l + 2 < 4) {
      async_tile(Ws[l & 1], WPF, par + p.offW[l + 2], H, H, H, tid, CT);
      cp_async_commit();
    }
  }
}

// =============================================================================================
// backward chain: delta_l = g * [h_l > 0];  g_{l-1} = delta_l . W_l
// =============================================================================================
template <int MT, bool SPLIT, int HC, int LC>
__global__ void __launch_bounds__(128 * MT) dec_chain_bwd_kernel(const ChainArgs p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CR = 16 * MT, CT = 128 * MT;
  constexpr bool FAST = HC > 0;
  constexpr int HPC = (HC + 7) & ~7;
  const int XP = FAST ? HPC + 4 : p.XP, Hp = FAST ? HPC : p.Hp;
  const int WPB = FAST ? (((HPC & 31) == 8 || (HPC & 31) == 24) ? HPC : p.WPB) : p.WPB;
  float* Ws0 = smem;                       // [Hp][WPB]  W_l natural: row j (out), col i (in)
  float* Ws1 = Ws0 + Hp * WPB;
  float* Gs = Ws1 + Hp * WPB;              // [CR][XP]   gradient wrt h_l, then delta_l in place
  float* Ms0 = Gs + CR * XP;               // [CR][XP]   h_l tile (ReLU mask)
  float* Ms1 = Ms0 + CR * XP;
  const int arm = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int warp = tid >> 5, wr = warp % MT, wc = warp / MT;
  const int g = lane >> 2, tig = lane & 3;
  const int H = FAST ? HC : p.H, L = FAST ? LC : p.L, B = p.B;
  const int row0 = blockIdx.x * CR;
  const int rows_valid = min(CR, B - row0);
  const float* par = p.params + (int64_t)arm * p.p_arm_stride;
  const int64_t rbase = (int64_t)arm * B + row0;

  pdl_trigger();
  for (int idx = tid; idx < 2 * Hp * WPB + 3 * CR * XP; idx += CT) smem[idx] = 0.f;
  __syncthreads();
  // group 0: g10, h10, W10 ; group 1: h9, W9
  async_tile(Ws0, WPB, par + p.offW[3], H, H, H, tid, CT);
  pdl_wait();          // PDL: the set-up and the first weight tile overlap the tail of the fc11 fix-up
  async_tile(Gs, XP, p.g10 + rbase * H, H, rows_valid, H, tid, CT);
  async_tile(Ms0, XP, p.act[3] + rbase * H, H, rows_valid, H, tid, CT);
  cp_async_commit();
  async_tile(Ms1, XP, p.act[2] + rbase * H, H, rows_valid, H, tid, CT);
  async_tile(Ws1, WPB, par + p.offW[2], H, H, H, tid, CT);
  cp_async_commit();

  const SmemPair Ws{Ws0, (int)(Ws1 - Ws0)};
  const SmemPair Ms{Ms0, (int)(Ms1 - Ms0)};
  constexpr int NTW = 4;
  constexpr int UNR = FAST ? 4 : 1;
#pragma unroll UNR
  for (int it = 0; it < 4; ++it) {
    const int l = 3 - it;                  // layer index 3..0 = fc10..fc7
    const int nin = l == 0 ? L : H;        // inputs of this layer
    if (it < 3) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    // delta = g * relu'(h_l): in place in Gs, and to global for the weight-gradient kernel
    float* dout = p.delta[l] + rbase * H;
    const float* Mcur = Ms[it & 1];
    for (int idx = tid; idx < CR * (H / 2 + (H & 1)); idx += CT) {
      const int hw = H / 2 + (H & 1);
      const int r = idx / hw, c = (idx - r * hw) * 2;
      float d0 = Mcur[r * XP + c] > 0.f ? Gs[r * XP + c] : 0.f;
      float d1 = (c + 1 < H && Mcur[r * XP + c + 1] > 0.f) ? Gs[r * XP + c + 1] : 0.f;
      Gs[r * XP + c] = d0;
      if (c + 1 < H) Gs[r * XP + c + 1] = d1;
      if (r < rows_valid) {
        dout[(int64_t)r * H + c] = d0;
        if (c + 1 < H) dout[(int64_t)r * H + c + 1] = d1;
      }
    }
    __syncthreads();
    float acc[NTW][4];
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const int nt_used = max(0, min(NTW, (nin + 7) / 8 - wc * NTW));
    {
      const float* Aw = Gs + wr * 16 * XP;
      const float* Bw = Ws[it & 1] + wc * NTW * 8;
      if (FAST) {
        const int nt_total = (nin + 7) / 8, full = nt_total / NTW, rem = nt_total % NTW;    // constants once unrolled
        if (wc < full) warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (H + 7) / 8, NTW, acc, lane);
        else if (wc == full && rem > 0) warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (H + 7) / 8, rem, acc, lane);
      } else {
        warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (H + 7) / 8, nt_used, acc, lane);
      }
    }
    __syncthreads();                       // all warps have read delta before it is overwritten by the new g
    const int ra = wr * 16 + g, rb = ra + 8;
    float* g6 = p.g6 + rbase * L;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt < nt_used) {
        const int c = (wc * NTW + nt) * 8 + 2 * tig;
        if (l > 0) {
          Gs[ra * XP + c] = acc[nt][0]; Gs[ra * XP + c + 1] = acc[nt][1];
          Gs[rb * XP + c] = acc[nt][2]; Gs[rb * XP + c + 1] = acc[nt][3];
        } else {
          if (ra < rows_valid) { if (c < L) g6[(int64_t)ra * L + c] = acc[nt][0]; if (c + 1 < L) g6[(int64_t)ra * L + c + 1] = acc[nt][1]; }
          if (rb < rows_valid) { if (c < L) g6[(int64_t)rb * L + c] = acc[nt][2]; if (c + 1 < L) g6[(int64_t)rb * L + c + 1] = acc[nt][3]; }
        }
      }
    }
    // stage the operands of iteration it+2 into the buffers just released
    if (it + 2 < 4) {
      const int l2 = l - 2;
      __syncthreads();
      async_tile(Ms[it & 1], XP, p.act[l2] + rbase * H, H, rows_valid, H, tid, CT);
      if (l2 == 0) {
        for (int idx = tid; idx < Hp * WPB; idx += CT) Ws[it & 1][idx] = 0.f;   // fc7 is [H][L]: clear the wider fc9
        __syncthreads();
        async_tile(Ws[it & 1], WPB, par + p.offW[0], L, H, L, tid, CT);
      } else {
        async_tile(Ws[it & 1], WPB, par + p.offW[l2], H, H, H, tid, CT);
      }
      cp_async_commit();
    }
  }
}

// =============================================================================================
// encoder middle: fc2..fc5 with the BatchNorm of the previous layer folded into the operand (nn_model.py:264-268).
// BatchNorm needs batch-global statistics between layers, so this is a COOPERATIVE kernel (one co-resident wave):
// every CTA keeps its rows in shared memory from a1 to a5, adds its fp64 column sums to the global accumulators and
// meets the other CTAs at a grid barrier before it normalises the next layer's operand.  Weights are streamed
// with cp.async one layer ahead, as in the decoder chain.
// =============================================================================================
struct EncChainArgs {
  int XP, Hp;
  const float* params; int64_t p_arm_stride;
  int64_t offW[4], offB[4];      // fc2..fc5
  int A, B, H, L;
  const float* a1;               // [A][B][H] relu(fc1), before batch_l1
  float* a1_out;                 // same buffer, written by the kernel when fx.valid (fc1 fix-up fused in)
  Fc1Deferred fx;                // fc1's stream-K partial tiles (ts_gemm.cu), or valid == 0
  float* aout[4];                // a2..a4 [A][B][H], a5 [A][B][L]
  double* sums;                  // acc_fwd: column sums / sums of squares, layer l at (l * A + arm) * 256
  float* bn_mean; float* bn_rstd;   // [5][A][128]
  unsigned int* bar;             // grid-barrier counters (zeroed with the accumulators)
  float eps;
};

__device__ __forceinline__ double group_sum_d(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < n);
    __threadfence();
  }
  __syncthreads();
}

// HC / LC > 0: fc_dim / lowD_dim known at compile time (the reference defaults 100 / 10): pitches, k-step counts and the
// number of column tiles per warp become constants, which shrinks the MMA loop ~5x (it is issue/latency-bound).
template <int MT, int NWC, int HC, int LC>
__global__ void __launch_bounds__(32 * MT * NWC) enc_chain_fwd_kernel(const EncChainArgs p) {
  extern __shared__ __align__(16) float smem[];
  pdl_trigger();
  CSTAMP(0, 0);
  constexpr int CR = 16 * MT, CT = 32 * MT * NWC, NTW = 16 / NWC;   // MT row groups x NWC column groups of warps
  constexpr bool FAST = HC > 0;
  const int XP = FAST ? (((HC + 7) & ~7) + 4) : p.XP, Hp = FAST ? ((HC + 7) & ~7) : p.Hp;
  float* Ws0 = smem;                       // [Hp][XP]  W natural [out][in]
  float* Ws1 = Ws0 + Hp * XP;
  float* Xs0 = Ws1 + Hp * XP;              // [CR][XP]
  float* Xs1 = Xs0 + CR * XP;
  float* bias = Xs1 + CR * XP;             // [4][128]
  float* mean = bias + 4 * 128;            // [128]
  float* rstd = mean + 128;                // [128]
  double* red = reinterpret_cast<double*>(rstd + 128);   // [MT][2][128]
  const int arm = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int warp = tid >> 5, wr = warp % MT, wc = warp / MT;
  const int g = lane >> 2, tig = lane & 3;
  const int H = FAST ? HC : p.H, L = FAST ? LC : p.L, B = p.B;
  const int row0 = blockIdx.x * CR;
  const int rows_valid = min(CR, B - row0);
  const float* par = p.params + (int64_t)arm * p.p_arm_stride;
  const unsigned int nctas = gridDim.x * gridDim.y;

  for (int idx = tid; idx < 2 * Hp * XP + 2 * CR * XP; idx += CT) smem[idx] = 0.f;
  for (int idx = tid; idx < 4 * 128; idx += CT) {
    const int l = idx >> 7, j = idx & 127;
    const int nout = l < 3 ? H : L;
    bias[idx] = j < nout ? par[p.offB[l] + j] : 0.f;
  }
  __syncthreads();
  async_tile(Ws0, XP, par + p.offW[0], H, H, H, tid, CT);
  if (!p.fx.valid) async_tile(Xs0, XP, p.a1 + ((int64_t)arm * B + row0) * H, H, rows_valid, H, tid, CT);
  cp_async_commit();
  async_tile(Ws1, XP, par + p.offW[1], H, H, H, tid, CT);
  cp_async_commit();

  if (p.fx.valid) {
    // ---- fc1 fix-up fused in (instead of a separate launch): a1 = relu(scale * (sum of the stream-K partials of fc1, in
    // CTA order) + b1) for this CTA's rows -> global (backward, weight gradients) and the layer-0 operand tile; fp64 column
    // sums for batch_l1 -> one more grid barrier.  Warp w owns rows w, w + nwarps, ...; lane = float4 column.  (H % 4 == 0.)
    constexpr int NWARP = MT * NWC;
    double* red0 = reinterpret_cast<double*>(Xs1);        // [NWARP][2][Hp] scratch: Xs1 is rewritten by layer 0's epilogue
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const bool col_ok = 4 * lane < H;
    float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok) bj = *reinterpret_cast<const float4*>(par + p.fx.offB + 4 * lane);
    // (rows in groups of 4 with up to 4 partials each issued together: the loop is bound by L2 latency, not by bytes)
    for (int rr0 = warp; rr0 < rows_valid; rr0 += 4 * NWARP) {
      const float* base[4];
      int np_[4];
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = rr0 + k * NWARP, r = row0 + rr;
        const int64_t t = (int64_t)(r >> 8) * p.fx.batch + arm;          // 256-row stream-K tile of fc1 (two 128-row blocks)
        const int c0 = (int)(((t * p.fx.ktiles + 1) * p.fx.G - 1) / p.fx.U);
        const int c1 = (int)(((t * p.fx.ktiles + p.fx.ktiles) * p.fx.G - 1) / p.fx.U);
        np_[k] = (rr < rows_valid && col_ok) ? c1 - c0 + 1 : 0;
        base[k] = p.fx.part + ((int64_t)(c0 + t) * 2 + ((r >> 7) & 1)) * (128 * 128) + (r & 127) * 128 + 4 * lane;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4 q[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          q[k][j] = j < np_[k] ? __ldcg(reinterpret_cast<const float4*>(base[k] + (int64_t)j * 2 * (128 * 128))) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {           // CTA order: j = 0 first (adding the zeros of absent partials changes nothing)
          v[k].x += q[k][j].x; v[k].y += q[k][j].y; v[k].z += q[k][j].z; v[k].w += q[k][j].w;
        }
        for (int j = 4; j < np_[k]; ++j) {      // a tile cut into more than 4 CTA shares (tiny grids)
          const float4 qq = __ldcg(reinterpret_cast<const float4*>(base[k] + (int64_t)j * 2 * (128 * 128)));
          v[k].x += qq.x; v[k].y += qq.y; v[k].z += qq.z; v[k].w += qq.w;
        }
        if (np_[k] > 0) {
          const int rr = rr0 + k * NWARP, r = row0 + rr;
          float4 o;
          o.x = fmaxf(fmaf(v[k].x, p.fx.scale, bj.x), 0.f); o.y = fmaxf(fmaf(v[k].y, p.fx.scale, bj.y), 0.f);
          o.z = fmaxf(fmaf(v[k].z, p.fx.scale, bj.z), 0.f); o.w = fmaxf(fmaf(v[k].w, p.fx.scale, bj.w), 0.f);
          *reinterpret_cast<float4*>(p.a1_out + ((int64_t)arm * B + r) * H + 4 * lane) = o;
          *reinterpret_cast<float4*>(Xs0 + rr * XP + 4 * lane) = o;
          s1[0] += (double)o.x; s1[1] += (double)o.y; s1[2] += (double)o.z; s1[3] += (double)o.w;
          s2[0] += (double)o.x * (double)o.x; s2[1] += (double)o.y * (double)o.y;
          s2[2] += (double)o.z * (double)o.z; s2[3] += (double)o.w * (double)o.w;
        }
      }
    }
    if (col_ok) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        red0[(warp * 2 + 0) * Hp + 4 * lane + e] = s1[e];
        red0[(warp * 2 + 1) * Hp + 4 * lane + e] = s2[e];
      }
    }
    __syncthreads();
    for (int j = tid; j < 256; j += CT) {
      const int which = j >> 7, cc = j & 127;
      if (cc < H) {
        double sacc = 0.0;
        for (int w = 0; w < NWARP; ++w) sacc += red0[(w * 2 + which) * Hp + cc];
        atomicAdd(p.sums + (int64_t)arm * 256 + which * 128 + cc, sacc);
      }
    }
    grid_barrier(p.bar + 3, nctas);
  }
  CSTAMP(0, 1);

  const SmemPair Ws{Ws0, (int)(Ws1 - Ws0)};
  const SmemPair Xs{Xs0, (int)(Xs1 - Xs0)};
  constexpr int UNR = FAST ? 4 : 1;
#pragma unroll UNR
  for (int l = 0; l < 4; ++l) {
    const int nout = l < 3 ? H : L;
    // ---- batch statistics of this layer's input (complete: previous kernel for l == 0, grid barrier otherwise)
    if (tid < H) {
      const double* sums = p.sums + (int64_t)(l * p.A + arm) * 256;
      const double s1 = __ldcg(sums + tid), s2 = __ldcg(sums + 128 + tid);
      const double m = s1 / (double)B;
      double var = s2 / (double)B - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)p.eps));
      mean[tid] = mf;
      rstd[tid] = rf;
      if (blockIdx.x == 0) {
        p.bn_mean[(l * p.A + arm) * 128 + tid] = mf;
        p.bn_rstd[(l * p.A + arm) * 128 + tid] = rf;
      }
    }
    if (l < 3) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    CSTAMP(0, 2 + 5 * l);
    // ---- normalise the operand tile in place (rows beyond the batch stay zero)
    float* Xc = Xs[l & 1];
    for (int idx = tid; idx < rows_valid * H; idx += CT) {
      const int r = idx / H, c = idx - r * H;
      Xc[r * XP + c] = (Xc[r * XP + c] - mean[c]) * rstd[c];
    }
    __syncthreads();
    CSTAMP(0, 3 + 5 * l);
    float acc[NTW][4];
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const int nt_used = max(0, min(NTW, (nout + 7) / 8 - wc * NTW));
    {
      const float* Aw = Xc + wr * 16 * XP;
      const float* Bw = Ws[l & 1] + wc * NTW * 8 * XP;
      if (FAST) {
        const int nt_total = (nout + 7) / 8, full = nt_total / NTW, rem = nt_total % NTW;   // constants once unrolled
        if (wc < full) warp_gemm2<NTW, true, true>(Aw, XP, Bw, XP, (H + 7) / 8, NTW, acc, lane);
        else if (wc == full && rem > 0) warp_gemm2<NTW, true, true>(Aw, XP, Bw, XP, (H + 7) / 8, rem, acc, lane);
      } else {
        warp_gemm2<NTW, true, true>(Aw, XP, Bw, XP, (H + 7) / 8, nt_used, acc, lane);
      }
    }
    CSTAMP(0, 4 + 5 * l);
    // ---- epilogue: bias + ReLU -> global a_{l+2}, the next operand tile, fp64 column sums of the valid rows
    float* out = p.aout[l] + ((int64_t)arm * B + row0) * nout;
    float* Xn = Xs[(l + 1) & 1];
    const int ra = wr * 16 + g, rb = ra + 8;
    const bool va = ra < rows_valid, vb = rb < rows_valid;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt < nt_used) {
        const int c = (wc * NTW + nt) * 8 + 2 * tig;
        const float b0 = bias[l * 128 + c], b1 = bias[l * 128 + c + 1];
        const float v00 = fmaxf(acc[nt][0] + b0, 0.f), v01 = fmaxf(acc[nt][1] + b1, 0.f);
        const float v10 = fmaxf(acc[nt][2] + b0, 0.f), v11 = fmaxf(acc[nt][3] + b1, 0.f);
        if (va) { if (c < nout) out[(int64_t)ra * nout + c] = v00; if (c + 1 < nout) out[(int64_t)ra * nout + c + 1] = v01; }
        if (vb) { if (c < nout) out[(int64_t)rb * nout + c] = v10; if (c + 1 < nout) out[(int64_t)rb * nout + c + 1] = v11; }
        if (l < 3) {
          Xn[ra * XP + c] = (va && c < nout) ? v00 : 0.f;
          Xn[ra * XP + c + 1] = (va && c + 1 < nout) ? v01 : 0.f;
          Xn[rb * XP + c] = (vb && c < nout) ? v10 : 0.f;
          Xn[rb * XP + c + 1] = (vb && c + 1 < nout) ? v11 : 0.f;
        }
        const double a0 = va ? (double)v00 : 0.0, a1 = va ? (double)v01 : 0.0;
        const double c0 = vb ? (double)v10 : 0.0, c1 = vb ? (double)v11 : 0.0;
        const double s0 = group_sum_d(a0 + c0), s1 = group_sum_d(a1 + c1);
        const double q0 = group_sum_d(a0 * a0 + c0 * c0), q1 = group_sum_d(a1 * a1 + c1 * c1);
        if (g == 0) {
          red[(wr * 2 + 0) * 128 + c] = s0; red[(wr * 2 + 0) * 128 + c + 1] = s1;
          red[(wr * 2 + 1) * 128 + c] = q0; red[(wr * 2 + 1) * 128 + c + 1] = q1;
        }
      }
    }
    __syncthreads();                       // Ws[l&1] / Xs[l&1] released, red complete
    CSTAMP(0, 5 + 5 * l);
    for (int j = tid; j < 256; j += CT) {
      const int which = j >> 7, c = j & 127;
      if (c < nout) {
        double sacc = 0.0;
        for (int w = 0; w < MT; ++w) sacc += red[(w * 2 + which) * 128 + c];
        atomicAdd(p.sums + (int64_t)((l + 1) * p.A + arm) * 256 + which * 128 + c, sacc);
      }
    }
    if (l + 2 < 4) {
      if (l + 2 == 3) {                    // fc5 is [L][H]: the rows beyond L keep stale (finite) fc3 weights, never stored
        async_tile(Ws[l & 1], XP, par + p.offW[3], H, L, H, tid, CT);
      } else {
        async_tile(Ws[l & 1], XP, par + p.offW[l + 2], H, H, H, tid, CT);
      }
      cp_async_commit();
    }
    if (l < 3) grid_barrier(p.bar + l, nctas);
    CSTAMP(0, 6 + 5 * l);
  }
}

// ---------------------------------------------------------------------------------------------
// The same chain for batches that need more row tiles than there are SMs (A * ceil(B / 80) > #SM: three or more arms at
// B = 5000, B = 16384): the grid is still one co-resident wave, and a CTA walks over SEVERAL row tiles of its arm inside
// every layer (tile j of CTA x: row block x + j * gridDim.x).  Its tiles cannot all stay in shared memory, so the layer's
// input tile comes back from global memory (it was written there for the backward pass anyway, by this same CTA) into a
// two-deep ring -- tile j + 1 is in flight while tile j is multiplied -- and the column sums of all its tiles are added
// up in shared memory before the one round of atomics per layer.
// ---------------------------------------------------------------------------------------------
template <int MT, int NWC, int HC, int LC>
__global__ void __launch_bounds__(32 * MT * NWC) enc_chain_fwd_multi_kernel(const EncChainArgs p) {
  extern __shared__ __align__(16) float smem[];
  pdl_trigger();
  constexpr int CR = 16 * MT, CT = 32 * MT * NWC, NTW = 16 / NWC;
  constexpr bool FAST = HC > 0;
  const int XP = FAST ? (((HC + 7) & ~7) + 4) : p.XP, Hp = FAST ? ((HC + 7) & ~7) : p.Hp;
  float* Ws0 = smem;                       // [Hp][XP]  W natural [out][in]
  float* Ws1 = Ws0 + Hp * XP;
  float* Xs0 = Ws1 + Hp * XP;              // [CR][XP]
  float* Xs1 = Xs0 + CR * XP;
  float* bias = Xs1 + CR * XP;             // [4][128]
  float* mean = bias + 4 * 128;            // [128]
  float* rstd = mean + 128;                // [128]
  double* red = reinterpret_cast<double*>(rstd + 128);   // [MT][2][128]
  double* tot = red + MT * 2 * 128;                      // [2][128] sums over this CTA's tiles of the layer
  const int arm = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int warp = tid >> 5, wr = warp % MT, wc = warp / MT;
  const int g = lane >> 2, tig = lane & 3;
  const int H = FAST ? HC : p.H, L = FAST ? LC : p.L, B = p.B;
  const float* par = p.params + (int64_t)arm * p.p_arm_stride;
  const unsigned int nctas = gridDim.x * gridDim.y;
  const int ntiles = (B + CR - 1) / CR;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  for (int idx = tid; idx < 2 * Hp * XP + 2 * CR * XP; idx += CT) smem[idx] = 0.f;
  for (int idx = tid; idx < 4 * 128; idx += CT) {
    const int l = idx >> 7, j = idx & 127;
    const int nout = l < 3 ? H : L;
    bias[idx] = j < nout ? par[p.offB[l] + j] : 0.f;
  }
  __syncthreads();
  const SmemPair Ws{Ws0, (int)(Ws1 - Ws0)};
  const SmemPair Xs{Xs0, (int)(Xs1 - Xs0)};
  auto load_tile = [&](int l, int j, float* dst) {      // input tile j of layer l (a1 / a2 / a3 / a4), valid rows only
    const int row0 = ((int)blockIdx.x + j * (int)gridDim.x) * CR;
    const float* src = (l == 0 ? p.a1 : p.aout[l - 1]) + ((int64_t)arm * B + row0) * H;
    async_tile(dst, XP, src, H, min(CR, B - row0), H, tid, CT);
  };
  async_tile(Ws0, XP, par + p.offW[0], H, H, H, tid, CT);
  if (mine > 0) load_tile(0, 0, Xs0);
  cp_async_commit();
  async_tile(Ws1, XP, par + p.offW[1], H, H, H, tid, CT);
  cp_async_commit();

  for (int l = 0; l < 4; ++l) {
    const int nout = l < 3 ? H : L;
    if (tid < H) {
      const double* sums = p.sums + (int64_t)(l * p.A + arm) * 256;
      const double s1 = __ldcg(sums + tid), s2 = __ldcg(sums + 128 + tid);
      const double m = s1 / (double)B;
      double var = s2 / (double)B - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)p.eps));
      mean[tid] = mf;
      rstd[tid] = rf;
      if (blockIdx.x == 0) {
        p.bn_mean[(l * p.A + arm) * 128 + tid] = mf;
        p.bn_rstd[(l * p.A + arm) * 128 + tid] = rf;
      }
    }
    for (int j = tid; j < 256; j += CT) tot[j] = 0.0;
    const int nt_used = max(0, min(NTW, (nout + 7) / 8 - wc * NTW));
    for (int j = 0; j < mine; ++j) {
      const int row0 = ((int)blockIdx.x + j * (int)gridDim.x) * CR;
      const int rows_valid = min(CR, B - row0);
      cp_async_wait<0>();                  // tile j and the weights of this layer have landed
      __syncthreads();                     // ... for everyone; the other tile buffer is free (tile j - 1 is finished)
      if (j + 1 < mine) {
        load_tile(l, j + 1, Xs[(j + 1) & 1]);
        cp_async_commit();
      }
      float* Xc = Xs[j & 1];
      for (int idx = tid; idx < rows_valid * H; idx += CT) {
        const int r = idx / H, c = idx - r * H;
        Xc[r * XP + c] = (Xc[r * XP + c] - mean[c]) * rstd[c];
      }
      __syncthreads();
      float acc[NTW][4];
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
      if (nt_used == NTW)                  // (a literal count folds the per-tile guards of the MMA loop away)
        warp_gemm2<NTW, true, true>(Xc + wr * 16 * XP, XP, Ws[l & 1] + wc * NTW * 8 * XP, XP, (H + 7) / 8, NTW, acc, lane);
      else if (nt_used > 0)
        warp_gemm2<NTW, true, true>(Xc + wr * 16 * XP, XP, Ws[l & 1] + wc * NTW * 8 * XP, XP, (H + 7) / 8, nt_used, acc, lane);
      float* out = p.aout[l] + ((int64_t)arm * B + row0) * nout;
      const int ra = wr * 16 + g, rb = ra + 8;
      const bool va = ra < rows_valid, vb = rb < rows_valid;
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        if (nt < nt_used) {
          const int c = (wc * NTW + nt) * 8 + 2 * tig;
          const float b0 = bias[l * 128 + c], b1 = bias[l * 128 + c + 1];
          const float v00 = fmaxf(acc[nt][0] + b0, 0.f), v01 = fmaxf(acc[nt][1] + b1, 0.f);
          const float v10 = fmaxf(acc[nt][2] + b0, 0.f), v11 = fmaxf(acc[nt][3] + b1, 0.f);
          if (va) { if (c < nout) out[(int64_t)ra * nout + c] = v00; if (c + 1 < nout) out[(int64_t)ra * nout + c + 1] = v01; }
          if (vb) { if (c < nout) out[(int64_t)rb * nout + c] = v10; if (c + 1 < nout) out[(int64_t)rb * nout + c + 1] = v11; }
          const double a0 = va ? (double)v00 : 0.0, a1 = va ? (double)v01 : 0.0;
          const double c0 = vb ? (double)v10 : 0.0, c1 = vb ? (double)v11 : 0.0;
          const double s0 = group_sum_d(a0 + c0), s1 = group_sum_d(a1 + c1);
          const double q0 = group_sum_d(a0 * a0 + c0 * c0), q1 = group_sum_d(a1 * a1 + c1 * c1);
          if (g == 0) {
            red[(wr * 2 + 0) * 128 + c] = s0; red[(wr * 2 + 0) * 128 + c + 1] = s1;
            red[(wr * 2 + 1) * 128 + c] = q0; red[(wr * 2 + 1) * 128 + c + 1] = q1;
          }
        }
      }
      __syncthreads();                     // red complete; everyone is done with this tile
      for (int jj = tid; jj < 256; jj += CT) {
        if ((jj & 127) < nout) {
          double sacc = 0.0;
          for (int w = 0; w < MT; ++w) sacc += red[(w * 2 + (jj >> 7)) * 128 + (jj & 127)];
          tot[jj] += sacc;                 // (thread jj owns tot[jj])
        }
      }
    }
    __syncthreads();                       // every warp has finished the layer: its weight buffer is free
    for (int jj = tid; jj < 256; jj += CT)
      if ((jj & 127) < nout && mine > 0)
        atomicAdd(p.sums + (int64_t)((l + 1) * p.A + arm) * 256 + jj, tot[jj]);
    if (l + 2 < 4) {
      if (l + 2 == 3) async_tile(Ws[l & 1], XP, par + p.offW[3], H, L, H, tid, CT);      // fc5 is [L][H]; stale rows are never stored
      else async_tile(Ws[l & 1], XP, par + p.offW[l + 2], H, H, H, tid, CT);
    }
    if (l < 3 && mine > 0) load_tile(l + 1, 0, Xs0);   // this CTA's own rows of the layer it has just written
    cp_async_commit();
    if (l < 3) grid_barrier(p.bar + l, nctas);
  }
}

// =============================================================================================
// encoder middle, backward: for l = 5..1 (array index 4..0)
//   g_l (gradient wrt the BatchNorm output of layer l) -> BatchNorm backward (needs the batch means of g and g*n:
//   grid-wide fp64 sums) -> ReLU mask -> delta_l (stored for the weight-gradient kernel) -> g_{l-1} = delta_l . W_l,
//   whose own BatchNorm sums are accumulated in the epilogue.  The gradient tile never leaves shared memory.
// =============================================================================================
struct EncBwdArgs {
  int XP, WPB, Hp;
  const float* params; int64_t p_arm_stride;
  int64_t offW[5];               // fc1..fc5 weights (index 0 unused)
  int A, B, H, L;
  const float* g_xlow;           // [A][B][L]
  const float* act[5];           // a1..a5 (post-ReLU, pre-BN)
  float* delta[5];               // delta1..delta5
  double* sums;                  // acc_bwd: layer l at (l * A + arm) * 256: sum g | sum g*n
  const float* bn_mean; const float* bn_rstd;   // [5][A][128]
  unsigned int* bar;
};

template <int MT, int NWC, bool SPLIT, int HC, int LC>
__global__ void __launch_bounds__(32 * MT * NWC) enc_chain_bwd_kernel(const EncBwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  pdl_trigger();
  CSTAMP(1, 0);
  constexpr int CR = 16 * MT, CT = 32 * MT * NWC, NTW = 16 / NWC;
  constexpr bool FAST = HC > 0;
  constexpr int HPC = (HC + 7) & ~7;
  const int XP = FAST ? HPC + 4 : p.XP, Hp = FAST ? HPC : p.Hp;
  const int WPB = FAST ? (((HPC & 31) == 8 || (HPC & 31) == 24) ? HPC : p.WPB) : p.WPB;
  float* Ws0 = smem;                       // [Hp][WPB]  W_l natural: row j (out), col i (in)
  float* Ws1 = Ws0 + Hp * WPB;
  float* Gs = Ws1 + Hp * WPB;              // [CR][XP]   g_l, then delta_l in place, then g_{l-1}
  float* As0 = Gs + CR * XP;               // [CR][XP]   activation tiles a_l / a_{l-1}
  float* As1 = As0 + CR * XP;
  float* c1 = As1 + CR * XP;               // [128] each: mean_b(g), mean_b(g*n), mean/rstd of layer l and of layer l-1
  float* c2 = c1 + 128;
  float* mo = c2 + 128;
  float* ro = mo + 128;
  float* mi = ro + 128;
  float* ri = mi + 128;
  double* red = reinterpret_cast<double*>(ri + 128);     // [MT][2][128]
  const int arm = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int warp = tid >> 5, wr = warp % MT, wc = warp / MT;
  const int g = lane >> 2, tig = lane & 3;
  const int H = FAST ? HC : p.H, L = FAST ? LC : p.L, B = p.B;
  const int row0 = blockIdx.x * CR;
  const int rows_valid = min(CR, B - row0);
  const float* par = p.params + (int64_t)arm * p.p_arm_stride;
  const int64_t rbase = (int64_t)arm * B + row0;
  const unsigned int nctas = gridDim.x * gridDim.y;

  for (int idx = tid; idx < 2 * Hp * WPB + 3 * CR * XP; idx += CT) smem[idx] = 0.f;
  __syncthreads();
  // group P0: g_xlow, a5, W5 (fc5: [L][H]); group P1: a4, W4
  async_tile(Gs, XP, p.g_xlow + rbase * L, L, rows_valid, L, tid, CT);
  async_tile(As0, XP, p.act[4] + rbase * L, L, rows_valid, L, tid, CT);
  async_tile(Ws0, WPB, par + p.offW[4], H, L, H, tid, CT);
  cp_async_commit();
  async_tile(As1, XP, p.act[3] + rbase * H, H, rows_valid, H, tid, CT);
  async_tile(Ws1, WPB, par + p.offW[3], H, H, H, tid, CT);
  cp_async_commit();

  const SmemPair Ws{Ws0, (int)(Ws1 - Ws0)};
  float* As[2] = {As0, As1};
  constexpr int UNR = FAST ? 5 : 1;
#pragma unroll UNR
  for (int it = 0; it < 5; ++it) {
    const int l = 4 - it;
    const int nout = l == 4 ? L : H;
    // ---- constants of the BatchNorm backward of layer l (sums complete: head kernel for l == 4, grid barrier otherwise)
    if (tid < nout) {
      const double* sums = p.sums + (int64_t)(l * p.A + arm) * 256;
      c1[tid] = (float)(__ldcg(sums + tid) / (double)B);
      c2[tid] = (float)(__ldcg(sums + 128 + tid) / (double)B);
      mo[tid] = p.bn_mean[(l * p.A + arm) * 128 + tid];
      ro[tid] = p.bn_rstd[(l * p.A + arm) * 128 + tid];
    }
    if (l > 0 && tid < H) {
      mi[tid] = p.bn_mean[((l - 1) * p.A + arm) * 128 + tid];
      ri[tid] = p.bn_rstd[((l - 1) * p.A + arm) * 128 + tid];
    }
    if (it < 4) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    CSTAMP(1, 1 + 5 * it);
    // ---- delta_l = bn_bwd(g_l) * relu'(a_l): in place in Gs and to global
    const float* Ac = As[it & 1];
    float* dout = p.delta[l] + rbase * nout;
    for (int idx = tid; idx < rows_valid * nout; idx += CT) {
      const int r = idx / nout, j = idx - r * nout;
      const float a = Ac[r * XP + j];
      const float n = (a - mo[j]) * ro[j];
      const float gg = ro[j] * (Gs[r * XP + j] - c1[j] - n * c2[j]);
      const float d = a > 0.f ? gg : 0.f;
      Gs[r * XP + j] = d;
      dout[(int64_t)r * nout + j] = d;
    }
    if (l == 0) break;
    cp_async_wait<0>();                    // a_{l-1} (needed by the epilogue) has landed
    __syncthreads();
    CSTAMP(1, 2 + 5 * it);
    float acc[NTW][4];
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const int nt_used = max(0, min(NTW, (H + 7) / 8 - wc * NTW));
    {
      const float* Aw = Gs + wr * 16 * XP;
      const float* Bw = Ws[it & 1] + wc * NTW * 8;
      if (FAST) {
        const int nt_total = (H + 7) / 8, full = nt_total / NTW, rem = nt_total % NTW;      // constants
        if (wc < full) warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (nout + 7) / 8, NTW, acc, lane);
        else if (wc == full && rem > 0) warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (nout + 7) / 8, rem, acc, lane);
      } else {
        warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (nout + 7) / 8, nt_used, acc, lane);
      }
    }
    __syncthreads();                       // all warps have read delta_l before g_{l-1} overwrites it
    CSTAMP(1, 3 + 5 * it);
    // ---- epilogue: g_{l-1} -> Gs; fp64 sums of g and g * n_{l-1} over the valid rows
    const float* An = As[(it + 1) & 1];
    const int ra = wr * 16 + g, rb = ra + 8;
    const bool va = ra < rows_valid, vb = rb < rows_valid;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt < nt_used) {
        const int c = (wc * NTW + nt) * 8 + 2 * tig;
        const bool c0ok = c < H, c1ok = c + 1 < H;
        double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
        if (c0ok) {
          Gs[ra * XP + c] = va ? acc[nt][0] : 0.f;
          Gs[rb * XP + c] = vb ? acc[nt][2] : 0.f;
          if (va) { const double n = (double)((An[ra * XP + c] - mi[c]) * ri[c]); s0 += (double)acc[nt][0]; q0 += (double)acc[nt][0] * n; }
          if (vb) { const double n = (double)((An[rb * XP + c] - mi[c]) * ri[c]); s0 += (double)acc[nt][2]; q0 += (double)acc[nt][2] * n; }
        }
        if (c1ok) {
          Gs[ra * XP + c + 1] = va ? acc[nt][1] : 0.f;
          Gs[rb * XP + c + 1] = vb ? acc[nt][3] : 0.f;
          if (va) { const double n = (double)((An[ra * XP + c + 1] - mi[c + 1]) * ri[c + 1]); s1 += (double)acc[nt][1]; q1 += (double)acc[nt][1] * n; }
          if (vb) { const double n = (double)((An[rb * XP + c + 1] - mi[c + 1]) * ri[c + 1]); s1 += (double)acc[nt][3]; q1 += (double)acc[nt][3] * n; }
        }
        s0 = group_sum_d(s0); s1 = group_sum_d(s1); q0 = group_sum_d(q0); q1 = group_sum_d(q1);
        if (g == 0) {
          red[(wr * 2 + 0) * 128 + c] = s0; red[(wr * 2 + 0) * 128 + c + 1] = s1;
          red[(wr * 2 + 1) * 128 + c] = q0; red[(wr * 2 + 1) * 128 + c + 1] = q1;
        }
      }
    }
    __syncthreads();
    CSTAMP(1, 4 + 5 * it);
    for (int j = tid; j < 256; j += CT) {
      const int which = j >> 7, c = j & 127;
      if (c < H) {
        double sacc = 0.0;
        for (int w = 0; w < MT; ++w) sacc += red[(w * 2 + which) * 128 + c];
        atomicAdd(p.sums + (int64_t)((l - 1) * p.A + arm) * 256 + which * 128 + c, sacc);
      }
    }
    // ---- stage a_{l-2} and W_{l-2} into the buffers of this iteration (layer 0 has no data gradient: no weights)
    if (l - 2 >= 0) {
      async_tile(As[it & 1], XP, p.act[l - 2] + rbase * H, H, rows_valid, H, tid, CT);
      if (l - 2 >= 1) async_tile(Ws[it & 1], WPB, par + p.offW[l - 2], H, H, H, tid, CT);
    }
    cp_async_commit();
    grid_barrier(p.bar + it, nctas);
    CSTAMP(1, 5 + 5 * it);
  }
  CSTAMP(1, 26);
}

// ---------------------------------------------------------------------------------------------
// Backward chain over several row tiles per CTA (see enc_chain_fwd_multi_kernel).  The gradient tile cannot stay in shared
// memory between layers here: g_{l-1} goes to a global scratch buffer [A][B][H] and comes back at the next layer (always
// this CTA's own rows).  Per tile: (g_l, a_l) and a_{l-1} arrive in two cp.async groups, so that the BatchNorm/ReLU
// backward of the tile runs while the second group is still in flight.
// ---------------------------------------------------------------------------------------------
template <int MT, int NWC, bool SPLIT, int HC, int LC>
__global__ void __launch_bounds__(32 * MT * NWC) enc_chain_bwd_multi_kernel(const EncBwdArgs p, float* __restrict__ gscr) {
  extern __shared__ __align__(16) float smem[];
  pdl_trigger();
  constexpr int CR = 16 * MT, CT = 32 * MT * NWC, NTW = 16 / NWC;
  constexpr bool FAST = HC > 0;
  constexpr int HPC = (HC + 7) & ~7;
  const int XP = FAST ? HPC + 4 : p.XP, Hp = FAST ? HPC : p.Hp;
  const int WPB = FAST ? (((HPC & 31) == 8 || (HPC & 31) == 24) ? HPC : p.WPB) : p.WPB;
  float* Ws0 = smem;                       // [Hp][WPB]  W_l natural: row j (out), col i (in)
  float* Ws1 = Ws0 + Hp * WPB;
  float* Gs = Ws1 + Hp * WPB;              // [CR][XP]   g_l, then delta_l in place
  float* Ac = Gs + CR * XP;                // [CR][XP]   a_l tile
  float* An = Ac + CR * XP;                // [CR][XP]   a_{l-1} tile
  float* c1 = An + CR * XP;                // [128] each: mean_b(g), mean_b(g*n), mean/rstd of layer l and of layer l-1
  float* c2 = c1 + 128;
  float* mo = c2 + 128;
  float* ro = mo + 128;
  float* mi = ro + 128;
  float* ri = mi + 128;
  double* red = reinterpret_cast<double*>(ri + 128);     // [MT][2][128]
  double* tot = red + MT * 2 * 128;                      // [2][128]
  const int arm = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int warp = tid >> 5, wr = warp % MT, wc = warp / MT;
  const int g = lane >> 2, tig = lane & 3;
  const int H = FAST ? HC : p.H, L = FAST ? LC : p.L, B = p.B;
  const float* par = p.params + (int64_t)arm * p.p_arm_stride;
  const unsigned int nctas = gridDim.x * gridDim.y;
  const int ntiles = (B + CR - 1) / CR;
  const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  for (int idx = tid; idx < 2 * Hp * WPB + 3 * CR * XP; idx += CT) smem[idx] = 0.f;
  __syncthreads();
  async_tile(Ws0, WPB, par + p.offW[4], H, L, H, tid, CT);       // fc5: [L][H]
  cp_async_commit();
  async_tile(Ws1, WPB, par + p.offW[3], H, H, H, tid, CT);
  cp_async_commit();
  const SmemPair Ws{Ws0, (int)(Ws1 - Ws0)};

  for (int it = 0; it < 5; ++it) {
    const int l = 4 - it;
    const int nout = l == 4 ? L : H;
    if (tid < nout) {
      const double* sums = p.sums + (int64_t)(l * p.A + arm) * 256;
      c1[tid] = (float)(__ldcg(sums + tid) / (double)B);
      c2[tid] = (float)(__ldcg(sums + 128 + tid) / (double)B);
      mo[tid] = p.bn_mean[(l * p.A + arm) * 128 + tid];
      ro[tid] = p.bn_rstd[(l * p.A + arm) * 128 + tid];
    }
    if (l > 0 && tid < H) {
      mi[tid] = p.bn_mean[((l - 1) * p.A + arm) * 128 + tid];
      ri[tid] = p.bn_rstd[((l - 1) * p.A + arm) * 128 + tid];
    }
    for (int j = tid; j < 256; j += CT) tot[j] = 0.0;
    const int nt_used = max(0, min(NTW, (H + 7) / 8 - wc * NTW));
    for (int j = 0; j < mine; ++j) {
      const int row0 = ((int)blockIdx.x + j * (int)gridDim.x) * CR;
      const int rows_valid = min(CR, B - row0);
      const int64_t rbase = (int64_t)arm * B + row0;
      // (the previous tile's epilogue ended with a __syncthreads: the three tile buffers are free)
      async_tile(Gs, XP, (l == 4 ? p.g_xlow : gscr) + rbase * nout, nout, rows_valid, nout, tid, CT);
      async_tile(Ac, XP, p.act[l] + rbase * nout, nout, rows_valid, nout, tid, CT);
      cp_async_commit();
      if (l > 0) async_tile(An, XP, p.act[l - 1] + rbase * H, H, rows_valid, H, tid, CT);
      cp_async_commit();
      cp_async_wait<1>();                  // g_l, a_l (and this layer's weights, issued a whole layer ago)
      __syncthreads();
      float* dout = p.delta[l] + rbase * nout;
      for (int idx = tid; idx < rows_valid * nout; idx += CT) {
        const int r = idx / nout, jj = idx - r * nout;
        const float a = Ac[r * XP + jj];
        const float n = (a - mo[jj]) * ro[jj];
        const float gg = ro[jj] * (Gs[r * XP + jj] - c1[jj] - n * c2[jj]);
        const float d = a > 0.f ? gg : 0.f;
        Gs[r * XP + jj] = d;
        dout[(int64_t)r * nout + jj] = d;
      }
      if (l == 0) { __syncthreads(); continue; }
      // rows beyond the batch of a partial tile must not carry another tile's values into the product's valid rows:
      // they do not (rows are independent), and their results are neither stored nor summed
      cp_async_wait<0>();
      __syncthreads();
      float acc[NTW][4];
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
      {
        const float* Aw = Gs + wr * 16 * XP;
        const float* Bw = Ws[it & 1] + wc * NTW * 8;
        if (nt_used == NTW) warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (nout + 7) / 8, NTW, acc, lane);
        else if (nt_used > 0) warp_gemm2<NTW, false, SPLIT>(Aw, XP, Bw, WPB, (nout + 7) / 8, nt_used, acc, lane);
      }
      float* gout = gscr + rbase * H;
      const int ra = wr * 16 + g, rb = ra + 8;
      const bool va = ra < rows_valid, vb = rb < rows_valid;
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        if (nt < nt_used) {
          const int c = (wc * NTW + nt) * 8 + 2 * tig;
          const bool c0ok = c < H, c1ok = c + 1 < H;
          double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
          if (c0ok) {
            if (va) { gout[(int64_t)ra * H + c] = acc[nt][0]; const double n = (double)((An[ra * XP + c] - mi[c]) * ri[c]); s0 += (double)acc[nt][0]; q0 += (double)acc[nt][0] * n; }
            if (vb) { gout[(int64_t)rb * H + c] = acc[nt][2]; const double n = (double)((An[rb * XP + c] - mi[c]) * ri[c]); s0 += (double)acc[nt][2]; q0 += (double)acc[nt][2] * n; }
          }
          if (c1ok) {
            if (va) { gout[(int64_t)ra * H + c + 1] = acc[nt][1]; const double n = (double)((An[ra * XP + c + 1] - mi[c + 1]) * ri[c + 1]); s1 += (double)acc[nt][1]; q1 += (double)acc[nt][1] * n; }
            if (vb) { gout[(int64_t)rb * H + c + 1] = acc[nt][3]; const double n = (double)((An[rb * XP + c + 1] - mi[c + 1]) * ri[c + 1]); s1 += (double)acc[nt][3]; q1 += (double)acc[nt][3] * n; }
          }
          s0 = group_sum_d(s0); s1 = group_sum_d(s1); q0 = group_sum_d(q0); q1 = group_sum_d(q1);
          if (g == 0) {
            red[(wr * 2 + 0) * 128 + c] = s0; red[(wr * 2 + 0) * 128 + c + 1] = s1;
            red[(wr * 2 + 1) * 128 + c] = q0; red[(wr * 2 + 1) * 128 + c + 1] = q1;
          }
        }
      }
      __syncthreads();                     // red complete, Gs / Ac / An free
      for (int jj = tid; jj < 256; jj += CT) {
        if ((jj & 127) < H) {
          double sacc = 0.0;
          for (int w = 0; w < MT; ++w) sacc += red[(w * 2 + (jj >> 7)) * 128 + (jj & 127)];
          tot[jj] += sacc;
        }
      }
      __syncthreads();                     // red is rewritten by the next tile
    }
    if (l == 0) break;
    for (int jj = tid; jj < 256; jj += CT)
      if ((jj & 127) < H && mine > 0) atomicAdd(p.sums + (int64_t)((l - 1) * p.A + arm) * 256 + jj, tot[jj]);
    // W_{l-2} into the weight buffer of this iteration (layer 0 has no data gradient: no weights)
    if (l - 2 >= 1) async_tile(Ws[it & 1], WPB, par + p.offW[l - 2], H, H, H, tid, CT);
    cp_async_commit();
    grid_barrier(p.bar + it, nctas);
  }
}

}  // namespace

static int b_pitch2(int n) {
  int p = (n + 7) & ~7;
  while ((p & 31) != 8 && (p & 31) != 24) p += 8;
  return p;
}

static void fill_chain(ChainArgs& c, const float* params, int64_t p_arm_stride, const int64_t* off, int B, int H, int L) {
  memset(&c, 0, sizeof(c));
  c.params = params; c.p_arm_stride = p_arm_stride; c.B = B; c.H = H; c.L = L;
  c.Hp = (H + 7) & ~7;
  c.XP = c.Hp + 4;
  c.WPB = b_pitch2(c.Hp);
  for (int l = 0; l < 4; ++l) {
    c.offW[l] = off[FC7_W + 2 * l];
    c.offB[l] = off[FC7_B + 2 * l];
  }
}

static int chain_mt(int A, int B) {
  const int t4 = (B + 63) / 64 * A, t5 = (B + 79) / 80 * A;
  return (t4 > 148 && t5 <= 148) ? 5 : 4;
}

int launch_dec_chain_fwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* h6, float* const hout[4], int split3, cudaStream_t s) {
  ChainArgs c;
  fill_chain(c, params, p_arm_stride, off, B, H, L);
  c.h6 = h6;
  for (int l = 0; l < 4; ++l) c.hout[l] = hout[l];
  const int mt = chain_mt(A, B), cr = 16 * mt;
  const size_t smem = (size_t)(2 * c.Hp * c.XP + 2 * cr * c.XP + 4 * 128) * 4;
#define CHAIN_LAUNCH(MTV, SP)                                                                                        \
  do {                                                                                                              \
    if (H == 100 && L == 10) {                                                                                      \
      MVAE_CUDA(cudaFuncSetAttribute(dec_chain_fwd_kernel<MTV, SP, 100, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      launch_pdl(dec_chain_fwd_kernel<MTV, SP, 100, 10>, dim3((B + cr - 1) / cr, A), dim3(128 * MTV), smem, s, c);                  \
    } else {                                                                                                        \
      MVAE_CUDA(cudaFuncSetAttribute(dec_chain_fwd_kernel<MTV, SP, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      launch_pdl(dec_chain_fwd_kernel<MTV, SP, 0, 0>, dim3((B + cr - 1) / cr, A), dim3(128 * MTV), smem, s, c);                     \
    }                                                                                                               \
  } while (0)
  if (mt == 5 && split3) CHAIN_LAUNCH(5, true);
  else if (mt == 5) CHAIN_LAUNCH(5, false);
  else if (split3) CHAIN_LAUNCH(4, true);
  else CHAIN_LAUNCH(4, false);
#undef CHAIN_LAUNCH
  MVAE_LAUNCH_CHECK();
  return 0;
}

int launch_dec_chain_bwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* g10, const float* const act[4], float* const delta[4], float* g6, int split3,
                         cudaStream_t s) {
  ChainArgs c;
  fill_chain(c, params, p_arm_stride, off, B, H, L);
  c.g10 = g10; c.g6 = g6;
  for (int l = 0; l < 4; ++l) { c.act[l] = act[l]; c.delta[l] = delta[l]; }
  const int mt = chain_mt(A, B), cr = 16 * mt;
  const size_t smem = (size_t)(2 * c.Hp * c.WPB + 3 * cr * c.XP) * 4;
#define CHAIN_LAUNCH(MTV, SP)                                                                                        \
  do {                                                                                                              \
    if (H == 100 && L == 10) {                                                                                      \
      MVAE_CUDA(cudaFuncSetAttribute(dec_chain_bwd_kernel<MTV, SP, 100, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      launch_pdl(dec_chain_bwd_kernel<MTV, SP, 100, 10>, dim3((B + cr - 1) / cr, A), dim3(128 * MTV), smem, s, c);                  \
    } else {                                                                                                        \
      MVAE_CUDA(cudaFuncSetAttribute(dec_chain_bwd_kernel<MTV, SP, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      launch_pdl(dec_chain_bwd_kernel<MTV, SP, 0, 0>, dim3((B + cr - 1) / cr, A), dim3(128 * MTV), smem, s, c);                     \
    }                                                                                                               \
  } while (0)
  if (mt == 5 && split3) CHAIN_LAUNCH(5, true);
  else if (mt == 5) CHAIN_LAUNCH(5, false);
  else if (split3) CHAIN_LAUNCH(4, true);
  else CHAIN_LAUNCH(4, false);
#undef CHAIN_LAUNCH
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae

namespace mvae {

// cooperative-launch capability and SM count of the current device (a process may drive several)
static int coop_sms() {
  static int cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 0;
  if (!cache[dev]) {
    int coop = 0, nsm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = coop ? nsm : -1;
  }
  return cache[dev] > 0 ? cache[dev] : 0;
}

int enc_chain_fwd_is_single(int A, int B, int H, int L) {
  const int nsm = coop_sms();
  if (!nsm || H > 128 || L > 64 || H % 4 != 0 || A > nsm) return 0;
  return (int64_t)((B + 79) / 80) * A <= nsm ? 1 : 0;
}

int launch_enc_chain_fwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* a1, float* const aout[4], double* acc_fwd, float* bn_mean, float* bn_rstd, float eps,
                         const Fc1Deferred* fc1, cudaStream_t s) {
  const int nsm = coop_sms();
  if (!nsm || H > 128 || L > 64 || H % 4 != 0 || A > nsm) return 1;
  EncChainArgs c;
  memset(&c, 0, sizeof(c));
  c.Hp = (H + 7) & ~7;
  c.XP = c.Hp + 4;
  c.params = params; c.p_arm_stride = p_arm_stride;
  for (int l = 0; l < 4; ++l) {
    c.offW[l] = off[FC2_W + 2 * l];
    c.offB[l] = off[FC2_B + 2 * l];
    c.aout[l] = aout[l];
  }
  c.A = A; c.B = B; c.H = H; c.L = L;
  c.a1 = a1; c.a1_out = const_cast<float*>(a1);
  if (fc1 && fc1->valid) c.fx = *fc1;
  c.sums = acc_fwd;
  c.bn_mean = bn_mean; c.bn_rstd = bn_rstd;
  c.bar = reinterpret_cast<unsigned int*>(acc_fwd + acc_sync(A));
  c.eps = eps;
  constexpr int MT = 5, NWC = 4, CR = 16 * MT;
  const int tiles = (B + CR - 1) / CR;
  // one co-resident wave (grid barrier): one row tile per CTA while that fits, else several tiles per CTA
  const bool multi = (int64_t)tiles * A > nsm;
  MVAE_CHECK_ARG(!(multi && c.fx.valid), "internal: the multi-tile encoder chain takes no deferred fc1 fix-up");
  const int gx = multi ? nsm / A : tiles;
  const size_t smem = (size_t)(2 * c.Hp * c.XP + 2 * CR * c.XP + 4 * 128 + 256) * 4 + (size_t)MT * 2 * 128 * 8 + (multi ? 2 * 128 * 8 : 0);
  const bool fast = (H == 100 && L == 10);
  void* args[] = {(void*)&c};
  const void* fn = multi ? (fast ? (const void*)enc_chain_fwd_multi_kernel<MT, NWC, 100, 10> : (const void*)enc_chain_fwd_multi_kernel<MT, NWC, 0, 0>)
                         : (fast ? (const void*)enc_chain_fwd_kernel<MT, NWC, 100, 10> : (const void*)enc_chain_fwd_kernel<MT, NWC, 0, 0>);
  MVAE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MVAE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(gx, A), dim3(32 * MT * NWC), args, smem, s));
  tl_pdl = 1;                              // (never a PDL secondary itself, but a primary: the kernel triggers)
  MVAE_LAUNCH_CHECK();
  return 0;
}

int launch_enc_chain_bwd(const float* params, int64_t p_arm_stride, const int64_t* off, int A, int B, int H, int L,
                         const float* g_xlow, const float* const act[5], float* const delta[5], double* acc_bwd,
                         const float* bn_mean, const float* bn_rstd, float* g_scratch, int split3, cudaStream_t s) {
  const int nsm = coop_sms();
  if (!nsm || H > 128 || L > 64 || H % 4 != 0 || A > nsm) return 1;
  EncBwdArgs c;
  memset(&c, 0, sizeof(c));
  c.Hp = (H + 7) & ~7;
  c.XP = c.Hp + 4;
  c.WPB = b_pitch2(c.Hp);
  c.params = params; c.p_arm_stride = p_arm_stride;
  for (int l = 0; l < 5; ++l) {
    c.offW[l] = off[FC1_W + 2 * l];
    c.act[l] = act[l];
    c.delta[l] = delta[l];
  }
  c.A = A; c.B = B; c.H = H; c.L = L;
  c.g_xlow = g_xlow;
  c.sums = acc_bwd;
  c.bn_mean = bn_mean; c.bn_rstd = bn_rstd;
  c.bar = reinterpret_cast<unsigned int*>(acc_bwd + accb_sync(A));
  constexpr int MT = 5, NWC = 4, CR = 16 * MT;
  const int tiles = (B + CR - 1) / CR;
  const bool multi = (int64_t)tiles * A > nsm;
  if (multi && !g_scratch) return 1;
  const int gx = multi ? nsm / A : tiles;
  const size_t smem = (size_t)(2 * c.Hp * c.WPB + 3 * CR * c.XP + 6 * 128) * 4 + (size_t)MT * 2 * 128 * 8 + (multi ? 2 * 128 * 8 : 0);
  const bool fast = (H == 100 && L == 10);
  const void* fn;
  if (multi) {
    fn = split3 ? (fast ? (const void*)enc_chain_bwd_multi_kernel<MT, NWC, true, 100, 10> : (const void*)enc_chain_bwd_multi_kernel<MT, NWC, true, 0, 0>)
                : (fast ? (const void*)enc_chain_bwd_multi_kernel<MT, NWC, false, 100, 10> : (const void*)enc_chain_bwd_multi_kernel<MT, NWC, false, 0, 0>);
  } else {
    fn = split3 ? (fast ? (const void*)enc_chain_bwd_kernel<MT, NWC, true, 100, 10> : (const void*)enc_chain_bwd_kernel<MT, NWC, true, 0, 0>)
                : (fast ? (const void*)enc_chain_bwd_kernel<MT, NWC, false, 100, 10> : (const void*)enc_chain_bwd_kernel<MT, NWC, false, 0, 0>);
  }
  void* args1[] = {(void*)&c};
  void* args2[] = {(void*)&c, (void*)&g_scratch};
  MVAE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MVAE_CUDA(cudaLaunchCooperativeKernel(fn, dim3(gx, A), dim3(32 * MT * NWC), multi ? args2 : args1, smem, s));
  tl_pdl = 1;
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae

#ifdef CHAIN_STAMPS
extern "C" int mvae_debug_chain_stamps(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, mvae::g_chain_stamps, sizeof(mvae::g_chain_stamps));
}
#endif
