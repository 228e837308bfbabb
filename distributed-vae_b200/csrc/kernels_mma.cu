// kernels_mma.cu — the narrow (<=128-wide) layers on warp-level tensor-core MMAs.
//
// These layers ([B,<=128] x [<=128,<=128]) are far too small for a TMA/tcgen05 pipeline (one
// 128-row tile has 13 K-steps) and are latency/issue-bound, not FLOP-bound: what matters is the
// number of issued instructions per cell.  mma.sync.m16n8k8 (TF32 operands, fp32 accumulate, run as
// the error-compensated 3xTF32 product so that results stay fp32-accurate: the categorical argmax
// must be bit-exact) needs ~12x fewer issue slots than the scalar-FMA kernels of kernels_rows.cu,
// which stay as the precision==3 ("fp32_simt") path.
//
//   dense_fwd_mma : out = act(W . bn(in) + b) + fp64 column sums for the next BatchNorm
//   dense_bwd_mma : delta = bn_bwd(g_out) * relu'(.) ; g_in = delta . W (+ sums for the next bn_bwd)
//   wgrad_mma     : dW = delta^T . in, db = delta^T . 1   (split over row chunks, fixed-order reduce)
#include "common.cuh"
#include "kernels.h"
#include "gemm_tc.h"

namespace mvae {

namespace {

constexpr int ROWS_PER_CTA = 64;   // 4 row groups of 16 rows x 4 column quarters = 16 warps
constexpr int MMA_THREADS = 512;
constexpr int NWC = 4;

// v = hi + lo: hi is v itself (mma.sync reads the upper 19 bits of a .tf32 operand register, i.e.
// truncates), lo = v - trunc(v) is exact in fp32 and is truncated to TF32 by the MMA in turn.
// (cvt.rna.tf32.f32 is a ~10-instruction emulation on sm_100; truncation keeps the split at 2 ops.)
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
// d += a . b with a = a_hi + a_lo, b = b_hi + b_lo (small terms first)
__device__ __forceinline__ void mma_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(d, al, bh);
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}

// pitch (floats) >= n with pitch % 32 in {8, 24}: B-fragment loads (k = tig, n = g) hit 32 banks
__host__ __device__ inline int b_pitch(int n) {
  int p = (n + 7) & ~7;
  while ((p & 31) != 8 && (p & 31) != 24) p += 8;
  return p;
}
// pitch (floats) >= k with pitch % 8 == 4: A-fragment loads (row = g, col = tig) hit 32 banks
__host__ __device__ inline int a_pitch(int k) { return ((k + 7) & ~7) + 4; }

// One warp: acc[NT][4] += A(16 x 8*ksteps) . B(8*ksteps x 8*NT)
//   A element (m,k): A_T ? As[k*ap + m] : As[m*ap + k]   (m relative to the warp's 16 rows)
//   B element (k,n): Bs[k*bp + n]
template <int NT, bool A_T, bool SPLIT = true>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ As, int ap, const float* __restrict__ Bs, int bp,
                                          int ksteps, int nt_used, float (&acc)[NT][4], int lane) {
  const int g = lane >> 2, tig = lane & 3;
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k0 = ks * 8;
    float av[4];
    if (!A_T) {
      av[0] = As[g * ap + k0 + tig];
      av[1] = As[(g + 8) * ap + k0 + tig];
      av[2] = As[g * ap + k0 + tig + 4];
      av[3] = As[(g + 8) * ap + k0 + tig + 4];
    } else {
      av[0] = As[(k0 + tig) * ap + g];
      av[1] = As[(k0 + tig) * ap + g + 8];
      av[2] = As[(k0 + tig + 4) * ap + g];
      av[3] = As[(k0 + tig + 4) * ap + g + 8];
    }
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nt_used) {
        const float b0 = Bs[(k0 + tig) * bp + nt * 8 + g];
        const float b1 = Bs[(k0 + tig + 4) * bp + nt * 8 + g];
        uint32_t bh[2], bl[2];
        if (SPLIT) {
          split_tf32(b0, bh[0], bl[0]);
          split_tf32(b1, bh[1], bl[1]);
          mma_3x(acc[nt], ah, al, bh, bl);
        } else {
          bh[0] = __float_as_uint(b0);
          bh[1] = __float_as_uint(b1);
          mma_tf32(acc[nt], ah, bh);
        }
      }
    }
  }
}

// Register staging of a [tile_rows x ncols] tile (row pitch ld): every load is issued before the first
// use so that a thread has NV independent requests in flight (the kernels are latency-bound).
template <int VEC, int NV, int NTHR>
__device__ __forceinline__ void tile_load(float (&v)[NV][VEC], const float* __restrict__ src, int64_t ld, int rows_valid,
                                          int ncols, int tile_rows, int tid) {
  const int cpr = (ncols + VEC - 1) / VEC;
  const int total = tile_rows * cpr;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int idx = tid + u * NTHR;
    const int r = idx / cpr, c = (idx - r * cpr) * VEC;
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[u][e] = 0.f;
    if (idx < total && r < rows_valid) {
      if (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(src + (int64_t)r * ld + c);
        v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
      } else {
        v[u][0] = src[(int64_t)r * ld + c];
      }
    }
  }
}
// visit the (row, col) of every staged value: f(u, e, r, c)
template <int VEC, int NV, int NTHR, typename F>
__device__ __forceinline__ void tile_visit(int ncols, int tile_rows, int tid, F&& f) {
  const int cpr = (ncols + VEC - 1) / VEC;
  const int total = tile_rows * cpr;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    const int idx = tid + u * NTHR;
    if (idx < total) {
      const int r = idx / cpr, c = (idx - r * cpr) * VEC;
#pragma unroll
      for (int e = 0; e < VEC; ++e) f(u, e, r, c + e);
    }
  }
}
__device__ __forceinline__ bool vec4_ok(const void* p, int64_t ld, int ncols) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0 && (ncols & 3) == 0;
}

// fp64 sum over the 8 row groups of a warp (lanes with equal tig), result valid in lanes g == 0
__device__ __forceinline__ double group_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// =============================================================================================
// forward
// =============================================================================================
template <int NT>
__global__ void __launch_bounds__(MMA_THREADS) dense_fwd_mma_kernel(const DenseFwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) & 3, wc = tid >> 7;
  const int g = lane >> 2, tig = lane & 3;
  const int nin = p.nin, nout = p.nout;
  const int Kp = (nin + 7) & ~7;
  const int ksteps = Kp / 8;
  const int bp = b_pitch(8 * NT), ap = a_pitch(nin);
  constexpr int NTW = NT / NWC;               // n-tiles per warp (column quarters)
  const int nt_used = max(0, min(NTW, (nout + 7) / 8 - wc * NTW));
  float* Wt = smem;                       // [Kp][bp]   Wt[k][n] = W[n][k]
  float* Xs = Wt + Kp * bp;               // [64][ap]
  float* bias = Xs + ROWS_PER_CTA * ap;   // [8*NT]
  float* mean = bias + 8 * NT;            // [Kp]
  float* rstd = mean + Kp;                // [Kp]
  double* red = reinterpret_cast<double*>(smem + (((rstd + Kp) - smem + 1) & ~(ptrdiff_t)1));  // [4][2][8*NT]

  const float* W = p.params + (int64_t)arm * p.p_arm_stride + p.offW;
  const float* bsrc = p.params + (int64_t)arm * p.p_arm_stride + p.offB;
  for (int idx = tid; idx < Kp * bp + ROWS_PER_CTA * ap; idx += MMA_THREADS) Wt[idx] = 0.f;   // Wt and Xs (padding stays 0)
  __syncthreads();
  // W[n][k] -> Wt[k][n]; consecutive threads take consecutive n: conflict-free stores, the weights are L2-resident
#pragma unroll 8
  for (int idx = tid; idx < nout * nin; idx += MMA_THREADS) {
    const int k = idx / nout, n = idx - k * nout;
    Wt[k * bp + n] = W[n * nin + k];
  }
  for (int j = tid; j < 8 * NT; j += MMA_THREADS) bias[j] = j < nout ? bsrc[j] : 0.f;
  if (p.bn_mode == 1) {
    const double* sums = p.bn_sums_in + (int64_t)arm * 256;
    for (int i = tid; i < nin; i += MMA_THREADS) {
      const double m = sums[i] / (double)p.B;
      double var = sums[128 + i] / (double)p.B - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)p.eps));
      mean[i] = mf;
      rstd[i] = rf;
      if (blockIdx.x == 0) {
        p.bn_mean[arm * 128 + i] = mf;
        p.bn_rstd[arm * 128 + i] = rf;
      }
    }
  } else if (p.bn_mode == 2) {
    for (int i = tid; i < nin; i += MMA_THREADS) {
      mean[i] = p.bn_mean[arm * 128 + i];
      rstd[i] = p.bn_rstd[arm * 128 + i];
    }
  }
  const float* in = p.in + (int64_t)arm * p.in_arm_stride;
  float* out = p.out + (int64_t)arm * p.out_arm_stride;
  const int ntiles = (p.B + ROWS_PER_CTA - 1) / ROWS_PER_CTA;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * ROWS_PER_CTA;
    __syncthreads();   // weights/stats ready; previous tile's Xs no longer read
    {
      const float* src = in + (int64_t)row0 * nin;
      const int rows_valid = min(ROWS_PER_CTA, p.B - row0);
      auto put = [&](float val, int r, int c) {
        if (p.bn_mode) val = (val - mean[c]) * rstd[c];
        Xs[r * ap + c] = (r < rows_valid) ? val : 0.f;
      };
      if (vec4_ok(src, nin, nin)) {
        float v[4][4];
        tile_load<4, 4, MMA_THREADS>(v, src, nin, rows_valid, nin, ROWS_PER_CTA, tid);
        tile_visit<4, 4, MMA_THREADS>(nin, ROWS_PER_CTA, tid, [&](int u, int e, int r, int c) { put(v[u][e], r, c); });
      } else {
        float v[16][1];
        tile_load<1, 16, MMA_THREADS>(v, src, nin, rows_valid, nin, ROWS_PER_CTA, tid);
        tile_visit<1, 16, MMA_THREADS>(nin, ROWS_PER_CTA, tid, [&](int u, int e, int r, int c) { put(v[u][e], r, c); });
      }
    }
    __syncthreads();
    float acc[NTW][4];
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    warp_gemm<NTW, false>(Xs + warp * 16 * ap, ap, Wt + wc * NTW * 8, bp, ksteps, nt_used, acc, lane);
    // ---- epilogue: bias, ReLU, store, fp64 column sums
    const int ra = row0 + warp * 16 + g, rb = ra + 8;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt < nt_used) {
        const int c = (wc * NTW + nt) * 8 + 2 * tig;
        float v00 = acc[nt][0] + bias[c], v01 = acc[nt][1] + bias[c + 1];
        float v10 = acc[nt][2] + bias[c], v11 = acc[nt][3] + bias[c + 1];
        if (p.relu) { v00 = fmaxf(v00, 0.f); v01 = fmaxf(v01, 0.f); v10 = fmaxf(v10, 0.f); v11 = fmaxf(v11, 0.f); }
        const bool va = ra < p.B, vb = rb < p.B;
        if (va) { if (c < nout) out[(int64_t)ra * nout + c] = v00; if (c + 1 < nout) out[(int64_t)ra * nout + c + 1] = v01; }
        if (vb) { if (c < nout) out[(int64_t)rb * nout + c] = v10; if (c + 1 < nout) out[(int64_t)rb * nout + c + 1] = v11; }
        if (p.stats_out) {
          const double a0 = va ? (double)v00 : 0.0, a1 = va ? (double)v01 : 0.0;
          const double b0 = vb ? (double)v10 : 0.0, b1 = vb ? (double)v11 : 0.0;
          const double s0 = group_sum(a0 + b0), s1 = group_sum(a1 + b1);
          const double q0 = group_sum(a0 * a0 + b0 * b0), q1 = group_sum(a1 * a1 + b1 * b1);
          if (g == 0) {
            red[(warp * 2 + 0) * 8 * NT + c] = s0;
            red[(warp * 2 + 0) * 8 * NT + c + 1] = s1;
            red[(warp * 2 + 1) * 8 * NT + c] = q0;
            red[(warp * 2 + 1) * 8 * NT + c + 1] = q1;
          }
        }
      }
    }
    if (p.stats_out) {
      __syncthreads();
      for (int j = tid; j < 2 * 8 * NT; j += MMA_THREADS) {
        const int which = j / (8 * NT), c = j - which * 8 * NT;
        if (c < nout) {
          double s = 0.0;
          for (int w = 0; w < 4; ++w) s += red[(w * 2 + which) * 8 * NT + c];
          atomicAdd(p.stats_out + (int64_t)arm * 256 + which * 128 + c, s);
        }
      }
    }
  }
}

// =============================================================================================
// backward (data gradient)
// =============================================================================================
template <int NT, bool SPLIT>
__global__ void __launch_bounds__(MMA_THREADS) dense_bwd_mma_kernel(const DenseBwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) & 3, wc = tid >> 7;
  const int g = lane >> 2, tig = lane & 3;
  const int nin = p.nin, nout = p.nout;
  const int Kp = (nout + 7) & ~7;          // reduction over this layer's outputs
  const int ksteps = Kp / 8;
  const int bp = b_pitch(8 * NT), ap = a_pitch(nout);
  constexpr int NTW = NT / NWC;
  const int nt_used = max(0, min(NTW, (nin + 7) / 8 - wc * NTW));
  float* Ws = smem;                        // [Kp][bp]   Ws[j][i] = W[j][i]
  float* Ds = Ws + Kp * bp;                // [64][ap]   delta tile
  float* c1 = Ds + ROWS_PER_CTA * ap;      // [Kp]
  float* c2 = c1 + Kp;
  float* mo = c2 + Kp;
  float* ro = mo + Kp;
  float* mi = ro + Kp;                     // [8*NT]
  float* ri = mi + 8 * NT;
  double* red = reinterpret_cast<double*>(smem + (((ri + 8 * NT) - smem + 1) & ~(ptrdiff_t)1));   // [4][2][8*NT]

  if (p.g_in) {
    const float* W = p.params + (int64_t)arm * p.p_arm_stride + p.offW;
    for (int idx = tid; idx < Kp * bp; idx += MMA_THREADS) Ws[idx] = 0.f;
    __syncthreads();
#pragma unroll 8
    for (int idx = tid; idx < nout * nin; idx += MMA_THREADS) {
      const int j = idx / nin, i = idx - j * nin;
      Ws[j * bp + i] = W[idx];
    }
  }
  for (int idx = tid; idx < ROWS_PER_CTA * ap; idx += MMA_THREADS) Ds[idx] = 0.f;
  if (p.bn_out) {
    const double* sums = p.bnb_sums + (int64_t)arm * 256;
    for (int j = tid; j < nout; j += MMA_THREADS) {
      c1[j] = (float)(sums[j] / (double)p.B);
      c2[j] = (float)(sums[128 + j] / (double)p.B);
      mo[j] = p.mean_out[arm * 128 + j];
      ro[j] = p.rstd_out[arm * 128 + j];
    }
  }
  if (p.bn_in) {
    for (int i = tid; i < 8 * NT; i += MMA_THREADS) {
      mi[i] = i < nin ? p.mean_in[arm * 128 + i] : 0.f;
      ri[i] = i < nin ? p.rstd_in[arm * 128 + i] : 0.f;
    }
  }
  const int64_t abo = (int64_t)arm * p.B;
  const float* g_out = p.g_out + abo * nout;
  const float* act_out = p.act_out + abo * nout;
  float* delta = p.delta + abo * nout;
  float* g_in = p.g_in ? p.g_in + abo * nin : nullptr;
  const float* act_in = p.bn_in ? p.act_in + abo * nin : nullptr;
  const int ntiles = (p.B + ROWS_PER_CTA - 1) / ROWS_PER_CTA;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * ROWS_PER_CTA;
    __syncthreads();
    {
      const float* gsrc = g_out + (int64_t)row0 * nout;
      const float* asrc = act_out + (int64_t)row0 * nout;
      float* dsts = delta + (int64_t)row0 * nout;
      const int rows_valid = min(ROWS_PER_CTA, p.B - row0);
      auto put = [&](float gg, float a, int r, int j) {
        float d = 0.f;
        if (r < rows_valid) {
          if (p.bn_out) {
            const float n = (a - mo[j]) * ro[j];
            gg = ro[j] * (gg - c1[j] - n * c2[j]);
          }
          d = a > 0.f ? gg : 0.f;
          dsts[(int64_t)r * nout + j] = d;
        }
        Ds[r * ap + j] = d;
      };
      if (vec4_ok(gsrc, nout, nout) && vec4_ok(asrc, nout, nout)) {
        float vg[4][4], va[4][4];
        tile_load<4, 4, MMA_THREADS>(vg, gsrc, nout, rows_valid, nout, ROWS_PER_CTA, tid);
        tile_load<4, 4, MMA_THREADS>(va, asrc, nout, rows_valid, nout, ROWS_PER_CTA, tid);
        tile_visit<4, 4, MMA_THREADS>(nout, ROWS_PER_CTA, tid, [&](int u, int e, int r, int j) { put(vg[u][e], va[u][e], r, j); });
      } else {
        float vg[16][1], va[16][1];
        tile_load<1, 16, MMA_THREADS>(vg, gsrc, nout, rows_valid, nout, ROWS_PER_CTA, tid);
        tile_load<1, 16, MMA_THREADS>(va, asrc, nout, rows_valid, nout, ROWS_PER_CTA, tid);
        tile_visit<1, 16, MMA_THREADS>(nout, ROWS_PER_CTA, tid, [&](int u, int e, int r, int j) { put(vg[u][e], va[u][e], r, j); });
      }
    }
    if (!g_in) continue;
    __syncthreads();
    float acc[NTW][4];
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    warp_gemm<NTW, false, SPLIT>(Ds + warp * 16 * ap, ap, Ws + wc * NTW * 8, bp, ksteps, nt_used, acc, lane);
    const int ra = row0 + warp * 16 + g, rb = ra + 8;
    const bool va = ra < p.B, vb = rb < p.B;
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      if (nt < nt_used) {
        const int c = (wc * NTW + nt) * 8 + 2 * tig;
        const bool c0ok = c < nin, c1ok = c + 1 < nin;
        if (va) { if (c0ok) g_in[(int64_t)ra * nin + c] = acc[nt][0]; if (c1ok) g_in[(int64_t)ra * nin + c + 1] = acc[nt][1]; }
        if (vb) { if (c0ok) g_in[(int64_t)rb * nin + c] = acc[nt][2]; if (c1ok) g_in[(int64_t)rb * nin + c + 1] = acc[nt][3]; }
        if (p.bn_in) {
          double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
          if (va) {
            if (c0ok) { const double n = (double)((act_in[(int64_t)ra * nin + c] - mi[c]) * ri[c]); s0 += (double)acc[nt][0]; q0 += (double)acc[nt][0] * n; }
            if (c1ok) { const double n = (double)((act_in[(int64_t)ra * nin + c + 1] - mi[c + 1]) * ri[c + 1]); s1 += (double)acc[nt][1]; q1 += (double)acc[nt][1] * n; }
          }
          if (vb) {
            if (c0ok) { const double n = (double)((act_in[(int64_t)rb * nin + c] - mi[c]) * ri[c]); s0 += (double)acc[nt][2]; q0 += (double)acc[nt][2] * n; }
            if (c1ok) { const double n = (double)((act_in[(int64_t)rb * nin + c + 1] - mi[c + 1]) * ri[c + 1]); s1 += (double)acc[nt][3]; q1 += (double)acc[nt][3] * n; }
          }
          s0 = group_sum(s0); s1 = group_sum(s1); q0 = group_sum(q0); q1 = group_sum(q1);
          if (g == 0) {
            red[(warp * 2 + 0) * 8 * NT + c] = s0;
            red[(warp * 2 + 0) * 8 * NT + c + 1] = s1;
            red[(warp * 2 + 1) * 8 * NT + c] = q0;
            red[(warp * 2 + 1) * 8 * NT + c + 1] = q1;
          }
        }
      }
    }
    if (p.bn_in) {
      __syncthreads();
      for (int j = tid; j < 2 * 8 * NT; j += MMA_THREADS) {
        const int which = j / (8 * NT), c = j - which * 8 * NT;
        if (c < nin) {
          double s = 0.0;
          for (int w = 0; w < 4; ++w) s += red[(w * 2 + which) * 8 * NT + c];
          atomicAdd(p.bnb_sums_next + (int64_t)arm * 256 + which * 128 + c, s);
        }
      }
    }
  }
}

// =============================================================================================
// weight gradients:  dW[j][i] = sum_b delta[b][j] in[b][i],  db[j] = sum_b delta[b][j] (a ones column)
// =============================================================================================
constexpr int WG_CHUNK = 32;

template <bool SPLIT>
__global__ void __launch_bounds__(512) wgrad_mma_kernel(const WgArgs p) {
  __shared__ __align__(16) float Ds[WG_CHUNK * 136];
  __shared__ __align__(16) float Is[WG_CHUNK * 136];
  __shared__ float bm[128], br[128];
  const WgProblem& pr = p.prob[blockIdx.y];
  const int arm = blockIdx.z, split = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) & 7, wn = tid >> 8;   // m-tile, n half
  const int g = lane >> 2, tig = lane & 3;
  const int nout = pr.nout, nin = pr.nin;
  const int nt_all = (nin + 1 + 7) / 8;       // + the ones column that yields the bias gradient
  const int nt_used = max(0, min(8, nt_all - wn * 8));
  const float* delta = p.work + pr.delta_off + (int64_t)arm * pr.delta_arm_stride;
  const float* in = nin > 0 ? p.work + pr.in_off + (int64_t)arm * pr.in_arm_stride : nullptr;
  if (pr.bn_layer >= 0) {
    for (int i = tid; i < nin; i += 512) {
      bm[i] = p.bn_mean[(pr.bn_layer * p.A + arm) * 128 + i];
      br[i] = p.bn_rstd[(pr.bn_layer * p.A + arm) * 128 + i];
    }
  }
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
  const int r0 = split * p.rows_per_split;
  const int r1 = min(p.B, r0 + p.rows_per_split);
  const bool active = warp * 16 < nout && nt_used > 0;
  for (int idx = tid; idx < WG_CHUNK * 136; idx += 512) { Ds[idx] = 0.f; Is[idx] = 0.f; }
  const bool dv4 = vec4_ok(delta, nout, nout);
  const bool iv4 = nin > 0 && vec4_ok(in, pr.in_ld, nin);
  // chunk staged through registers: the loads of chunk c+1 are in flight while chunk c is multiplied
  float vd[8][1], vi[8][1];              // scalar path: 32x128 / 512 threads
  float vd4[2][4], vi4[2][4];            // float4 path
  // float4 path: the (row, column) of this thread's two float4 per operand are the same for every chunk
  int drow[2] = {0, 0}, dcol[2] = {0, 0}, irow[2] = {0, 0}, icol[2] = {0, 0};
  bool dok[2] = {false, false}, iok[2] = {false, false};
  if (dv4) {
    const int cpr = nout >> 2;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + u * 512;
      dok[u] = idx < WG_CHUNK * cpr;
      drow[u] = idx / cpr;
      dcol[u] = (idx - drow[u] * cpr) * 4;
    }
  }
  if (iv4) {
    const int cpr = nin >> 2;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + u * 512;
      iok[u] = idx < WG_CHUNK * cpr;
      irow[u] = idx / cpr;
      icol[u] = (idx - irow[u] * cpr) * 4;
    }
  }
  auto load_chunk = [&](int rb) {
    const int nr = min(WG_CHUNK, r1 - rb);
    if (dv4) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dok[u] && drow[u] < nr) t4 = *reinterpret_cast<const float4*>(delta + (int64_t)(rb + drow[u]) * nout + dcol[u]);
        vd4[u][0] = t4.x; vd4[u][1] = t4.y; vd4[u][2] = t4.z; vd4[u][3] = t4.w;
      }
    } else {
      tile_load<1, 8, 512>(vd, delta + (int64_t)rb * nout, nout, nr, nout, WG_CHUNK, tid);
    }
    if (nin > 0) {
      if (iv4) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (iok[u] && irow[u] < nr) t4 = *reinterpret_cast<const float4*>(in + (int64_t)(rb + irow[u]) * pr.in_ld + icol[u]);
          vi4[u][0] = t4.x; vi4[u][1] = t4.y; vi4[u][2] = t4.z; vi4[u][3] = t4.w;
        }
      } else {
        tile_load<1, 8, 512>(vi, in + (int64_t)rb * pr.in_ld, pr.in_ld, nr, nin, WG_CHUNK, tid);
      }
    }
  };
  auto store_chunk = [&](int rb) {
    const int nr = min(WG_CHUNK, r1 - rb);
    if (dv4) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (dok[u])
          *reinterpret_cast<float4*>(Ds + drow[u] * 136 + dcol[u]) = make_float4(vd4[u][0], vd4[u][1], vd4[u][2], vd4[u][3]);
    } else {
      tile_visit<1, 8, 512>(nout, WG_CHUNK, tid, [&](int u, int e, int r, int j) { Ds[r * 136 + j] = vd[u][e]; });
    }
    auto puti = [&](float v, int r, int i) {
      if (pr.bn_layer >= 0) v = (v - bm[i]) * br[i];
      Is[r * 136 + i] = r < nr ? v : 0.f;
    };
    if (nin > 0) {
      if (iv4) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (iok[u]) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float v = vi4[u][e];
              if (pr.bn_layer >= 0) v = (v - bm[icol[u] + e]) * br[icol[u] + e];
              o[e] = irow[u] < nr ? v : 0.f;
            }
            *reinterpret_cast<float4*>(Is + irow[u] * 136 + icol[u]) = make_float4(o[0], o[1], o[2], o[3]);
          }
      } else {
        tile_visit<1, 8, 512>(nin, WG_CHUNK, tid, [&](int u, int e, int r, int i) { puti(vi[u][e], r, i); });
      }
    }
    if (tid < WG_CHUNK) Is[tid * 136 + nin] = tid < nr ? 1.f : 0.f;   // ones column -> bias gradient
  };
  if (r0 < r1) load_chunk(r0);
  for (int rb = r0; rb < r1; rb += WG_CHUNK) {
    __syncthreads();                       // previous chunk fully consumed
    store_chunk(rb);
    __syncthreads();
    if (rb + WG_CHUNK < r1) load_chunk(rb + WG_CHUNK);
    if (active) {
      // constant tile counts for the common shapes (nin = 100: 8 + 5 tiles): the per-tile guards fold away
      if (nt_used == 8) warp_gemm<8, true, SPLIT>(Ds + warp * 16, 136, Is + wn * 64, 136, WG_CHUNK / 8, 8, acc, lane);
      else if (nt_used == 5) warp_gemm<8, true, SPLIT>(Ds + warp * 16, 136, Is + wn * 64, 136, WG_CHUNK / 8, 5, acc, lane);
      else warp_gemm<8, true, SPLIT>(Ds + warp * 16, 136, Is + wn * 64, 136, WG_CHUNK / 8, nt_used, acc, lane);
    }
  }
  if (!active) return;
  float* part = p.part + (int64_t)split * p.part_split_stride + (int64_t)arm * p.part_arm_stride;
  const int ja = warp * 16 + g, jb = ja + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < nt_used) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = (e & 2) ? jb : ja;
        const int i = (wn * 8 + nt) * 8 + 2 * tig + (e & 1);
        if (j < nout) {
          if (i < nin) part[pr.poffW - p.base_off + (int64_t)j * nin + i] = acc[nt][e];
          else if (i == nin) part[pr.poffB - p.base_off + j] = acc[nt][e];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Second generation (default): the first kernel above staged chunks through registers with 126 registers per
// thread (one CTA per SM), gave every problem the same 20 row splits (520 CTAs = 3.5 waves) and spent most of its
// issue slots on index arithmetic and barrier waits (ncu: 31 M instructions for 0.75 M MMAs).  Here
//   * chunks of 32 rows arrive through a 3-stage cp.async ring (16-byte pieces when rows are 16-byte aligned,
//     8/4-byte pieces otherwise; rows beyond the split are zero-filled by cp.async itself), one barrier per chunk;
//   * the BatchNorm normalisation of the input is applied in place by the thread that copied the piece;
//   * split counts are per problem (wide problems get twice the splits of narrow ones) so that the whole grid is
//     ONE wave of two co-resident CTAs per SM; chunks are dealt to the splits evenly.
// Same partial layout and the same fixed-order reduce as before.
// ---------------------------------------------------------------------------------------------
constexpr int WG2_STAGES = 3;
constexpr int WG2_PITCH = 136;
constexpr int WG2_STAGE_FLOATS = 2 * WG_CHUNK * WG2_PITCH;          // Ds | Is
constexpr int WG2_SMEM_FLOATS = 256 + WG2_STAGES * WG2_STAGE_FLOATS;  // bm | br | stages

template <int PF>
__device__ __forceinline__ void wg2_cp(float* dst, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? PF * 4 : 0;       // src-size 0: the destination is zero-filled
  if (PF == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
  else if (PF == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ int wg2_piece_floats(const void* p, int64_t ld, int ncols) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if ((a & 15) == 0 && (ld & 3) == 0 && (ncols & 3) == 0) return 4;
  if ((a & 7) == 0 && (ld & 1) == 0 && (ncols & 1) == 0) return 2;
  return 1;
}
// One row of a chunk per 16 threads: thread (r, l) copies the pieces l, l + 16, ... of row r (PF floats each).  No
// per-piece index arithmetic: the loop is unrolled, the column offsets are immediates.
template <int PF>
__device__ __forceinline__ void wg2_copy_row(float* srow, const float* grow, int l, int ncols, bool row_ok) {
#pragma unroll
  for (int j = 0; j < 128 / (16 * PF); ++j) {
    const int c = (l + 16 * j) * PF;
    if (c < ncols) wg2_cp<PF>(srow + c, grow + c, row_ok);
  }
}
// (v - mean) * rstd on the pieces this thread copied itself
template <int PF>
__device__ __forceinline__ void wg2_norm_row(float* srow, const float* bm, const float* br, int l, int ncols) {
#pragma unroll
  for (int j = 0; j < 128 / (16 * PF); ++j) {
    const int c = (l + 16 * j) * PF;
    if (c < ncols) {
      if (PF == 4) {
        float4 v = *reinterpret_cast<float4*>(srow + c);
        const float4 m = *reinterpret_cast<const float4*>(bm + c), rs = *reinterpret_cast<const float4*>(br + c);
        v.x = (v.x - m.x) * rs.x; v.y = (v.y - m.y) * rs.y; v.z = (v.z - m.z) * rs.z; v.w = (v.w - m.w) * rs.w;
        *reinterpret_cast<float4*>(srow + c) = v;
      } else {
#pragma unroll
        for (int e = 0; e < PF; ++e) srow[c + e] = (srow[c + e] - bm[c + e]) * br[c + e];
      }
    }
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(512, SPLIT ? 1 : 2) wgrad2_kernel(const WgArgs p) {
  extern __shared__ __align__(16) float wsm[];
  float* bm = wsm;
  float* br = wsm + 128;
  float* stages = wsm + 256;
  pdl_trigger();
  pdl_wait();
  int pi = 0;
  while (pi + 1 < p.nprob && (int)blockIdx.x >= p.prob[pi + 1].cta_begin) ++pi;
  const WgProblem& pr = p.prob[pi];
  const int split = (int)blockIdx.x - pr.cta_begin, arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nout = pr.nout, nin = pr.nin, in_ld = pr.in_ld;
  const int nt_all = (nin + 1 + 7) / 8;       // + the ones column that yields the bias gradient
  // tile ownership: 8 m-tiles x 2 halves of the n-tiles, or -- one m-tile only (nout <= 16) -- one n-tile per warp, so
  // that the k-steps of a chunk are not serialised on two warps while fourteen wait at the barrier
  const bool thin_m = nout <= 16;
  const int mt = thin_m ? 0 : (tid >> 5) & 7;
  const int nt0 = thin_m ? (tid >> 5) : (tid >> 8) * 8;
  const int nt_used = thin_m ? ((tid >> 5) < nt_all ? 1 : 0) : max(0, min(8, nt_all - nt0));
  const float* delta = p.work + pr.delta_off + (int64_t)arm * pr.delta_arm_stride;
  const float* in = nin > 0 ? p.work + pr.in_off + (int64_t)arm * pr.in_arm_stride : nullptr;
  const bool bn = pr.bn_layer >= 0;
  // chunks of this split: dealt evenly (sizes differ by at most one chunk)
  const int T = (p.B + WG_CHUNK - 1) / WG_CHUNK;
  const int c0 = (int)((int64_t)split * T / pr.nsplit), c1 = (int)((int64_t)(split + 1) * T / pr.nsplit);
  const int nchunks = c1 - c0;

  {
    // cp.async fills columns [0, nout) / [0, nin) of all 32 rows of a stage (rows beyond the split as zeros); the MMA
    // tiles also read the pad columns up to the next multiple of 16 / 8: zero exactly those (not the whole 105 KB)
    const int r = tid >> 4, l = tid & 15;
    const int mpad = (nout + 15) & ~15, npad = nt_all * 8;
    for (int st = 0; st < WG2_STAGES; ++st) {
      float* drow = stages + st * WG2_STAGE_FLOATS + r * WG2_PITCH;
      float* irow = drow + WG_CHUNK * WG2_PITCH;
      if (nout + l < mpad) drow[nout + l] = 0.f;
      if (nin + 1 + l < npad) irow[nin + 1 + l] = 0.f;
    }
  }
  if (tid < 128) {
    const bool have = bn && tid < nin;
    bm[tid] = have ? p.bn_mean[(pr.bn_layer * p.A + arm) * 128 + tid] : 0.f;
    br[tid] = have ? p.bn_rstd[(pr.bn_layer * p.A + arm) * 128 + tid] : 1.f;
  }
  __syncthreads();
  if (tid < WG2_STAGES * WG_CHUNK)        // ones column -> bias gradient (rows beyond the split have delta = 0)
    stages[(tid / WG_CHUNK) * WG2_STAGE_FLOATS + WG_CHUNK * WG2_PITCH + (tid % WG_CHUNK) * WG2_PITCH + nin] = 1.f;

  const int pfd = wg2_piece_floats(delta, nout, nout);
  const int pfi = nin > 0 ? wg2_piece_floats(in, in_ld, nin) : 4;
  const int lr = tid >> 4, ll = tid & 15;                 // row of the chunk, piece lane
  const float* gd = delta + (int64_t)lr * nout;           // this thread's row in chunk 0 of the matrix
  const float* gi = nin > 0 ? in + (int64_t)lr * in_ld : nullptr;
  float* const srow_d = stages + lr * WG2_PITCH;
  float* const srow_i = stages + WG_CHUNK * WG2_PITCH + lr * WG2_PITCH;
  auto issue = [&](int c) {
    if (c < nchunks) {
      const int so = (c % WG2_STAGES) * WG2_STAGE_FLOATS;
      const int rb = (c0 + c) * WG_CHUNK;
      const bool ok = rb + lr < p.B;
      const float* d = ok ? gd + (int64_t)rb * nout : delta;   // zero-filled rows still get a valid address (row 0)
      if (pfd == 4) wg2_copy_row<4>(srow_d + so, d, ll, nout, ok);
      else if (pfd == 2) wg2_copy_row<2>(srow_d + so, d, ll, nout, ok);
      else wg2_copy_row<1>(srow_d + so, d, ll, nout, ok);
      if (nin > 0) {
        const float* i = ok ? gi + (int64_t)rb * in_ld : in;
        if (pfi == 4) wg2_copy_row<4>(srow_i + so, i, ll, nin, ok);
        else if (pfi == 2) wg2_copy_row<2>(srow_i + so, i, ll, nin, ok);
        else wg2_copy_row<1>(srow_i + so, i, ll, nin, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
  const bool active = mt * 16 < nout && nt_used > 0;
  __syncthreads();                          // the zero fill is ordered before the first cp.async lands
#pragma unroll 1
  for (int c = 0; c < WG2_STAGES - 1; ++c) issue(c);
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    asm volatile("cp.async.wait_group %0;" ::"n"(WG2_STAGES - 2) : "memory");   // this thread's pieces of chunk c landed
    const int so = (c % WG2_STAGES) * WG2_STAGE_FLOATS;
    if (bn) {
      if (pfi == 4) wg2_norm_row<4>(srow_i + so, bm, br, ll, nin);
      else if (pfi == 2) wg2_norm_row<2>(srow_i + so, bm, br, ll, nin);
      else wg2_norm_row<1>(srow_i + so, bm, br, ll, nin);
    }
    __syncthreads();                        // chunk c visible to all warps; chunk c-1 fully consumed
    issue(c + WG2_STAGES - 1);              // refills the stage chunk c-1 used
    if (active) {
      const float* Ds = stages + so;
      const float* Is = Ds + WG_CHUNK * WG2_PITCH;
      if (nt_used == 8) warp_gemm<8, true, SPLIT>(Ds + mt * 16, WG2_PITCH, Is + nt0 * 8, WG2_PITCH, WG_CHUNK / 8, 8, acc, lane);
      else if (nt_used == 1) warp_gemm<8, true, SPLIT>(Ds + mt * 16, WG2_PITCH, Is + nt0 * 8, WG2_PITCH, WG_CHUNK / 8, 1, acc, lane);
      else warp_gemm<8, true, SPLIT>(Ds + mt * 16, WG2_PITCH, Is + nt0 * 8, WG2_PITCH, WG_CHUNK / 8, nt_used, acc, lane);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (!active) return;
  float* part = p.part + (int64_t)split * p.part_split_stride + (int64_t)arm * p.part_arm_stride;
  const int g = lane >> 2, tig = lane & 3;
  const int ja = mt * 16 + g, jb = ja + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (nt < nt_used) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = (e & 2) ? jb : ja;
        const int i = (nt0 + nt) * 8 + 2 * tig + (e & 1);
        if (j < nout) {
          if (i < nin) part[pr.poffW - p.base_off + (int64_t)j * nin + i] = acc[nt][e];
          else if (i == nin) part[pr.poffB - p.base_off + j] = acc[nt][e];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce2_kernel(const WgArgs p) {
  const WgProblem& pr = p.prob[blockIdx.y >> 1];
  const int which = blockIdx.y & 1, arm = blockIdx.z;
  const int64_t n = which ? pr.nout : (int64_t)pr.nout * pr.nin;
  const int64_t poff = which ? pr.poffB : pr.poffW;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (e >= n) return;
  const float* part = p.part + (int64_t)arm * p.part_arm_stride + (poff - p.base_off) + e;
  float s = 0.f;
  for (int sp = 0; sp < pr.nsplit; ++sp) s += part[(int64_t)sp * p.part_split_stride];
  p.grads[(int64_t)arm * p.g_arm_stride + poff + e] = s;
}

template <int NT>
size_t fwd_smem(int nin) {
  const int Kp = (nin + 7) & ~7;
  size_t fl = (size_t)Kp * b_pitch(8 * NT) + (size_t)ROWS_PER_CTA * a_pitch(nin) + 8 * NT + 2 * Kp + 2;
  return fl * 4 + (size_t)4 * 2 * 8 * NT * 8;
}
template <int NT>
size_t bwd_smem(int nout) {
  const int Kp = (nout + 7) & ~7;
  size_t fl = (size_t)Kp * b_pitch(8 * NT) + (size_t)ROWS_PER_CTA * a_pitch(nout) + 4 * Kp + 2 * 8 * NT + 2;
  return fl * 4 + (size_t)4 * 2 * 8 * NT * 8;
}

int mma_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

int launch_dense_fwd_mma(const DenseFwdArgs& a, int A, cudaStream_t s) {
  // one CTA per SM (registers): keep the grid inside ONE wave and let CTAs loop over their tiles -- 79 tiles per arm on
  // 74 SMs per arm cost one extra tile pass for a few CTAs instead of a second wave of whole-CTA latency
  const int ntiles = (a.B + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  int per_arm = mma_sm_count() / (A > 0 ? A : 1);
  if (per_arm < 1) per_arm = 1;
  dim3 grid(ntiles > per_arm ? per_arm : ntiles, A);
#define LAUNCH(NT)                                                                                                 \
  do {                                                                                                             \
    static bool attr[64] = {};                                                                                      \
    if (first_on_device(attr)) {                                                                                                   \
      MVAE_CUDA(cudaFuncSetAttribute(dense_fwd_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
    }                                                                                                              \
    dense_fwd_mma_kernel<NT><<<grid, MMA_THREADS, fwd_smem<NT>(a.nin), s>>>(a);                                    \
  } while (0)
  if (a.nout <= 32) LAUNCH(4);
  else if (a.nout <= 64) LAUNCH(8);
  else if (a.nout <= 128) LAUNCH(16);
  else { set_error("dense_fwd_mma: nout=%d too wide", a.nout); return -1; }
#undef LAUNCH
  MVAE_LAUNCH_CHECK();
  return 0;
}

int launch_dense_bwd_mma(const DenseBwdArgs& a, int A, int split3, cudaStream_t s) {
  // one CTA per SM (registers): keep the grid inside ONE wave and let CTAs loop over their tiles -- 79 tiles per arm on
  // 74 SMs per arm cost one extra tile pass for a few CTAs instead of a second wave of whole-CTA latency
  const int ntiles = (a.B + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  int per_arm = mma_sm_count() / (A > 0 ? A : 1);
  if (per_arm < 1) per_arm = 1;
  dim3 grid(ntiles > per_arm ? per_arm : ntiles, A);
  const int nin = a.g_in ? a.nin : 1;
#define LAUNCH(NT, SP)                                                                                             \
  do {                                                                                                             \
    MVAE_CUDA(cudaFuncSetAttribute(dense_bwd_mma_kernel<NT, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
    dense_bwd_mma_kernel<NT, SP><<<grid, MMA_THREADS, bwd_smem<NT>(a.nout), s>>>(a);                               \
  } while (0)
  if (split3) {
    if (nin <= 32) LAUNCH(4, true);
    else if (nin <= 64) LAUNCH(8, true);
    else if (nin <= 128) LAUNCH(16, true);
    else { set_error("dense_bwd_mma: nin=%d too wide", a.nin); return -1; }
  } else {
    if (nin <= 32) LAUNCH(4, false);
    else if (nin <= 64) LAUNCH(8, false);
    else if (nin <= 128) LAUNCH(16, false);
    else { set_error("dense_bwd_mma: nin=%d too wide", a.nin); return -1; }
  }
#undef LAUNCH
  MVAE_LAUNCH_CHECK();
  return 0;
}

int launch_wgrad_mma(const WgArgs& a0, int split3, cudaStream_t s, const WgFork* fork) {
  WgArgs a = a0;
  const bool thin_side = fork != nullptr, red_side = fork != nullptr;
  if (thin_side) MVAE_CUDA(cudaEventRecord(fork->fork_ev, s));
  bool fits = true;
  for (int i = 0; i < a.nprob; ++i) fits = fits && a.prob[i].nout <= 128 && a.prob[i].nin <= 127;
  if (fits) {
    // the wide problems (many MMA tiles per chunk of cells) run on tcgen05 when TMA can read their operands
    int tc_idx[8], ntc = 0;
    bool on_tc[13];
    for (int i = 0; i < a.nprob; ++i) {
      const int tiles = ((a.prob[i].nout + 15) / 16) * ((a.prob[i].nin + 1 + 7) / 8);
      on_tc[i] = tiles > 40 && ntc < 8 && tc_narrow_wgrad_ok(a, a.prob[i]);
      if (on_tc[i]) tc_idx[ntc++] = i;
    }
    if (ntc > 0) {
      const int rc = tc_narrow_wgrad(a, tc_idx, ntc, split3, s, thin_side);
      if (rc) return rc;
    }
    // the rest: split counts for one wave of two CTAs per SM; wide problems get twice the splits of narrow ones
    int wsum = 0, wt[13];
    for (int i = 0; i < a.nprob; ++i) {
      const int tiles = ((a.prob[i].nout + 15) / 16) * ((a.prob[i].nin + 1 + 7) / 8);
      wt[i] = on_tc[i] ? 0 : (tiles > 40 ? 2 : 1);
      wsum += wt[i];
    }
    const int T = (a.B + WG_CHUNK - 1) / WG_CHUNK;
    int per = wsum > 0 ? (2 * mma_sm_count()) / (a.A * wsum) : 1;
    if (per < 1) per = 1;
    int ctas = 0;
    for (int i = 0; i < a.nprob; ++i) {
      a.prob[i].cta_begin = ctas;
      if (on_tc[i]) continue;               // keeps the split count tc_narrow_wgrad chose; no CTA here
      int n = per * wt[i];
      if (n > kWgMaxSplit) n = kWgMaxSplit;
      if (n > T) n = T;
      a.prob[i].nsplit = n;
      ctas += n;
    }
    if (ctas > 0) {
      static bool attr[64] = {};
      if (first_on_device(attr)) {
        MVAE_CUDA(cudaFuncSetAttribute(wgrad2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG2_SMEM_FLOATS * 4));
        MVAE_CUDA(cudaFuncSetAttribute(wgrad2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG2_SMEM_FLOATS * 4));
      }
      cudaStream_t s2 = s;
      const int main_pdl = tl_pdl;
      if (thin_side) {           // beside the grouped GEMM: an ordinary launch on the side stream
        MVAE_CUDA(cudaStreamWaitEvent(fork->side, fork->fork_ev, 0));
        s2 = fork->side;
        tl_pdl = 0;
      }
      if (split3) launch_pdl(wgrad2_kernel<true>, dim3(ctas, a.A), dim3(512), (size_t)WG2_SMEM_FLOATS * 4, s2, a);
      else launch_pdl(wgrad2_kernel<false>, dim3(ctas, a.A), dim3(512), (size_t)WG2_SMEM_FLOATS * 4, s2, a);
      if (thin_side) {
        MVAE_CUDA(cudaEventRecord(fork->thin_done, fork->side));
        tl_pdl = main_pdl;
        if (!red_side) {
          MVAE_CUDA(cudaStreamWaitEvent(s, fork->thin_done, 0));
          tl_pdl = 0;
        }
      }
    }
  } else {
    for (int i = 0; i < a.nprob; ++i) a.prob[i].nsplit = a.nsplit;
    if (split3) wgrad_mma_kernel<true><<<dim3(a.nsplit, a.nprob, a.A), 512, 0, s>>>(a);
    else wgrad_mma_kernel<false><<<dim3(a.nsplit, a.nprob, a.A), 512, 0, s>>>(a);
  }
  MVAE_LAUNCH_CHECK();
  int64_t maxn = 0;
  for (int i = 0; i < a.nprob; ++i) {
    int64_t n = (int64_t)a.prob[i].nout * (a.prob[i].nin > 0 ? a.prob[i].nin : 0);
    if (n > maxn) maxn = n;
    if (a.prob[i].nout > maxn) maxn = a.prob[i].nout;
  }
  if (red_side) {
    // the sum of the partials runs beside the fc1 weight gradient: behind both producers on the side stream
    MVAE_CUDA(cudaEventRecord(fork->wide_done, s));
    MVAE_CUDA(cudaStreamWaitEvent(fork->side, fork->wide_done, 0));
    const int main_pdl = tl_pdl;
    tl_pdl = 0;
    launch_pdl(wgrad_reduce2_kernel, dim3((unsigned)((maxn + 255) / 256), a.nprob * 2, a.A), dim3(256), 0, fork->side, a);
    MVAE_CUDA(cudaEventRecord(fork->reduce_done, fork->side));
    tl_pdl = main_pdl;
  } else {
    launch_pdl(wgrad_reduce2_kernel, dim3((unsigned)((maxn + 255) / 256), a.nprob * 2, a.A), dim3(256), 0, s, a);
  }
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae
