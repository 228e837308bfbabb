// kernels_rows.cu — the narrow (<=128-wide) layers of the cpl-mixVAE step as row-tiled fp32 kernels.
//
// These are HBM/latency-bound: every cell (row) is independent except for the batch statistics,
// which are carried between launches as fp64 column sums accumulated with atomics and finalised in
// the prologue of the consumer.  One warp works on kRowsPerWarp rows at a time with the layer's
// weights resident in shared memory (transposed so that lanes read consecutive addresses).
//
// Reference arithmetic: mmidas/nn_model.py:263-287 (layers), :337-351 (categorical/state heads),
// :413-493 (noise), autograd of the same for the *_bwd kernels.
#include "common.cuh"
#include "kernels.h"

namespace mvae {

// =============================================================================================
// generic narrow dense layer, forward:  out = act( W * bn(in) + b )
// =============================================================================================
template <int KO>
__global__ void __launch_bounds__(kRowWarps * 32) dense_fwd_kernel(const DenseFwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nin = p.nin, nout = p.nout;
  constexpr int P = 32 * KO + 1;
  const int ninp = (nin + 3) & ~3;
  float* Wt = smem;                       // [nin][P]
  float* bias = Wt + nin * P;             // [32*KO]
  float* mean = bias + 32 * KO;           // [nin]
  float* rstd = mean + ninp;              // [nin]
  float* rows = rstd + ninp;              // [warps][rows][ninp]
  float* rows_end = rows + kRowWarps * kRowsPerWarp * ninp;
  double* red = reinterpret_cast<double*>(smem + (((rows_end - smem) + 1) & ~(ptrdiff_t)1));

  const float* W = p.params + (int64_t)arm * p.p_arm_stride + p.offW;
  const float* bsrc = p.params + (int64_t)arm * p.p_arm_stride + p.offB;
  for (int idx = tid; idx < nin * P; idx += blockDim.x) Wt[idx] = 0.f;
  __syncthreads();
  for (int idx = tid; idx < nout * nin; idx += blockDim.x) {
    int j = idx / nin, i = idx - j * nin;
    Wt[i * P + j] = W[idx];
  }
  for (int j = tid; j < 32 * KO; j += blockDim.x) bias[j] = j < nout ? bsrc[j] : 0.f;
  if (p.bn_mode == 1) {
    const double* sums = p.bn_sums_in + (int64_t)arm * 256;
    for (int i = tid; i < nin; i += blockDim.x) {
      double m = sums[i] / (double)p.B;
      double var = sums[128 + i] / (double)p.B - m * m;
      if (var < 0.0) var = 0.0;
      float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)p.eps));
      mean[i] = mf;
      rstd[i] = rf;
      if (blockIdx.x == 0) {
        p.bn_mean[arm * 128 + i] = mf;
        p.bn_rstd[arm * 128 + i] = rf;
      }
    }
  } else if (p.bn_mode == 2) {
    for (int i = tid; i < nin; i += blockDim.x) {
      mean[i] = p.bn_mean[arm * 128 + i];
      rstd[i] = p.bn_rstd[arm * 128 + i];
    }
  }
  __syncthreads();

  const float* in = p.in + (int64_t)arm * p.in_arm_stride;
  float* out = p.out + (int64_t)arm * p.out_arm_stride;
  float* myrows = rows + warp * kRowsPerWarp * ninp;
  double s1[KO], s2[KO];
#pragma unroll
  for (int k = 0; k < KO; ++k) s1[k] = s2[k] = 0.0;

  for (int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp; row0 < p.B; row0 += gridDim.x * kRowWarps * kRowsPerWarp) {
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int row = row0 + r;
      for (int i = lane; i < nin; i += 32) {
        float v = row < p.B ? in[(int64_t)row * nin + i] : 0.f;
        if (p.bn_mode) v = (v - mean[i]) * rstd[i];
        myrows[r * ninp + i] = v;
      }
    }
    __syncwarp();
    float acc[kRowsPerWarp][KO];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
      for (int k = 0; k < KO; ++k) acc[r][k] = 0.f;
#pragma unroll 4
    for (int i = 0; i < nin; ++i) {
      float w[KO];
#pragma unroll
      for (int k = 0; k < KO; ++k) w[k] = Wt[i * P + lane + 32 * k];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const float x = myrows[r * ninp + i];
#pragma unroll
        for (int k = 0; k < KO; ++k) acc[r][k] = fmaf(x, w[k], acc[r][k]);
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int row = row0 + r;
      if (row < p.B) {
#pragma unroll
        for (int k = 0; k < KO; ++k) {
          const int j = lane + 32 * k;
          if (j < nout) {
            float v = acc[r][k] + bias[j];
            if (p.relu) v = fmaxf(v, 0.f);
            out[(int64_t)row * nout + j] = v;
            s1[k] += (double)v;
            s2[k] += (double)v * (double)v;
          }
        }
      }
    }
    __syncwarp();
  }
  if (p.stats_out) {
#pragma unroll
    for (int k = 0; k < KO; ++k) {
      red[(warp * 2 + 0) * 32 * KO + lane + 32 * k] = s1[k];
      red[(warp * 2 + 1) * 32 * KO + lane + 32 * k] = s2[k];
    }
    __syncthreads();
    for (int j = tid; j < 2 * 32 * KO; j += blockDim.x) {
      const int which = j / (32 * KO), jj = j - which * 32 * KO;
      if (jj < nout) {
        double s = 0.0;
        for (int w = 0; w < kRowWarps; ++w) s += red[(w * 2 + which) * 32 * KO + jj];
        atomicAdd(p.stats_out + (int64_t)arm * 256 + which * 128 + jj, s);
      }
    }
  }
}

size_t dense_fwd_smem(int nin, int nout) {
  const int KO = (nout + 31) / 32;
  const int P = 32 * KO + 1;
  const int ninp = (nin + 3) & ~3;
  size_t fl = (size_t)nin * P + 32 * KO + 2 * ninp + kRowWarps * kRowsPerWarp * ninp + 2;
  return fl * 4 + (size_t)kRowWarps * 2 * 32 * KO * 8;
}

int launch_dense_fwd(const DenseFwdArgs& a, int A, cudaStream_t s) {
  const int KO = (a.nout + 31) / 32;
  const size_t smem = dense_fwd_smem(a.nin, a.nout);
  int gx = (a.B + kRowWarps * kRowsPerWarp - 1) / (kRowWarps * kRowsPerWarp);
  if (gx > 296) gx = 296;
  dim3 grid(gx, A);
#define LAUNCH_DF(K)                                                                                   \
  case K: {                                                                                            \
    static bool attr_done[64] = {};                                                                     \
    if (first_on_device(attr_done)) {                                                                                  \
      MVAE_CUDA(cudaFuncSetAttribute(dense_fwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
    }                                                                                                  \
    dense_fwd_kernel<K><<<grid, kRowWarps * 32, smem, s>>>(a);                                         \
  } break;
  switch (KO) {
    LAUNCH_DF(1)
    LAUNCH_DF(2)
    LAUNCH_DF(3)
    LAUNCH_DF(4)
    default:
      set_error("dense_fwd: nout=%d too wide", a.nout);
      return -1;
  }
#undef LAUNCH_DF
  MVAE_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// generic narrow dense layer, backward (data gradient):
//   g   = bn_out ? rstd*(g_out - mean_b(g_out) - n*mean_b(g_out*n)) : g_out
//   dlt = g * 1[act_out > 0]           (stored: operand of the weight-gradient kernel)
//   g_in = dlt * W                      (+ column sums of g_in and g_in*n_in for the next BN backward)
// =============================================================================================
template <int KI>
__global__ void __launch_bounds__(kRowWarps * 32) dense_bwd_kernel(const DenseBwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nin = p.nin, nout = p.nout;
  const int noutp = (nout + 3) & ~3;
  const int Pw = p.g_in ? nin : 0;
  float* Ws = smem;                                 // [nout][nin]
  float* c1 = Ws + ((nout * Pw + 3) & ~3);          // [nout] mean_b(g_out)
  float* c2 = c1 + noutp;                           // [nout] mean_b(g_out*n)
  float* mo = c2 + noutp;                           // [nout] mean of this layer's BN
  float* ro = mo + noutp;                           // [nout] rstd
  float* mi = ro + noutp;                           // [128]  mean of the input BN
  float* ri = mi + 128;                             // [128]
  float* rows = ri + 128;                           // [warps][rows][noutp]
  float* rows_end = rows + kRowWarps * kRowsPerWarp * noutp;
  double* red = reinterpret_cast<double*>(smem + (((rows_end - smem) + 1) & ~(ptrdiff_t)1));

  if (p.g_in) {
    const float* W = p.params + (int64_t)arm * p.p_arm_stride + p.offW;
    for (int idx = tid; idx < nout * nin; idx += blockDim.x) Ws[idx] = W[idx];
  }
  if (p.bn_out) {
    const double* sums = p.bnb_sums + (int64_t)arm * 256;
    for (int j = tid; j < nout; j += blockDim.x) {
      c1[j] = (float)(sums[j] / (double)p.B);
      c2[j] = (float)(sums[128 + j] / (double)p.B);
      mo[j] = p.mean_out[arm * 128 + j];
      ro[j] = p.rstd_out[arm * 128 + j];
    }
  }
  if (p.bn_in) {
    for (int i = tid; i < nin; i += blockDim.x) {
      mi[i] = p.mean_in[arm * 128 + i];
      ri[i] = p.rstd_in[arm * 128 + i];
    }
  }
  __syncthreads();

  const int64_t abo = (int64_t)arm * p.B;
  const float* g_out = p.g_out + abo * nout;
  const float* act_out = p.act_out + abo * nout;
  float* delta = p.delta + abo * nout;
  float* g_in = p.g_in ? p.g_in + abo * nin : nullptr;
  const float* act_in = p.bn_in ? p.act_in + abo * nin : nullptr;
  float* delta_t = p.delta_t ? p.delta_t + (int64_t)arm * p.delta_t_arm_stride : nullptr;
  float* myrows = rows + warp * kRowsPerWarp * noutp;
  double s1[KI], s2[KI];
#pragma unroll
  for (int k = 0; k < KI; ++k) s1[k] = s2[k] = 0.0;

  for (int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp; row0 < p.B; row0 += gridDim.x * kRowWarps * kRowsPerWarp) {
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int row = row0 + r;
      for (int j = lane; j < nout; j += 32) {
        float d = 0.f;
        if (row < p.B) {
          float g = g_out[(int64_t)row * nout + j];
          const float a = act_out[(int64_t)row * nout + j];
          if (p.bn_out) {
            const float n = (a - mo[j]) * ro[j];
            g = ro[j] * (g - c1[j] - n * c2[j]);
          }
          d = a > 0.f ? g : 0.f;
          delta[(int64_t)row * nout + j] = d;
          if (delta_t) delta_t[(int64_t)j * p.delta_t_ld + row] = d;
        }
        myrows[r * noutp + j] = d;
      }
    }
    __syncwarp();
    if (g_in) {
      float acc[kRowsPerWarp][KI];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
        for (int k = 0; k < KI; ++k) acc[r][k] = 0.f;
#pragma unroll 4
      for (int j = 0; j < nout; ++j) {
        float w[KI];
#pragma unroll
        for (int k = 0; k < KI; ++k) {
          const int i = lane + 32 * k;
          w[k] = i < nin ? Ws[j * nin + i] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
          const float d = myrows[r * noutp + j];
#pragma unroll
          for (int k = 0; k < KI; ++k) acc[r][k] = fmaf(d, w[k], acc[r][k]);
        }
      }
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const int row = row0 + r;
        if (row < p.B) {
#pragma unroll
          for (int k = 0; k < KI; ++k) {
            const int i = lane + 32 * k;
            if (i < nin) {
              const float g = acc[r][k];
              g_in[(int64_t)row * nin + i] = g;
              if (p.bn_in) {
                const float n = (act_in[(int64_t)row * nin + i] - mi[i]) * ri[i];
                s1[k] += (double)g;
                s2[k] += (double)g * (double)n;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  }
  if (p.bn_in && g_in) {
#pragma unroll
    for (int k = 0; k < KI; ++k) {
      red[(warp * 2 + 0) * 32 * KI + lane + 32 * k] = s1[k];
      red[(warp * 2 + 1) * 32 * KI + lane + 32 * k] = s2[k];
    }
    __syncthreads();
    for (int j = tid; j < 2 * 32 * KI; j += blockDim.x) {
      const int which = j / (32 * KI), jj = j - which * 32 * KI;
      if (jj < nin) {
        double s = 0.0;
        for (int w = 0; w < kRowWarps; ++w) s += red[(w * 2 + which) * 32 * KI + jj];
        atomicAdd(p.bnb_sums_next + (int64_t)arm * 256 + which * 128 + jj, s);
      }
    }
  }
}

int launch_dense_bwd(const DenseBwdArgs& a, int A, cudaStream_t s) {
  const int nin_eff = a.g_in ? a.nin : 1;
  const int KI = (nin_eff + 31) / 32;
  const int noutp = (a.nout + 3) & ~3;
  size_t fl = (size_t)((a.nout * (a.g_in ? a.nin : 0) + 3) & ~3) + 4 * noutp + 256 + kRowWarps * kRowsPerWarp * noutp + 2;
  size_t smem = fl * 4 + (size_t)kRowWarps * 2 * 32 * KI * 8;
  int gx = (a.B + kRowWarps * kRowsPerWarp - 1) / (kRowWarps * kRowsPerWarp);
  if (gx > 296) gx = 296;
  dim3 grid(gx, A);
#define LAUNCH_DB(K)                                                                                   \
  case K: {                                                                                            \
    static bool attr_done[64] = {};                                                                     \
    if (first_on_device(attr_done)) {                                                                                  \
      MVAE_CUDA(cudaFuncSetAttribute(dense_bwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
    }                                                                                                  \
    dense_bwd_kernel<K><<<grid, kRowWarps * 32, smem, s>>>(a);                                         \
  } break;
  switch (KI) {
    LAUNCH_DB(1)
    LAUNCH_DB(2)
    LAUNCH_DB(3)
    LAUNCH_DB(4)
    default:
      set_error("dense_bwd: nin=%d too wide", a.nin);
      return -1;
  }
#undef LAUNCH_DB
  MVAE_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// fc1 epilogue: a1 = relu(sum_k partial_k + b1), column sums for batch_l1.
// partials: [nsplit][A][ldp rows][ldc] (tensor-core path, padded) or [1][A][B][H] (SIMT path).
// =============================================================================================
__global__ void __launch_bounds__(256) fc1_epilogue_kernel(const Fc1EpiArgs p) {
  __shared__ double red[8][2][128];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* bias = p.params + (int64_t)arm * p.p_arm_stride + p.offB;
  float* out = p.out + (int64_t)arm * p.B * p.H;
  double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  for (int row = blockIdx.x * 8 + warp; row < p.B; row += gridDim.x * 8) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = lane + 32 * k;
      if (j < p.H) {
        float v = 0.f;
        for (int sp = 0; sp < p.nsplit; ++sp)
          v += p.part[(int64_t)sp * p.split_stride + (int64_t)arm * p.arm_stride + (int64_t)row * p.ld + j];
        v = fmaxf(v + bias[j], 0.f);
        out[(int64_t)row * p.H + j] = v;
        s1[k] += (double)v;
        s2[k] += (double)v * (double)v;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    red[warp][0][lane + 32 * k] = s1[k];
    red[warp][1][lane + 32 * k] = s2[k];
  }
  __syncthreads();
  {
    const int which = tid >> 7, j = tid & 127;
    if (j < p.H) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w][which][j];
      atomicAdd(p.stats_out + (int64_t)arm * 256 + which * 128 + j, s);
    }
  }
}

int launch_fc1_epilogue(const Fc1EpiArgs& a, int A, cudaStream_t s) {
  int gx = (a.B + 7) / 8;
  if (gx > 296) gx = 296;
  fc1_epilogue_kernel<<<dim3(gx, A), 256, 0, s>>>(a);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// BatchNorm bookkeeping
// =============================================================================================
// eval mode: mean/rstd from the running statistics (nn.BatchNorm1d in .eval()).
__global__ void bn_eval_prep_kernel(const float* bn_running, int64_t bn_stride, const BnOff bn_off,
                                    float* bn_mean, float* bn_rstd, int A, int H, int L, float eps) {
  const int layer = blockIdx.x, arm = blockIdx.y;
  const int n = layer < 4 ? H : L;
  const float* rm = bn_running + (int64_t)arm * bn_stride + bn_off.off[layer];
  const float* rv = rm + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    bn_mean[(layer * A + arm) * 128 + i] = rm[i];
    bn_rstd[(layer * A + arm) * 128 + i] = 1.0f / sqrtf(rv[i] + eps);
  }
}

// training mode: running = (1-m)*running + m*batch  (unbiased variance), num_batches_tracked += 1
__global__ void bn_update_running_kernel(float* bn_running, int64_t bn_stride, const BnOff bn_off,
                                         int64_t* nbt, const double* bn_sums, int A, int B, int H, int L,
                                         float momentum) {
  const int layer = blockIdx.x, arm = blockIdx.y;
  const int n = layer < 4 ? H : L;
  float* rm = bn_running + (int64_t)arm * bn_stride + bn_off.off[layer];
  float* rv = rm + n;
  const double* sums = bn_sums + acc_bn(layer, A, arm);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = sums[i] / (double)B;
    double var = sums[128 + i] / (double)B - m * m;
    if (var < 0.0) var = 0.0;
    const float mean_f = (float)m;
    const float var_u = (float)(var * (double)B / (double)(B - 1));
    rm[i] = (1.0f - momentum) * rm[i] + momentum * mean_f;
    rv[i] = (1.0f - momentum) * rv[i] + momentum * var_u;
  }
  if (threadIdx.x == 0) nbt[arm * 6 + layer] += 1;
}

int launch_bn_eval_prep(const float* bn_running, int64_t bn_stride, BnOff off, float* bn_mean,
                        float* bn_rstd, int A, int H, int L, float eps, cudaStream_t s) {
  bn_eval_prep_kernel<<<dim3(5, A), 128, 0, s>>>(bn_running, bn_stride, off, bn_mean, bn_rstd, A, H, L, eps);
  MVAE_LAUNCH_CHECK();
  return 0;
}
int launch_bn_update_running(float* bn_running, int64_t bn_stride, BnOff off, int64_t* nbt,
                             const double* bn_sums, int A, int B, int H, int L, float momentum, cudaStream_t s) {
  bn_update_running_kernel<<<dim3(5, A), 128, 0, s>>>(bn_running, bn_stride, off, nbt, bn_sums, A, B, H, L, momentum);
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae
