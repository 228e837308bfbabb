// kernels_loss.cu — coupling / entropy terms of mixVAE_model.loss, loss finalisation, and the
// element-wise reconstruction loss used when the reconstruction is materialised.
//
// Reference: mmidas/nn_model.py:39-86 (helpers), :539-598 (loss).  All batch reductions are fp64
// atomics into a block that mvae_loss zeroes first.
#include "common.cuh"
#include "kernels.h"

namespace mvae {

constexpr int KC = 4;

// ---------------------------------------------------------------------------------------------
// column sums of q(c|x) of every arm: sum_b q, sum_b q^2  -> inv_var (nn_model.py:75-77)
// ---------------------------------------------------------------------------------------------
// 4 row groups x 128 columns per CTA, one wave of CTAs: the loads of a thread's rows are independent (unrolled), the
// row groups are combined in shared memory and each CTA issues ONE fp64 atomic per column and moment.
__global__ void __launch_bounds__(512) qstats_kernel(const CouplingArgs p) {
  __shared__ double red[3][2][128];
  const int arm = blockIdx.y;  // global arm index
  const int k = threadIdx.x & 127, rg = threadIdx.x >> 7;
  const float* q = p.qc_all + (int64_t)arm * p.B * p.C;
  const int rows_per = (p.B + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per, r1 = min(p.B, r0 + rows_per);
  double s1 = 0.0, s2 = 0.0;
  if (k < p.C) {
#pragma unroll 4
    for (int r = r0 + rg; r < r1; r += 4) {
      const double v = (double)q[(int64_t)r * p.C + k];
      s1 += v;
      s2 += v * v;
    }
  }
  if (rg > 0) {
    red[rg - 1][0][k] = s1;
    red[rg - 1][1][k] = s2;
  }
  __syncthreads();
  if (rg == 0 && k < p.C && r1 > r0) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      s1 += red[g][0][k];
      s2 += red[g][1][k];
    }
    atomicAdd(p.acc + accl_qs(arm) + k, s1);
    atomicAdd(p.acc + accl_qs(arm) + 128 + k, s2);
  }
}

int launch_qstats(const CouplingArgs& a, cudaStream_t s) {
  int gx = (a.B + 15) / 16;
  const int cap = (2 * 148) / (a.At > 0 ? a.At : 1);
  if (gx > cap) gx = cap > 0 ? cap : 1;
  qstats_kernel<<<dim3(gx, a.At), 512, 0, s>>>(a);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// per cell: r_a = log(q_a+eps)*w_a, pairwise ||r_a-r_b||^2 and ||c_a-c_b||^2, per-arm entropy,
// and for the local arms Gd_a = sum_b (r_a - r_b) and T_a[k] = sum_cells G_a[cell,k]*log(q_a[cell,k]+eps),
// G_a = (2 lam / B) Gd_a   (gradient of the distance through inv_var).
// Gd_a is formed from differences against arm 0 (d_b = r_b - r_0, S = sum_b d_b, Gd_a = A d_a - S): categories that no
// arm uses have r = log(eps) / sqrt(eps) ~ -1.8e5 in EVERY arm, and A r_a - sum_b r_b in fp32 would round at that
// magnitude while the true difference is tiny (it showed up as 3e-4 relative error in the encoder gradients at A = 3).
// ---------------------------------------------------------------------------------------------
// ATC: compile-time number of arms (0: run-time); C96: n_categories > 96 (only the last category slot of a lane is guarded)
// One CTA of kCoupWarps warps per SM: every CTA ends with ~200 fp64 atomics on the same 200 addresses, whose cost grows
// with the number of CTAs (444 CTAs of 8 warps spent a third of the kernel there).
constexpr int kCoupWarps = 24;
template <int ATC, bool C96>
__global__ void __launch_bounds__(kCoupWarps * 32) coupling_rows_kernel(const CouplingArgs p) {
  __shared__ float w[MVAE_MAX_ARMS][128];
  extern __shared__ float sTw[];            // [kCoupWarps][A local arms][128]: every lane owns its categories of its warp's slice
  __shared__ double sPair[kMaxPairs][2];
  __shared__ double sEnt[MVAE_MAX_ARMS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int At = ATC ? ATC : p.At, B = p.B, C = p.C;
  if (C96) {
    __builtin_assume(C > 96);
    __builtin_assume(C <= 128);
  }
  pdl_trigger();       // PDL: the fc11 row pass may start filling its rings with x and fc11.weight tiles
  pdl_wait();
  for (int idx = tid; idx < At * 128; idx += blockDim.x) {
    const int a = idx >> 7, k = idx & 127;
    float wv = 0.f;
    if (k < C) {
      const double s1 = p.acc[accl_qs(a) + k], s2 = p.acc[accl_qs(a) + 128 + k];
      const double mean = s1 / (double)B;
      double var = (s2 - s1 * mean) / (double)(B - 1);      // unbiased, torch.var default
      if (var < 0.0) var = 0.0;
      wv = (float)sqrt(1.0 / (var + (double)p.eps));
      if (blockIdx.x == 0) p.wcat[a * 128 + k] = wv;
    }
    w[a][k] = wv;
  }
  for (int idx = tid; idx < kCoupWarps * p.A * 128; idx += blockDim.x) sTw[idx] = 0.f;
  for (int idx = tid; idx < kMaxPairs * 2; idx += blockDim.x) (&sPair[0][0])[idx] = 0.0;
  if (tid < MVAE_MAX_ARMS) sEnt[tid] = 0.0;
  __syncthreads();

  const float gcoef = 2.f * p.lam / (float)B;
  for (int row = blockIdx.x * kCoupWarps + warp; row < B; row += gridDim.x * kCoupWarps) {
    float r0[KC] = {0.f, 0.f, 0.f, 0.f}, rs[KC] = {0.f, 0.f, 0.f, 0.f};
    float lqc[ATC ? ATC : 1][KC];      // compile-time arm count: log(q + eps) of pass 1 kept for pass 2 (same values)
    // pass 1: r_0 and S = sum_b (r_b - r_0) ; entropy per arm
    for (int a = 0; a < At; ++a) {
      const float* q = p.qc_all + ((int64_t)a * B + row) * C;
      float ent = 0.f;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = lane + 32 * k;
        if (kk < C) {
          const float qv = q[kk];
          const float lq = logf(qv + p.eps);
          if (ATC) lqc[ATC ? a : 0][k] = lq;
          const float rv = lq * w[a][kk];
          if (a == 0) r0[k] = rv;
          else rs[k] += rv - r0[k];
          ent = fmaf(qv, lq, ent);
        }
      }
      ent = warp_sum(ent);
      if (lane == 0) atomicAdd(&sEnt[a], (double)ent);
    }
    // pass 2: pairs and T (q rows are L1/L2 resident)
    int pair = 0;
    for (int a = 0; a < At; ++a) {
      const float* qa = p.qc_all + ((int64_t)a * B + row) * C;
      const float* ca = p.csmp_all + ((int64_t)a * B + row) * C;
      float ra[KC], ya[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = lane + 32 * k;
        ra[k] = 0.f; ya[k] = 0.f;
        if (kk < C) {
          const float lq = ATC ? lqc[ATC ? a : 0][k] : logf(qa[kk] + p.eps);
          ra[k] = lq * w[a][kk];
          ya[k] = ca[kk];
          const int la = a - p.arm_off;
          if (la >= 0 && la < p.A) {
            const float Gd = (float)At * (ra[k] - r0[k]) - rs[k];
            p.gdiff[((int64_t)la * B + row) * C + kk] = Gd;
            sTw[(warp * p.A + la) * 128 + kk] += gcoef * Gd * lq;   // a handful of rows per warp: fp32 here, fp64 across warps
          }
        }
      }
      for (int b = a + 1; b < At; ++b, ++pair) {
        const float* qb = p.qc_all + ((int64_t)b * B + row) * C;
        const float* cb = p.csmp_all + ((int64_t)b * B + row) * C;
        float dist = 0.f, l2 = 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int kk = lane + 32 * k;
          if (kk < C) {
            const float rb = (ATC ? lqc[ATC ? b : 0][k] : logf(qb[kk] + p.eps)) * w[b][kk];
            const float d = ra[k] - rb;
            dist = fmaf(d, d, dist);
            const float e = ya[k] - cb[kk];
            l2 = fmaf(e, e, l2);
          }
        }
        dist = warp_sum(dist);
        l2 = warp_sum(l2);
        if (lane == 0) {
          atomicAdd(&sPair[pair][0], (double)dist);
          atomicAdd(&sPair[pair][1], (double)l2);
        }
      }
    }
  }
  __syncthreads();
  const int npairs = At * (At - 1) / 2;
  for (int idx = tid; idx < npairs * 2; idx += blockDim.x) atomicAdd(p.acc + accl_pair(0) + idx, (&sPair[0][0])[idx]);
  if (tid < At) atomicAdd(p.acc + accl_ent(tid), sEnt[tid]);
  for (int idx = tid; idx < p.A * 128; idx += blockDim.x) {
    const int a = idx >> 7, k = idx & 127;
    if (k < C) {
      double t = 0.0;
#pragma unroll
      for (int wp = 0; wp < kCoupWarps; ++wp) t += (double)sTw[(wp * p.A + a) * 128 + k];
      atomicAdd(p.acc + accl_T(a) + k, t);
    }
  }
}

int launch_coupling_rows(const CouplingArgs& a, cudaStream_t s) {
  int gx = (a.B + kCoupWarps - 1) / kCoupWarps;
  if (gx > 148) gx = 148;              // one CTA per SM, a couple of rows per warp
  const size_t dyn = (size_t)kCoupWarps * a.A * 128 * sizeof(float);
  const bool c96 = a.C > 96;
  const void* fn = a.At == 2 ? (c96 ? (const void*)coupling_rows_kernel<2, true> : (const void*)coupling_rows_kernel<2, false>)
                             : (c96 ? (const void*)coupling_rows_kernel<0, true> : (const void*)coupling_rows_kernel<0, false>);
  MVAE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  CouplingArgs arg = a;
  void* args[] = {(void*)&arg};
  MVAE_CUDA(cudaLaunchKernel(fn, dim3(gx), dim3(kCoupWarps * 32), args, dyn, s));   // (At == 2: compile-time arm loops)
  tl_pdl = 1;          // (an ordinary launch -- it follows the join with the side branch -- that can be a PDL primary)
  MVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// loss vector (nn_model.py:579-598) + per-category constants of the coupling gradient
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) loss_finalize_kernel(const LossFinalArgs p) {
  const int tid = threadIdx.x;
  const int A = p.A, At = p.At, B = p.B, C = p.C;
  const double Bd = (double)B;
  // coupling-gradient constants for the local arms
  for (int idx = tid; idx < A * 128; idx += blockDim.x) {
    const int a = idx >> 7, k = idx & 127;
    float wv = 0.f, cv = 0.f, mn = 0.f, T = 0.f;
    if (k < C) {
      const int ga = a + p.arm_off;
      const double s1 = p.acc_loss[accl_qs(ga) + k], s2 = p.acc_loss[accl_qs(ga) + 128 + k];
      const double mean = s1 / Bd;
      double var = (s2 - s1 * mean) / (Bd - 1.0);
      if (var < 0.0) var = 0.0;
      const double ve = var + (double)p.eps;
      wv = (float)sqrt(1.0 / ve);
      cv = (float)(1.0 / (ve * sqrt(ve)) / (Bd - 1.0));
      mn = (float)mean;
      T = (float)p.acc_loss[accl_T(a) + k];
    }
    p.colc[(int64_t)a * 512 + k] = wv;
    p.colc[(int64_t)a * 512 + 128 + k] = cv;
    p.colc[(int64_t)a * 512 + 256 + k] = mn;
    p.colc[(int64_t)a * 512 + 384 + k] = T;
  }
  // stage every accumulator the scalar part reads in shared memory first: one round of independent loads by all threads
  // instead of ~30 serial global loads by thread 0 (this kernel sits on the critical path between loss and backward)
  __shared__ double sh_rec[MVAE_MAX_ARMS][2], sh_kl[MVAE_MAX_ARMS][kMaxS], sh_pair[kMaxPairs][2], sh_ent[MVAE_MAX_ARMS];
  const int npairs_all = At * (At - 1) / 2;
  for (int idx = tid; idx < A * 2; idx += blockDim.x) sh_rec[idx >> 1][idx & 1] = p.acc_loss[accl_recon(idx >> 1) + (idx & 1)];
  for (int idx = tid; idx < A * p.S; idx += blockDim.x) sh_kl[idx / p.S][idx % p.S] = p.kl_sums[(int64_t)(idx / p.S) * 16 + idx % p.S];
  for (int idx = tid; idx < npairs_all * 2; idx += blockDim.x) sh_pair[idx >> 1][idx & 1] = p.acc_loss[accl_pair(idx >> 1) + (idx & 1)];
  for (int idx = tid; idx < At; idx += blockDim.x) sh_ent[idx] = p.acc_loss[accl_ent(idx)];
  __syncthreads();
  if (tid == 0) {
    const double log2pi = 1.8378770664093453;
    float* out = p.loss_out;
    for (int i = 0; i < 5 + 3 * At; ++i) out[i] = 0.f;
    double sum_ind = 0.0;
    for (int a = 0; a < A; ++a) {
      const int ga = a + p.arm_off;
      const double sse = sh_rec[a][0], mism = sh_rec[a][1];
      const double numel = Bd * (double)p.D;
      const float rec = (float)(0.5 * sse / Bd + 0.5 * (100.0 * mism / numel));
      const float ll = (float)(sse / numel + Bd * log2pi);
      double kl = 0.0;
      for (int s = 0; s < p.S; ++s) kl += -0.5 * (sh_kl[a][s] / Bd);
      out[5 + ga] = rec;
      out[5 + At + ga] = (float)kl;
      out[5 + 2 * At + ga] = ll;
      sum_ind += (double)rec + (double)p.beta * kl;
    }
    const int npairs = npairs_all;
    double sum_dist = 0.0, sum_l2 = 0.0, sum_ent = 0.0;
    for (int i = 0; i < npairs; ++i) {
      sum_dist += sh_pair[i][0] / Bd;
      sum_l2 += sh_pair[i][1] / Bd;
    }
    for (int a = 0; a < At; ++a) sum_ent += (double)(At - 1) * sh_ent[a] / Bd;
    const double np_f = npairs > 1 ? (double)npairs : 1.0;
    const double joint = (double)p.lam * sum_dist + sum_ent +
                         np_f * (((double)C / 2.0) * log2pi - 0.5 * log(2.0 * (double)p.lam));
    const double scale = At - 1 > 1 ? (double)(At - 1) : 1.0;
    out[0] = (float)(scale * sum_ind + joint);
    out[1] = (float)joint;
    if (npairs > 0) {
      out[2] = (float)(sum_ent / npairs);
      out[3] = (float)(sum_dist / npairs);
      out[4] = (float)(sum_l2 / npairs);
    }
  }
}

int launch_loss_finalize(const LossFinalArgs& a, cudaStream_t s) {
  loss_finalize_kernel<<<1, 128, 0, s>>>(a);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// element-wise reconstruction loss on a materialised pre-activation (nn_model.py:287, :542-546)
//   x_hat = relu(pre + b11);  sse += (x_hat-x)^2;  mism += [x_hat>0.1] != [x>0.1]
//   dY = gscale * (x_hat - x) * [x_hat > 0]   (the BCE half has no gradient: its inputs are constants)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) recon_elem_kernel(const ReconElemArgs p) {
  __shared__ double red[8][2];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* pre = p.pre + (int64_t)arm * p.B * p.D;
  const float* x = p.x + (int64_t)arm * p.x_arm_stride;
  const float* bias = p.params + (int64_t)arm * p.p_arm_stride + p.offB;
  float* xr = p.x_rec ? p.x_rec + (int64_t)arm * p.B * p.D : nullptr;
  double sse = 0.0, mism = 0.0;
  const int64_t n = (int64_t)p.B * p.D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + tid; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = idx / p.D;
    const int col = (int)(idx - row * p.D);
    const float xv = x[row * p.x_row_stride + col];
    const float xh = fmaxf(pre[idx] + bias[col], 0.f);
    const float d = xh - xv;
    sse += (double)d * (double)d;
    mism += ((xh > 0.1f) != (xv > 0.1f)) ? 1.0 : 0.0;
    if (xr) xr[idx] = xh;
    if (p.want_grad) pre[idx] = xh > 0.f ? p.gscale * d : 0.f;
  }
  sse = warp_sum(sse);
  mism = warp_sum(mism);
  if (lane == 0) {
    red[warp][0] = sse;
    red[warp][1] = mism;
  }
  __syncthreads();
  if (tid < 2 && p.recon_acc) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][tid];
    atomicAdd(p.recon_acc + accl_recon(arm) + tid, s);
  }
}

int launch_recon_elem(const ReconElemArgs& a, int A, cudaStream_t s) {
  int64_t n = (int64_t)a.B * a.D;
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (gx > 148 * 8) gx = 148 * 8;
  if (gx < 1) gx = 1;
  recon_elem_kernel<<<dim3(gx, A), 256, 0, s>>>(a);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// column sums of a [B][D] matrix per arm (d fc11.bias = sum_b dY)
__global__ void __launch_bounds__(256) colsum_kernel(const float* src, int64_t src_arm_stride, float* dst_base,
                                                     int64_t dst_arm_stride, int B, int D) {
  __shared__ float red[8][32];
  const int arm = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const float* s = src + (int64_t)arm * src_arm_stride;
  float acc = 0.f;
  if (col < D)
    for (int r = warp; r < B; r += 8) acc += s[(int64_t)r * D + col];
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && col < D) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    dst_base[(int64_t)arm * dst_arm_stride + col] = t;
  }
}

int launch_colsum(const float* src, int64_t src_arm_stride, float* dst_base, int64_t dst_arm_stride, int B, int D,
                  int A, cudaStream_t s) {
  colsum_kernel<<<dim3((D + 31) / 32, A), 256, 0, s>>>(src, src_arm_stride, dst_base, dst_arm_stride, B, D);
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae
