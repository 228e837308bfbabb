// kernels_head.cu — categorical head (double softmax + Gumbel-softmax), Gaussian state head with
// reparameterisation, and the first decoder layer, one warp per cell; forward and backward.
//
// Reference: mmidas/nn_model.py:269 (fcc+softmax), :337 (softmax(./tau)), :430-493 (Gumbel-softmax,
// straight-through hard sample), :347-351 + :413-428 (state head, UNIFORM reparameterisation noise),
// :277-280 (state dropout, concat, fc6).  Backward = autograd of the same, plus the coupling /
// entropy / KL gradients of mixVAE_model.loss (:550-569), see DESIGN.md for the formulas.
#include "common.cuh"
#include "kernels.h"

namespace mvae {

constexpr int KC = 4;          // categories per lane (C <= 128)
constexpr int CP = 32 * KC + 1;

struct HeadSmem {
  float* WcT;   // [L][CP]
  float* bc;    // [128]
  float* Wmu;   // [S][L+C]
  float* Wsig;  // [S][L+C]
  float* W6;    // [L][C+S]
  float* bmu;   // [8]
  float* bsig;  // [8]
  float* b6;    // [32]
  float* mean5; // [32]
  float* rstd5; // [32]
  float* colc;  // [4][128]
};

__device__ __forceinline__ HeadSmem head_carve(float* smem, int L, int C, int S) {
  HeadSmem h;
  float* p = smem;
  h.WcT = p; p += L * CP;
  h.bc = p; p += 128;
  h.Wmu = p; p += S * (L + C);
  h.Wsig = p; p += S * (L + C);
  h.W6 = p; p += L * (C + S);
  h.bmu = p; p += 8;
  h.bsig = p; p += 8;
  h.b6 = p; p += 32;
  h.mean5 = p; p += 32;
  h.rstd5 = p; p += 32;
  h.colc = p; p += 4 * 128;
  return h;
}
static size_t head_smem_bytes(int L, int C, int S) {
  return sizeof(float) * ((size_t)L * CP + 128 + 2 * S * (L + C) + L * (C + S) + 8 + 8 + 32 + 32 + 32 + 4 * 128);
}

__device__ __forceinline__ void head_load_weights(const HeadArgs& p, const HeadSmem& h, int arm, bool with_colc) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int L = p.L, C = p.C, S = p.S;
  const float* base = p.params + (int64_t)arm * p.p_arm_stride;
  for (int idx = tid; idx < L * CP; idx += nt) h.WcT[idx] = 0.f;
  __syncthreads();
  for (int idx = tid; idx < C * L; idx += nt) {
    const int k = idx / L, i = idx - k * L;
    h.WcT[i * CP + k] = base[p.oWc + idx];
  }
  for (int k = tid; k < 128; k += nt) h.bc[k] = k < C ? base[p.oBc + k] : 0.f;
  for (int idx = tid; idx < S * (L + C); idx += nt) {
    h.Wmu[idx] = base[p.oWmu + idx];
    h.Wsig[idx] = base[p.oWsig + idx];
  }
  for (int idx = tid; idx < L * (C + S); idx += nt) h.W6[idx] = base[p.oW6 + idx];
  for (int s = tid; s < S; s += nt) {
    h.bmu[s] = base[p.oBmu + s];
    h.bsig[s] = base[p.oBsig + s];
  }
  for (int i = tid; i < L; i += nt) h.b6[i] = base[p.oB6 + i];
  if (with_colc) {
    pdl_wait();        // the coupling constants come from loss_finalize_kernel; the weights above are step constants
    for (int idx = tid; idx < 4 * 128; idx += nt) h.colc[idx] = p.colc[(int64_t)arm * 4 * 128 + idx];
  }
}

// fc6 of one cell: lane i < L returns relu(W6[i,:C] . c + W6[i,C:] . s + b6[i]); c is spread over the lanes
// (category lane + 32 k in c[k]).  Same operation order per output as a warp_sum per row of W6.
template <int NL>
__device__ __forceinline__ float fc6_rows(const HeadSmem& h, const float (&c)[KC], const float (&sdp)[kMaxS], int L, int C,
                                          int S, int lane) {
  float part[NL];
  const int CS = C + S;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    part[i] = 0.f;
    if (i < L) {
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = lane + 32 * k;
        if (kk < C) part[i] = fmaf(h.W6[i * CS + kk], c[k], part[i]);
      }
    }
  }
  const float tot = warp_multi_sum<NL>(part, lane);
  float v = __shfl_sync(0xffffffffu, tot, (lane * (32 / NL)) & 31);
  float d6 = 0.f;
  if (lane < L) {
#pragma unroll
    for (int s = 0; s < kMaxS; ++s) if (s < S) v = fmaf(h.W6[lane * CS + C + s], sdp[s], v);
    d6 = fmaxf(v + h.b6[lane], 0.f);
  }
  return d6;
}
// fcc backward of one cell: lane i < L returns sum_k gq[k] WcT[i][k]
template <int NL>
__device__ __forceinline__ float fcc_bwd_rows(const HeadSmem& h, const float (&gq)[KC], int L, int lane) {
  float part[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    part[i] = 0.f;
    if (i < L) {
#pragma unroll
      for (int k = 0; k < KC; ++k) part[i] = fmaf(gq[k], h.WcT[i * CP + lane + 32 * k], part[i]);
    }
  }
  const float tot = warp_multi_sum<NL>(part, lane);
  return __shfl_sync(0xffffffffu, tot, (lane * (32 / NL)) & 31);
}

// =============================================================================================
// forward
// =============================================================================================
// NL: reduction width of the L-row products, 16 (lowD_dim <= 16) or 32.  C96: n_categories > 96, so that only the last of
// the four category slots of a lane (lane + 96) needs its `< C` guard -- the other three guards fold away.
// LC / SC: compile-time lowD_dim / state_dim (0: run-time values); the reference's 10 / 2 get their own instance.
template <int NL, bool C96, int LC, int SC>
__global__ void __launch_bounds__(kRowWarps * 32, 3) head_fwd_kernel(const HeadArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = LC ? LC : p.L, C = p.C, S = SC ? SC : p.S, B = p.B;
  if (C96) {
    __builtin_assume(C > 96);
    __builtin_assume(C <= 128);
  }
  const HeadSmem h = head_carve(smem, L, C, S);
  pdl_trigger();
  head_load_weights(p, h, arm, false);     // PDL: the weight staging overlaps the tail of the encoder chain
  pdl_wait();
  if (p.bn_mode == 1) {
    const double* sums = p.bn_sums5 + (int64_t)arm * 256;
    for (int i = tid; i < L; i += blockDim.x) {
      const double m = sums[i] / (double)B;
      double var = sums[128 + i] / (double)B - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)p.eps));
      h.mean5[i] = mf;
      h.rstd5[i] = rf;
      if (blockIdx.x == 0) {
        p.bn_mean5[arm * 128 + i] = mf;
        p.bn_rstd5[arm * 128 + i] = rf;
      }
    }
  } else {
    for (int i = tid; i < L; i += blockDim.x) {
      h.mean5[i] = p.bn_mean5[arm * 128 + i];
      h.rstd5[i] = p.bn_rstd5[arm * 128 + i];
    }
  }
  __syncthreads();
  if (p.bn_running && blockIdx.x == 0) {
    // running = (1 - m) running + m batch (unbiased variance), num_batches_tracked += 1: the batch sums of all five
    // BatchNorms are final when this kernel starts
    for (int layer = 0; layer < 5; ++layer) {
      const int n = layer < 4 ? p.H : L;
      float* rm = p.bn_running + (int64_t)arm * p.bn_stride + p.bn_off.off[layer];
      float* rv = rm + n;
      const double* sums = p.bn_sums_all + acc_bn(layer, p.A, arm);
      for (int i = tid; i < n; i += blockDim.x) {
        const double m = sums[i] / (double)B;
        double var = sums[128 + i] / (double)B - m * m;
        if (var < 0.0) var = 0.0;
        const float mean_f = (float)m;
        const float var_u = (float)(var * (double)B / (double)(B - 1));
        rm[i] = (1.0f - p.momentum) * rm[i] + p.momentum * mean_f;
        rv[i] = (1.0f - p.momentum) * rv[i] + p.momentum * var_u;
      }
      if (tid == 0) p.nbt[arm * 6 + layer] += 1;
    }
  }

  const int64_t ab = (int64_t)arm * B;
  __shared__ double klacc_sm[kRowWarps][kMaxS];   // per-warp KL sums (lane 0 only): kept out of the register file
  double* klacc = klacc_sm[warp];
  if (lane < kMaxS) klacc[lane] = 0.0;
  __syncwarp();

  // category mask of the pruning path (nn_model.py:332-335): q = softmax over the kept categories only, 0 elsewhere
  bool kq[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int kk = lane + 32 * k;
    kq[k] = kk < C && (p.cat_mask == nullptr || p.cat_mask[kk] != 0);
  }

  for (int row = blockIdx.x * kRowWarps + warp; row < B; row += gridDim.x * kRowWarps) {
    const int64_t r = ab + row;
    // ---- x_low = batch_l5(relu(fc5)) : nn_model.py:268
    float xl = 0.f;
    if (lane < L) xl = (p.a5[r * L + lane] - h.mean5[lane]) * h.rstd5[lane];
    // ---- z = fcc(x_low), p = softmax(z) : :269
    float z[KC], pk[KC], q[KC], y[KC], c[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) z[k] = h.bc[lane + 32 * k];
    for (int i = 0; i < L; ++i) {
      const float xi = __shfl_sync(0xffffffffu, xl, i);
#pragma unroll
      for (int k = 0; k < KC; ++k) z[k] = fmaf(h.WcT[i * CP + lane + 32 * k], xi, z[k]);
    }
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KC; ++k) if (lane + 32 * k < C) m = fmaxf(m, z[k]);
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      pk[k] = (lane + 32 * k < C) ? expf(z[k] - m) : 0.f;
      sum += pk[k];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int k = 0; k < KC; ++k) pk[k] = pk[k] / sum;
    // ---- q = softmax(p / tau) : :337
    m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      z[k] = pk[k] / p.tau;
      if (kq[k]) m = fmaxf(m, z[k]);
    }
    m = warp_max(m);
    sum = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      q[k] = kq[k] ? expf(z[k] - m) : 0.f;
      sum += q[k];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int k = 0; k < KC; ++k) q[k] = q[k] / sum;
    // ---- Gumbel-softmax sample : :430-455
    if (p.training) {
      m = -INFINITY;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = lane + 32 * k;
        if (kk < C) {
          const float u = p.U ? p.U[r * C + kk]
                              : noise_uniform(p.ukeys[arm], (uint64_t)row * C + kk);
          const float g = -logf(-logf(u + p.eps) + p.eps);
          z[k] = (logf(q[k] + p.eps) + g) / p.temp;
          m = fmaxf(m, z[k]);
        }
      }
      m = warp_max(m);
      sum = 0.f;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        y[k] = (lane + 32 * k < C) ? expf(z[k] - m) : 0.f;
        sum += y[k];
      }
      sum = warp_sum(sum);
#pragma unroll
      for (int k = 0; k < KC; ++k) y[k] = y[k] / sum;
    } else {
#pragma unroll
      for (int k = 0; k < KC; ++k) y[k] = q[k];
    }
    // ---- straight-through one-hot : :486-493 (always in eval, nn_model.py:341-343)
    if (p.hard || !p.training) {
      float vm = -INFINITY;
#pragma unroll
      for (int k = 0; k < KC; ++k) if (lane + 32 * k < C) vm = fmaxf(vm, y[k]);
      vm = warp_max(vm);
      int idx = 1 << 30;
#pragma unroll
      for (int k = 0; k < KC; ++k) if (lane + 32 * k < C && y[k] == vm) idx = min(idx, lane + 32 * k);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) idx = min(idx, __shfl_xor_sync(0xffffffffu, idx, o));
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const float yh = (lane + 32 * k == idx) ? 1.f : 0.f;
        c[k] = (yh - y[k]) + y[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < KC; ++k) c[k] = y[k];
    }
    // ---- category stores (here: p, q, y are dead afterwards)
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int kk = lane + 32 * k;
      if (kk < C) {
        p.c_prob[r * C + kk] = pk[k];
        p.qc[r * C + kk] = q[k];
        p.c_smp[r * C + kk] = c[k];
        p.ysoft[r * C + kk] = y[k];
        p.yy[r * (L + C) + L + kk] = c[k];
        p.zc[r * (C + S) + kk] = c[k];
      }
    }
    // ---- state head on [x_low | c_smp] : :347-351
    float sdp[kMaxS];
#pragma unroll
    for (int s = 0; s < kMaxS; ++s) {
      sdp[s] = 0.f;
      if (s < S) {
        float pm = 0.f, ps = 0.f;
        if (lane < L) {
          pm = h.Wmu[s * (L + C) + lane] * xl;
          ps = h.Wsig[s * (L + C) + lane] * xl;
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int kk = lane + 32 * k;
          if (kk < C) {
            pm = fmaf(h.Wmu[s * (L + C) + L + kk], c[k], pm);
            ps = fmaf(h.Wsig[s * (L + C) + L + kk], c[k], ps);
          }
        }
        const float mu = warp_sum(pm) + h.bmu[s];
        const float us = warp_sum(ps) + h.bsig[s];
        const float var = 1.f / (1.f + expf(-us));
        const float lv = logf(var + p.eps);
        const float elv = expf(lv);
        const float sd = sqrtf(elv);
        const float e = p.E ? p.E[r * S + s] : noise_uniform(p.ekeys[arm], (uint64_t)row * S + s);
        const float smp = e * sd + mu;                     // uniform noise, nn_model.py:427
        float sd_in = smp;
        if (p.training && p.keep_s) sd_in = p.keep_s[r * S + s] ? smp * p.s_scale : 0.f;
        sdp[s] = sd_in;
        if (lane == 0) {
          p.s_mean[r * S + s] = mu;
          p.s_logvar[r * S + s] = lv;
          p.s_smp[r * S + s] = smp;
          p.svar[r * S + s] = var;
          p.zc[r * (C + S) + C + s] = sd_in;
          klacc[s] += (double)(1.f + lv - mu * mu - elv);
        }
      }
    }
    // ---- d6 = relu(fc6([c_smp | dropout(s)])) : :278-280   (the L dot products are reduced together)
    float d6 = 0.f;
    d6 = fc6_rows<NL>(h, c, sdp, L, C, S, lane);
    // ---- stores
    if (lane < L) {
      p.x_low[r * L + lane] = xl;
      p.yy[r * (L + C) + lane] = xl;
      p.d6[r * L + lane] = d6;
    }
  }
  if (p.kl_sums && lane == 0) {
#pragma unroll
    for (int s = 0; s < kMaxS; ++s)
      if (s < S) atomicAdd(p.kl_sums + (int64_t)arm * 16 + s, klacc[s]);
  }
}

// =============================================================================================
// backward
// =============================================================================================
template <int NL, bool C96, int LC, int SC>
__global__ void __launch_bounds__(kRowWarps * 32, 3) head_bwd_kernel(const HeadArgs p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double red[kRowWarps][2][32];
  const int arm = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = LC ? LC : p.L, C = p.C, S = SC ? SC : p.S, B = p.B;
  if (C96) {
    __builtin_assume(C > 96);
    __builtin_assume(C <= 128);
  }
  const HeadSmem h = head_carve(smem, L, C, S);
  pdl_trigger();
  head_load_weights(p, h, arm, true);      // (waits for the previous kernels before it reads the coupling constants)
  __syncthreads();
  const float* cw = h.colc;            // w
  const float* ccv = h.colc + 128;     // (var+eps)^-1.5 / (B-1)
  const float* cmean = h.colc + 256;   // column mean of q
  const float* cT = h.colc + 384;      // sum_b G*log q
  const int64_t ab = (int64_t)arm * B;
  double s1 = 0.0, s2 = 0.0;

  for (int row = blockIdx.x * kRowWarps + warp; row < B; row += gridDim.x * kRowWarps) {
    const int64_t r = ab + row;
    // ---- fc6 backward
    float dl6 = 0.f, xl = 0.f;
    if (lane < L) {
      dl6 = p.d6[r * L + lane] > 0.f ? p.g_d6[r * L + lane] : 0.f;
      p.delta6[r * L + lane] = dl6;
      xl = p.x_low[r * L + lane];
    }
    float gc[KC], gsd[kMaxS];
#pragma unroll
    for (int k = 0; k < KC; ++k) gc[k] = 0.f;
#pragma unroll
    for (int s = 0; s < kMaxS; ++s) gsd[s] = 0.f;
    for (int i = 0; i < L; ++i) {
      const float di = __shfl_sync(0xffffffffu, dl6, i);
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int kk = lane + 32 * k;
        if (kk < C) gc[k] = fmaf(di, h.W6[i * (C + S) + kk], gc[k]);
      }
#pragma unroll
      for (int s = 0; s < kMaxS; ++s) if (s < S) gsd[s] = fmaf(di, h.W6[i * (C + S) + C + s], gsd[s]);
    }
    // ---- state head backward (+ KL gradient)
    float gx = 0.f;
#pragma unroll
    for (int s = 0; s < kMaxS; ++s) {
      if (s < S) {
        float gs = gsd[s];
        if (p.keep_s) gs = p.keep_s[r * S + s] ? gs * p.s_scale : 0.f;
        const float mu = p.s_mean[r * S + s], lv = p.s_logvar[r * S + s], var = p.svar[r * S + s];
        const float e = p.E ? p.E[r * S + s] : noise_uniform(p.ekeys[arm], (uint64_t)row * S + s);
        const float elv = expf(lv), sd = sqrtf(elv);
        const float gmu = gs + p.kl_coef * mu;
        const float glv = gs * e * 0.5f * sd + p.kl_coef * (-0.5f) * (1.f - elv);
        const float gvar = glv / (var + p.eps);
        const float dsg = gvar * var * (1.f - var);
        if (lane == 0) {
          p.delta_mu[r * S + s] = gmu;
          p.delta_sig[r * S + s] = dsg;
        }
        if (lane < L) gx = fmaf(gmu, h.Wmu[s * (L + C) + lane], fmaf(dsg, h.Wsig[s * (L + C) + lane], gx));
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int kk = lane + 32 * k;
          if (kk < C) gc[k] = fmaf(gmu, h.Wmu[s * (L + C) + L + kk], fmaf(dsg, h.Wsig[s * (L + C) + L + kk], gc[k]));
        }
      }
    }
    // ---- Gumbel-softmax backward (straight-through: d c_smp / d y = 1), coupling + entropy
    float q[KC], gq[KC];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int kk = lane + 32 * k;
      q[k] = kk < C ? p.qc[r * C + kk] : 0.f;
      gq[k] = kk < C ? p.ysoft[r * C + kk] : 0.f;   // y
      dot = fmaf(gq[k], gc[k], dot);
    }
    dot = warp_sum(dot);
    float dot2 = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int kk = lane + 32 * k;
      if (kk < C) {
        const float dl = gq[k] * (gc[k] - dot);                 // d loss / d logits of y
        const float qe = q[k] + p.eps;
        const float lq = logf(qe);
        const float G = p.g_coef * p.gdiff[r * C + kk];             // sum_b (r_a - r_b), coupling_rows_kernel
        float g = dl / (p.temp * qe);
        g += G * cw[kk] / qe - cT[kk] * ccv[kk] * (q[k] - cmean[kk]);
        g += p.ent_coef * (lq + q[k] / qe);
        gq[k] = g;
        dot2 = fmaf(q[k], g, dot2);
      } else {
        gq[k] = 0.f;
      }
    }
    dot2 = warp_sum(dot2);
    float pk[KC], dot3 = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int kk = lane + 32 * k;
      pk[k] = kk < C ? p.c_prob[r * C + kk] : 0.f;
      gq[k] = q[k] * (gq[k] - dot2) / p.tau;                    // d loss / d p
      dot3 = fmaf(pk[k], gq[k], dot3);
    }
    dot3 = warp_sum(dot3);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int kk = lane + 32 * k;
      gq[k] = pk[k] * (gq[k] - dot3);                           // delta_z
      if (kk < C) p.delta_z[r * C + kk] = gq[k];
    }
    // ---- fcc backward into x_low (the L dot products are reduced together)
    {
      const float part = fcc_bwd_rows<NL>(h, gq, L, lane);
      if (lane < L) gx += part;
    }
    if (lane < L) {
      p.g_xlow[r * L + lane] = gx;
      s1 += (double)gx;
      s2 += (double)gx * (double)xl;
    }
  }
  red[warp][0][lane] = s1;
  red[warp][1][lane] = s2;
  __syncthreads();
  if (tid < 64) {
    const int which = tid >> 5, i = tid & 31;
    if (i < L) {
      double s = 0.0;
      for (int w = 0; w < kRowWarps; ++w) s += red[w][which][i];
      atomicAdd(p.bnb_sums5 + (int64_t)arm * 256 + which * 128 + i, s);
    }
  }
}

// persistent: at most 3 resident CTAs per SM over all arms, each warp loops over its cells (the weights
// are loaded once per CTA)
static int head_grid(int B, int A) {
  int gx = (B + kRowWarps - 1) / kRowWarps;
  const int cap = (148 * 3) / (A > 0 ? A : 1);
  return gx > cap ? (cap > 0 ? cap : 1) : gx;
}

template <typename F>
static int head_set_attr(F* f) {
  MVAE_CUDA(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  return 0;
}
static int head_attrs() {
  static bool attr_done[64] = {};
  if (first_on_device(attr_done)) {
#define HEAD_ATTR(...)                                                     \
  if (int rc = head_set_attr(head_fwd_kernel<__VA_ARGS__>)) return rc;     \
  if (int rc = head_set_attr(head_bwd_kernel<__VA_ARGS__>)) return rc;
    HEAD_ATTR(16, false, 0, 0) HEAD_ATTR(16, true, 0, 0) HEAD_ATTR(32, false, 0, 0) HEAD_ATTR(32, true, 0, 0)
    HEAD_ATTR(16, false, 10, 2) HEAD_ATTR(16, true, 10, 2)
#undef HEAD_ATTR
  }
  return 0;
}

#define HEAD_DISPATCH(KERNEL)                                                                       \
  do {                                                                                              \
    const size_t sm = head_smem_bytes(a.L, a.C, a.S);                                               \
    const bool c96 = a.C > 96;                                                                      \
    if (a.L == 10 && a.S == 2) {                                                                    \
      if (c96) launch_pdl(KERNEL<16, true, 10, 2>, grid, dim3(kRowWarps * 32), sm, s, a);                         \
      else launch_pdl(KERNEL<16, false, 10, 2>, grid, dim3(kRowWarps * 32), sm, s, a);                            \
    } else if (a.L <= 16) {                                                                         \
      if (c96) launch_pdl(KERNEL<16, true, 0, 0>, grid, dim3(kRowWarps * 32), sm, s, a);                          \
      else launch_pdl(KERNEL<16, false, 0, 0>, grid, dim3(kRowWarps * 32), sm, s, a);                             \
    } else {                                                                                        \
      if (c96) launch_pdl(KERNEL<32, true, 0, 0>, grid, dim3(kRowWarps * 32), sm, s, a);                          \
      else launch_pdl(KERNEL<32, false, 0, 0>, grid, dim3(kRowWarps * 32), sm, s, a);                             \
    }                                                                                               \
  } while (0)

int launch_head_fwd(const HeadArgs& a, cudaStream_t s) {
  if (int rc = head_attrs()) return rc;
  const dim3 grid(head_grid(a.B, a.A), a.A);
  HEAD_DISPATCH(head_fwd_kernel);
  MVAE_LAUNCH_CHECK();
  return 0;
}

int launch_head_bwd(const HeadArgs& a, cudaStream_t s) {
  if (int rc = head_attrs()) return rc;
  const dim3 grid(head_grid(a.B, a.A), a.A);
  HEAD_DISPATCH(head_bwd_kernel);
  MVAE_LAUNCH_CHECK();
  return 0;
}
#undef HEAD_DISPATCH

}  // namespace mvae
