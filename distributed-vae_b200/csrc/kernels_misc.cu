// kernels_misc.cu — weight gradients of the narrow layers, fused Adam, small utilities and the
// fp32 SIMT gene-dimension GEMM (used for precision==3 and for shapes the tensor-core path rejects).
#include "common.cuh"
#include "kernels.h"

namespace mvae {

// =============================================================================================
// narrow-layer weight gradients:  dW[j][i] = sum_b delta[b][j] * in[b][i],  db[j] = sum_b delta[b][j]
// grid (split, problem, arm); each CTA reduces `rows_per_split` rows into a partial, a second kernel
// sums the partials in a fixed order (deterministic).
// =============================================================================================
constexpr int WG_RB = 32;  // rows staged per iteration

__global__ void __launch_bounds__(256) wgrad_partial_kernel(const WgArgs p) {
  __shared__ float ds[WG_RB][129];
  __shared__ float is[WG_RB][129];
  __shared__ float bm[128], br[128];
  const WgProblem& pr = p.prob[blockIdx.y];
  const int arm = blockIdx.z, split = blockIdx.x;
  const int tid = threadIdx.x, tj = tid >> 4, ti = tid & 15;
  const int nout = pr.nout, nin = pr.nin;
  const float* delta = p.work + pr.delta_off + (int64_t)arm * pr.delta_arm_stride;
  const float* in = nin > 0 ? p.work + pr.in_off + (int64_t)arm * pr.in_arm_stride : nullptr;
  if (pr.bn_layer >= 0) {
    for (int i = tid; i < nin; i += 256) {
      bm[i] = p.bn_mean[(pr.bn_layer * p.A + arm) * 128 + i];
      br[i] = p.bn_rstd[(pr.bn_layer * p.A + arm) * 128 + i];
    }
  }
  float acc[8][8];
  float accb[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    accb[u] = 0.f;
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[u][v] = 0.f;
  }
  const int r0 = split * p.rows_per_split;
  const int r1 = min(p.B, r0 + p.rows_per_split);
  __syncthreads();
  for (int rb = r0; rb < r1; rb += WG_RB) {
    const int nr = min(WG_RB, r1 - rb);
    for (int idx = tid; idx < WG_RB * nout; idx += 256) {
      const int r = idx / nout, j = idx - r * nout;
      ds[r][j] = r < nr ? delta[(int64_t)(rb + r) * nout + j] : 0.f;
    }
    for (int idx = tid; idx < WG_RB * nin; idx += 256) {
      const int r = idx / nin, i = idx - r * nin;
      float v = 0.f;
      if (r < nr) {
        v = in[(int64_t)(rb + r) * pr.in_ld + i];
        if (pr.bn_layer >= 0) v = (v - bm[i]) * br[i];
      }
      is[r][i] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < WG_RB; ++r) {
      float dv[8], iv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) dv[u] = ds[r][tj + 16 * u];
#pragma unroll
      for (int v = 0; v < 8; ++v) iv[v] = is[r][ti + 16 * v];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        accb[u] += dv[u];
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = fmaf(dv[u], iv[v], acc[u][v]);
      }
    }
    __syncthreads();
  }
  float* part = p.part + (int64_t)split * p.part_split_stride + (int64_t)arm * p.part_arm_stride;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = tj + 16 * u;
    if (j < nout) {
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const int i = ti + 16 * v;
        if (i < nin) part[pr.poffW - p.base_off + (int64_t)j * nin + i] = acc[u][v];
      }
      if (ti == 0) part[pr.poffB - p.base_off + j] = accb[u];
    }
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgArgs p) {
  const WgProblem& pr = p.prob[blockIdx.y >> 1];
  const int which = blockIdx.y & 1, arm = blockIdx.z;
  const int64_t n = which ? pr.nout : (int64_t)pr.nout * pr.nin;
  const int64_t poff = which ? pr.poffB : pr.poffW;
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  const float* part = p.part + (int64_t)arm * p.part_arm_stride + (poff - p.base_off) + e;
  float s = 0.f;
  for (int sp = 0; sp < p.nsplit; ++sp) s += part[(int64_t)sp * p.part_split_stride];
  p.grads[(int64_t)arm * p.g_arm_stride + poff + e] = s;
}

int launch_wgrad(const WgArgs& a, cudaStream_t s) {
  wgrad_partial_kernel<<<dim3(a.nsplit, a.nprob, a.A), 256, 0, s>>>(a);
  MVAE_LAUNCH_CHECK();
  int64_t maxn = 0;
  for (int i = 0; i < a.nprob; ++i) {
    int64_t n = (int64_t)a.prob[i].nout * (a.prob[i].nin > 0 ? a.prob[i].nin : 0);
    if (n > maxn) maxn = n;
    if (a.prob[i].nout > maxn) maxn = a.prob[i].nout;
  }
  wgrad_reduce_kernel<<<dim3((unsigned)((maxn + 255) / 256), a.nprob * 2, a.A), 256, 0, s>>>(a);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// fused Adam over a flat fp32 buffer (torch.optim.Adam, amsgrad=False, maximize=False)
//   m = lerp(m, g, 1-b1); v = b2*v + (1-b2) g^2; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// =============================================================================================
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n4,
                                                   float lr, float b1, float b2, float eps, float wd, int adamw,
                                                   float step_size, float inv_sqrt_bc2, const uint64_t* step_dev) {
  // PDL: parameters and moments are only ever written by this kernel, so the first round of their loads (21 of the 28
  // bytes per parameter) goes out while the kernel that finishes the gradients is still running
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), mv = pv, vv = pv;
  if (i0 < n4) {
    pv = reinterpret_cast<float4*>(p)[i0];
    mv = reinterpret_cast<float4*>(m)[i0];
    vv = reinterpret_cast<float4*>(v)[i0];
  }
  pdl_wait();
  if (step_dev) {   // step counter on the device (replayed CUDA graph): same double-precision bias corrections as the host path
    __shared__ float sh[2];
    if (threadIdx.x == 0) {
      const double st = (double)*step_dev;
      const double bc1 = 1.0 - pow((double)b1, st), bc2 = 1.0 - pow((double)b2, st);
      sh[0] = (float)((double)lr / bc1);
      sh[1] = (float)(1.0 / sqrt(bc2));
    }
    __syncthreads();
    step_size = sh[0];
    inv_sqrt_bc2 = sh[1];
  }
  for (int64_t i = i0; i < n4; i += stride) {
    if (i != i0) {
      pv = reinterpret_cast<float4*>(p)[i];
      mv = reinterpret_cast<float4*>(m)[i];
      vv = reinterpret_cast<float4*>(v)[i];
    }
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = gp[k];
      if (wd != 0.f) {
        if (adamw) pp[k] *= (1.f - lr * wd);
        else gg = fmaf(wd, pp[k], gg);
      }
      mp[k] = fmaf(1.f - b1, gg - mp[k], mp[k]);
      vp[k] = fmaf(vp[k], b2, (1.f - b2) * gg * gg);
      const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
      pp[k] = pp[k] - step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                float wd, int adamw, int64_t step, const uint64_t* step_dev, cudaStream_t s) {
  MVAE_CHECK_ARG(n % 4 == 0, "adam: n=%lld must be a multiple of 4", (long long)n);
  MVAE_CHECK_ARG(step_dev != nullptr || step >= 1, "adam: step must be >= 1");
  float step_size = 0.f, inv_sqrt_bc2 = 0.f;
  if (!step_dev) {
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    step_size = (float)((double)lr / bc1);
    inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  const int64_t n4 = n / 4;
  int gx = (int)((n4 + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  if (gx < 1) gx = 1;
  launch_pdl(adam_kernel, dim3(gx), dim3(256), 0, s, p, g, m, v, n4, lr, b1, b2, eps, wd, adamw, step_size, inv_sqrt_bc2, step_dev);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// Gradient all-reduce fused with Adam over peer memory (data-parallel replicas on NVLink / NVSwitch).
// Rank r owns the float4 range [i0, i1) of the flat buffer: it sums that range of EVERY replica's gradient buffer (peer
// loads, fixed rank order -> every replica ends up with bit-identical parameters), averages, applies Adam with its
// slice of the moments, and stores the new parameters into every replica's parameter buffer (peer stores).  Two-shot
// traffic: (W-1)/W of the buffer in and out per GPU instead of a separate all-reduce pass + a local Adam pass.
// The caller brackets the launch with cross-rank barriers on the stream (gradients final before, parameters final after).
// =============================================================================================
constexpr int kMaxPeers = 16;
struct PeerPtrs { float* p[kMaxPeers]; const float* g[kMaxPeers]; };
__global__ void __launch_bounds__(256) adam_peer_kernel(const PeerPtrs peers, float* __restrict__ m, float* __restrict__ v,
                                                        int64_t i0, int64_t i1, int rank, int world, float inv_world, float lr,
                                                        float b1, float b2, float eps, float step_size, float inv_sqrt_bc2,
                                                        const uint64_t* step_dev) {
  if (step_dev) {
    __shared__ float sh[2];
    if (threadIdx.x == 0) {
      const double st = (double)*step_dev;
      const double bc1 = 1.0 - pow((double)b1, st), bc2 = 1.0 - pow((double)b2, st);
      sh[0] = (float)((double)lr / bc1);
      sh[1] = (float)(1.0 / sqrt(bc2));
    }
    __syncthreads();
    step_size = sh[0];
    inv_sqrt_bc2 = sh[1];
  }
  float4* pl = reinterpret_cast<float4*>(peers.p[rank]);
  for (int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (int64_t)gridDim.x * blockDim.x) {
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < world; ++q) {
      const float4 t = __ldcv(reinterpret_cast<const float4*>(peers.g[q]) + i);   // peer memory: never from a stale cache line
      gv.x += t.x; gv.y += t.y; gv.z += t.z; gv.w += t.w;
    }
    float4 pv = pl[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gg = gp[k] * inv_world;
      mp[k] = fmaf(1.f - b1, gg - mp[k], mp[k]);
      vp[k] = fmaf(vp[k], b2, (1.f - b2) * gg * gg);
      const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
      pp[k] = pp[k] - step_size * (mp[k] / denom);
    }
    for (int q = 0; q < world; ++q) reinterpret_cast<float4*>(peers.p[q])[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  __threadfence_system();
}

int launch_adam_peer(float* const* peer_params, const float* const* peer_grads, float* m, float* v, int64_t n, int rank,
                     int world, float lr, float b1, float b2, float eps, int64_t step, const uint64_t* step_dev,
                     cudaStream_t s) {
  MVAE_CHECK_ARG(n % 4 == 0, "adam_peer: n=%lld must be a multiple of 4", (long long)n);
  MVAE_CHECK_ARG(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "adam_peer: rank %d of %d", rank, world);
  MVAE_CHECK_ARG(step_dev != nullptr || step >= 1, "adam_peer: step must be >= 1");
  PeerPtrs peers;
  memset(&peers, 0, sizeof(peers));
  for (int q = 0; q < world; ++q) {
    MVAE_CHECK_ARG(peer_params[q] && peer_grads[q], "adam_peer: null peer buffer %d", q);
    peers.p[q] = peer_params[q];
    peers.g[q] = peer_grads[q];
  }
  float step_size = 0.f, inv_sqrt_bc2 = 0.f;
  if (!step_dev) {
    const double bc1 = 1.0 - pow((double)b1, (double)step);
    const double bc2 = 1.0 - pow((double)b2, (double)step);
    step_size = (float)((double)lr / bc1);
    inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  }
  const int64_t n4 = n / 4;
  const int64_t i0 = n4 * rank / world, i1 = n4 * (rank + 1) / world;
  int gx = (int)((i1 - i0 + 255) / 256);
  if (gx > 148 * 4) gx = 148 * 4;
  if (gx < 1) gx = 1;
  adam_peer_kernel<<<gx, 256, 0, s>>>(peers, m, v, i0, i1, rank, world, 1.0f / (float)world, lr, b1, b2, eps, step_size,
                                      inv_sqrt_bc2, step_dev);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// Generator keys of one step (common.cuh: stream_key) + the step counters, written to Work::keys.  With device counters
// the kernel increments them first: every launch of a replayed graph then draws fresh noise / uses the next Adam step
// although its arguments never change.
__global__ void step_prep_kernel(uint64_t seed, uint64_t step_host, uint64_t* counters, int bump_adam, int arm_off,
                                 uint64_t* keys_out) {
  __shared__ uint64_t st;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    uint64_t s = step_host;
    if (counters) {
      s = counters[0] + 1;
      counters[0] = s;
      if (bump_adam) counters[1] += 1;
    }
    st = s;
    keys_out[kNumStreams * MVAE_MAX_ARMS] = s;
    keys_out[kNumStreams * MVAE_MAX_ARMS + 1] = counters ? counters[1] : 0;
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < (int)kNumStreams * MVAE_MAX_ARMS)
    keys_out[t] = stream_key(seed, st, (uint32_t)(t % MVAE_MAX_ARMS + arm_off), (uint32_t)(t / MVAE_MAX_ARMS));
}
// =============================================================================================
// Bit-exact expansion of a row-packed sparse batch (bitmap + non-zero values) into the dense [rows][D] fp32 matrix the
// gene kernels stream: the host->device copy of a Smart-seq-shaped batch (35 % non-zeros) moves 36 MB instead of 101 MB.
// One warp per row; per 32-word chunk the lanes popcount their word, scan, and then every word is expanded by all 32
// lanes (lane = bit): coalesced 128-byte stores, value loads contiguous per word.
// =============================================================================================
__global__ void __launch_bounds__(256) unpack_rows_kernel(const uint32_t* __restrict__ bitmap, const float* __restrict__ values,
                                                          const int64_t* __restrict__ row_ptr, int64_t rows, int D, int W,
                                                          float* __restrict__ out, int64_t out_ld) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t below = (1u << lane) - 1u;
  for (int64_t row = warp; row < rows; row += nwarps) {
    const uint32_t* bm = bitmap + row * W;
    const float* v = values + row_ptr[row];
    float* o = out + row * out_ld;
    int base = 0;
    for (int w0 = 0; w0 < W; w0 += 32) {
      const uint32_t word = (w0 + lane < W) ? bm[w0 + lane] : 0u;
      const int cnt = __popc(word);
      int incl = cnt;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += t;
      }
      const int excl = base + incl - cnt;
      base += __shfl_sync(0xffffffffu, incl, 31);
      const int nw = W - w0 < 32 ? W - w0 : 32;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        if (j < nw) {                 // (warp-uniform)
          const uint32_t wj = __shfl_sync(0xffffffffu, word, j);
          const int oj = __shfl_sync(0xffffffffu, excl, j);
          const int col = (w0 + j) * 32 + lane;
          float val = 0.f;
          if ((wj >> lane) & 1u) val = __ldg(v + oj + __popc(wj & below));
          if (col < D) o[col] = val;
        }
      }
    }
  }
}
int launch_unpack_rows(const uint32_t* bitmap, const float* values, const int64_t* row_ptr, int64_t rows, int D, float* out,
                       int64_t out_ld, cudaStream_t s) {
  if (rows == 0) return 0;
  const int W = (D + 31) / 32;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  unpack_rows_kernel<<<(int)blocks, 256, 0, s>>>(bitmap, values, row_ptr, rows, D, W, out, out_ld);
  MVAE_LAUNCH_CHECK();
  return 0;
}

__global__ void counter_inc_kernel(uint64_t* c) { *c += 1; }
int launch_counter_inc(uint64_t* counter, cudaStream_t s) {
  counter_inc_kernel<<<1, 1, 0, s>>>(counter);
  MVAE_LAUNCH_CHECK();
  return 0;
}
int launch_step_prep(uint64_t seed, uint64_t step_host, uint64_t* counters, int bump_adam, int arm_off, uint64_t* keys_out,
                     cudaStream_t s) {
  launch_pdl(step_prep_kernel, dim3(1), dim3(64), 0, s, seed, step_host, counters, bump_adam, arm_off, keys_out);
  MVAE_LAUNCH_CHECK();
  return 0;
}

__global__ void scale_kernel(float* p, int64_t n, const float* scale) {
  const float sc = *scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] *= sc;
}
int launch_scale(float* p, int64_t n, const float* scale_dev, cudaStream_t s) {
  int gx = (int)((n + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  scale_kernel<<<gx, 256, 0, s>>>(p, n, scale_dev);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// argmax per row, first maximal index (mmidas/_utils.py:78 classify == np.argmax)
__global__ void __launch_bounds__(256) argmax_kernel(const float* q, int32_t* labels, int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float best = -INFINITY;
  int bi = 1 << 30;
  for (int k = lane; k < cols; k += 32) {
    const float v = q[row * cols + k];
    if (v > best) { best = v; bi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) labels[row] = bi;
}
int launch_argmax(const float* q, int32_t* labels, int64_t rows, int cols, cudaStream_t s) {
  argmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(q, labels, rows, cols);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// Confusion (co-assignment) counts of the arm pairs: counts[pair][la][lb] += 1 per cell (mmidas/_utils.py:83
// compute_confmat, np.add.at).  labels [A][n] int32; pairs enumerate a < b in order.  Integer work: bit-exact.
__global__ void __launch_bounds__(256) confmat_kernel(const int32_t* labels, int64_t n, int A, int K, int32_t* counts) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  int pair = 0;
  for (int a = 0; a < A; ++a) {
    const int la = labels[(int64_t)a * n + i];
    for (int b = a + 1; b < A; ++b, ++pair) {
      const int lb = labels[(int64_t)b * n + i];
      if (la >= 0 && la < K && lb >= 0 && lb < K) atomicAdd(counts + ((int64_t)pair * K + la) * K + lb, 1);
    }
  }
}
int launch_confmat(const int32_t* labels, int64_t n, int A, int K, int32_t* counts, cudaStream_t s) {
  if (n <= 0) return 0;
  confmat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(labels, n, A, K, counts);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// dst[c][r] = src[r][c]  (batched), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const float* src, int64_t src_ld, int64_t src_bs, float* dst,
                                                        int64_t dst_ld, int64_t dst_bs, int rows, int cols) {
  __shared__ float t[32][33];
  const float* s = src + (int64_t)blockIdx.z * src_bs;
  float* d = dst + (int64_t)blockIdx.z * dst_bs;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    t[i][tx] = (r < rows && c < cols) ? s[(int64_t)r * src_ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) d[(int64_t)c * dst_ld + r] = t[tx][i];
  }
}
int launch_transpose(const float* src, int64_t src_ld, int64_t src_bs, float* dst, int64_t dst_ld, int64_t dst_bs,
                     int rows, int cols, int batch, cudaStream_t s) {
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32, batch), 256, 0, s>>>(src, src_ld, src_bs, dst, dst_ld,
                                                                                 dst_bs, rows, cols);
  MVAE_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// fp32 SIMT GEMM: C[m][n] = sum_k A(m,k) * B(k,n), arbitrary strides, optional dropout on the x operand.
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread.
// =============================================================================================
constexpr int GBM = 64, GBN = 64, GBK = 16;

__device__ __forceinline__ float drop_apply(const DropSpec& d, int arm, int64_t row, int64_t col, float v) {
  if (d.mode == 0) return v;
  bool keep;
  if (d.mode == 1) keep = d.keep[(int64_t)arm * d.keep_arm_stride + row * d.D + col] != 0;
  else keep = drop_keep(d.keys[arm], row, col, d.D, d.thresh16);
  return keep ? v * d.scale : 0.f;
}

__global__ void __launch_bounds__(256) sgemm_simt_kernel(const GemmArgs g) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int batch = blockIdx.z;
  const float* A = g.A + (int64_t)batch * g.A_batch;
  const float* Bm = g.Bm + (int64_t)batch * g.B_batch;
  float* C = g.C + (int64_t)batch * g.C_batch;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int tid = threadIdx.x;
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kcontig = (g.sAk == 1);
  const bool b_kcontig = (g.sBk == 1);
  for (int k0 = 0; k0 < g.K; k0 += GBK) {
#pragma unroll
    for (int t = 0; t < (GBM * GBK) / 256; ++t) {
      const int idx = tid + t * 256;
      int mm, kk;
      if (a_kcontig) { kk = idx % GBK; mm = idx / GBK; } else { mm = idx % GBM; kk = idx / GBM; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < g.K) {
        v = A[(int64_t)m * g.sAm + (int64_t)k * g.sAk];
        if (g.drop_operand == 1) v = drop_apply(g.drop, batch, m, k, v);
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int t = 0; t < (GBN * GBK) / 256; ++t) {
      const int idx = tid + t * 256;
      int nn, kk;
      if (b_kcontig) { kk = idx % GBK; nn = idx / GBK; } else { nn = idx % GBN; kk = idx / GBN; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < g.K) {
        v = Bm[(int64_t)k * g.sBk + (int64_t)n * g.sBn];
        if (g.drop_operand == 2) v = drop_apply(g.drop, batch, k, n, v);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][tm]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tn]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m < g.M) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tn + j;
        if (n < g.N) C[(int64_t)m * g.sCm + (int64_t)n * g.sCn] = acc[i][j];
      }
    }
  }
}

__global__ void dropout_mask_kernel(DropSpec d, uint64_t key, int B, uint8_t* out) {
  const int64_t n = (int64_t)B * d.D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / d.D, col = i - row * d.D;
    out[i] = drop_keep(key, row, col, d.D, d.thresh16) ? 1 : 0;
  }
}
int launch_dropout_mask(const DropSpec& d, uint64_t key, int B, uint8_t* out, cudaStream_t s) {
  dropout_mask_kernel<<<148 * 4, 256, 0, s>>>(d, key, B, out);
  MVAE_LAUNCH_CHECK();
  return 0;
}

int launch_sgemm_simt(const GemmArgs& a, int batch, cudaStream_t s) {
  dim3 grid((a.N + GBN - 1) / GBN, (a.M + GBM - 1) / GBM, batch);
  sgemm_simt_kernel<<<grid, 256, 0, s>>>(a);
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae
