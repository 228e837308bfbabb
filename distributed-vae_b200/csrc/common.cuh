// common.cuh — shared host/device helpers for libmixvae_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/mixvae_b200.h"

namespace mvae {

constexpr int kMaxH = 128;   // fc_dim
constexpr int kMaxL = 32;    // lowD_dim
constexpr int kMaxC = 128;   // n_categories
constexpr int kMaxS = 8;     // state_dim
constexpr int kMaxW = 128;   // max width of any narrow layer
constexpr int kMaxPairs = MVAE_MAX_ARMS * (MVAE_MAX_ARMS - 1) / 2;
constexpr int kRowWarps = 8;       // warps per CTA in the row-wise kernels
constexpr int kRowsPerWarp = 4;    // rows a warp processes at once

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern int64_t g_launches;

#define MVAE_CHECK_ARG(cond, ...)                \
  do {                                           \
    if (!(cond)) {                               \
      ::mvae::set_error(__VA_ARGS__);            \
      return -1;                                 \
    }                                            \
  } while (0)

#define MVAE_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::mvae::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                       \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

#define MVAE_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    ::mvae::g_launches++;                                                                \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ::mvae::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),      \
                        __FILE__, __LINE__);                                             \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// parameter tensor indices (per arm) — order of the reference's ModuleLists, nn_model.py:184-208
// ---------------------------------------------------------------------------------------------
enum ParamId {
  FC1_W = 0, FC1_B, FC2_W, FC2_B, FC3_W, FC3_B, FC4_W, FC4_B, FC5_W, FC5_B,
  FCC_W, FCC_B, FCMU_W, FCMU_B, FCSIG_W, FCSIG_B, FC6_W, FC6_B, FC7_W, FC7_B,
  FC8_W, FC8_B, FC9_W, FC9_B, FC10_W, FC10_B, FC11_W, FC11_B
};

// ---------------------------------------------------------------------------------------------
// workspace map (offsets in floats from st->work).  Everything that is per arm is laid out
// [A_local][...]; `acc` is the fp64 accumulator block that is zeroed by memset nodes.
// ---------------------------------------------------------------------------------------------
constexpr int kWgMaxSplit = 32;   // row splits per narrow weight-gradient problem (capacity of Work::wg_part)
struct Work {
  // forward activations kept for backward
  int64_t a[5];        // a1..a4 [A][B][H], a5 [A][B][L]   post-ReLU, pre-BN
  int64_t bn_mean;     // [5][A][128]  batch mean   (training) or running mean (eval)
  int64_t bn_rstd;     // [5][A][128]
  int64_t ysoft;       // [A][B][C]  soft Gumbel sample (== c_smp unless hard)
  int64_t svar;        // [A][B][S]  sigma^2
  int64_t yy;          // [A][B][L+C]  state-head input  [x_low | c_smp]
  int64_t zc;          // [A][B][C+S]  decoder input     [c_smp | dropout(s)]
  int64_t d[5];        // d6 [A][B][L], d7..d10 [A][B][H]
  // backward
  int64_t g_d10;       // [A][B][H]  dLoss/d h10 (post-ReLU), written by the fc11 kernels
  int64_t gtmp[2];     // [A][B][H]  ping-pong gradient wrt layer inputs
  int64_t delta_dec[5];// delta6 [A][B][L], delta7..10 [A][B][H]  (pre-activation grads)
  int64_t delta_mu, delta_sig;  // [A][B][S]
  int64_t delta_z;     // [A][B][C]
  int64_t g_xlow;      // [A][B][L]
  int64_t delta_enc[5];// delta1..4 [A][B][H], delta5 [A][B][L]
  int64_t delta1_t;    // [A][Hpad128][Bpad]  delta1 transposed (tensor-core fc1 dW operand)
  int64_t d10_t;       // [A][Hpad128][Bpad]  h10 transposed     (tensor-core fc11 dW operand)
  int64_t w11_t;       // [A][Hpad128][Dpad]  fc11.weight transposed (tensor-core d h10 operand)
  int64_t rsum;        // [A][B][C]  Gd_a = sum over all arms b of (r_a - r_b), r = log(q+eps)*w
  int64_t colc;        // [A][4][128]  per category: w, cvar, mean, T   (coupling-gradient constants)
  int64_t wcat;        // [At][128]    w of every arm of the model
  int64_t fc1_part;    // [splitk][A][B][Hpad] split-K partials of fc1 (tensor-core path)
  int64_t db_part;     // [8][A][Dpad]  d fc11.bias partials of the gene-owner kernel
  int64_t big;         // [A][B][D]  materialised x_hat / dY, SIMT path only (else -1)
  int64_t wg_part;     // [nsplit][A][wg_floats]  weight-gradient partials of the narrow layers
  // fp64 accumulators (offsets still in floats; 8-byte aligned)
  int64_t acc_fwd, acc_fwd_floats;   // bn_sum[5][A][2][128], (unused q block), kl_sum[A][16]
  int64_t acc_loss, acc_loss_floats; // see accl_* below
  int64_t acc_bwd, acc_bwd_floats;   // bnb_sum[5][A][2][128]
  int64_t keys;        // uint64 [kNumStreams][MVAE_MAX_ARMS] generator keys of the current step + [2] {rng step, Adam step}
  int64_t total;
  int32_t Bpad, Dpad, Hpad, wg_nsplit, wg_rows, fc1_splitk;
  int64_t wg_floats;   // narrow-layer params per arm (contiguous range FC2_W .. FC10_B), see wgrad
};

// sub-offsets inside the fp64 accumulator blocks, in doubles
__host__ __device__ inline int64_t acc_bn(int layer, int A, int a) { return ((int64_t)(layer * A + a)) * 2 * 128; }
__host__ __device__ inline int64_t acc_q(int A, int a) { return (int64_t)5 * A * 256 + (int64_t)a * 256; }
__host__ __device__ inline int64_t acc_kl(int A, int a) { return (int64_t)6 * A * 256 + (int64_t)a * 16; }
__host__ __device__ inline int64_t acc_sync(int A) { return (int64_t)6 * A * 256 + (int64_t)A * 16; }   // grid-barrier counters (8 doubles)
__host__ __device__ inline int64_t acc_fwd_doubles(int A) { return (int64_t)6 * A * 256 + (int64_t)A * 16 + 8; }
// acc_loss block (doubles, fixed capacity MVAE_MAX_ARMS): recon[16][2] | ent[16] | pair[120][2] | T[16][128] | qs[16][2][128]
__host__ __device__ inline int64_t accl_recon(int a) { return (int64_t)a * 2; }
__host__ __device__ inline int64_t accl_ent(int a) { return 32 + (int64_t)a; }
__host__ __device__ inline int64_t accl_pair(int p) { return 48 + (int64_t)p * 2; }
__host__ __device__ inline int64_t accl_T(int a) { return 48 + kMaxPairs * 2 + (int64_t)a * 128; }
__host__ __device__ inline int64_t accl_qs(int a) { return 48 + kMaxPairs * 2 + 16 * 128 + (int64_t)a * 256; }
__host__ __device__ inline int64_t acc_loss_doubles() { return 48 + kMaxPairs * 2 + 16 * 128 + 16 * 256; }
__host__ __device__ inline int64_t accb_bn(int layer, int A, int a) { return ((int64_t)(layer * A + a)) * 2 * 128; }
__host__ __device__ inline int64_t accb_sync(int A) { return (int64_t)5 * A * 256; }   // grid-barrier counters (8 doubles)
__host__ __device__ inline int64_t acc_bwd_doubles(int A) { return (int64_t)5 * A * 256 + 8; }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Inside one API call the library's kernels follow each other on one stream; a
// kernel launched through launch_pdl() right after another kernel of the library carries the programmatic-stream-
// serialization attribute (captured as a programmatic edge in a CUDA graph): its CTAs may become resident -- and run
// the part of the kernel in front of pdl_wait() -- while the kernels before it are still running.  Rules:
//   * every kernel launched through launch_pdl() executes pdl_wait() before it reads anything an earlier kernel of
//     the step wrote and before it writes global memory at all; in front of it only step constants (parameters, x)
//     may be read -- the kernels of a PDL chain cascade, so "earlier" is not just the direct predecessor;
//   * tl_pdl says "the previous operation enqueued by this thread for this step was a kernel of the library on the same
//     stream"; api.cu clears it at every entry point and behind memsets, and keeps the main branch's value across the
//     launches it puts on the side branch.  A kernel launched behind a join (cudaStreamWaitEvent) keeps the attribute: the
//     event edge stays a full dependency and pdl_wait() covers the kernel edge.  Cooperative launches are never
//     secondaries (measured: no gain) but trigger like every other kernel.
// ---------------------------------------------------------------------------------------------
extern thread_local int tl_pdl;
bool pdl_enabled();   // mvae_pdl_enable(0) turns the attribute off (a test compares both ways)
template <typename... KP, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (tl_pdl && pdl_enabled()) ? 1 : 0;
  tl_pdl = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncSetAttribute is per device: "set it once" guards are per device too (one process may drive several GPUs)
inline bool first_on_device(bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// first partial-tile slot of the fc11 gene pass: behind the row pass's slots (CTA + tile < A * ceil(B / 128) + #SM)
inline int64_t f11_gene_slot0(int A, int B) { return (int64_t)A * ((B + 127) / 128) + 160; }

// optional per-group device timing (CUDA events on the launching stream), for bench.py's roofline
enum TimedGroup { TG_FC1_FWD = 0, TG_FC11, TG_FC1_WGRAD, TG_NARROW_FWD, TG_NARROW_BWD, TG_COUPLING, TG_WGRAD, TG_ADAM, TG_COUNT };
bool timing_enabled();
void timing_begin(int group, cudaStream_t s);
void timing_end(int group, cudaStream_t s);
struct TimedScope {
  int g; cudaStream_t s;
  TimedScope(int g_, cudaStream_t s_) : g(g_), s(s_) { timing_begin(g, s); }
  ~TimedScope() { timing_end(g, s); }
};

int compute_layout(const mvae_dims& d, mvae_layout* L);
Work make_work(const mvae_dims& d);

// ---------------------------------------------------------------------------------------------
// generator keys (host and device: mvae_dropout_mask derives them on the host, step_prep_kernel on the device)
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define MVAE_HD __host__ __device__
#else
#define MVAE_HD
#endif
MVAE_HD inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
constexpr uint32_t kStreamDrop = 0, kStreamU = 1, kStreamE = 2, kNumStreams = 3;
MVAE_HD inline uint64_t stream_key(uint64_t seed, uint64_t step, uint32_t arm_global, uint32_t stream) {
  uint64_t k = splitmix64(seed);
  k = splitmix64(k ^ step);
  return splitmix64(k ^ (((uint64_t)stream << 32) | (uint64_t)arm_global));
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// PDL: wait until every kernel this one depends on has completed and its writes are visible / let the dependent kernel's
// CTAs become resident (they block in their own pdl_wait until this grid has completed)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// N independent sums over the 32 lanes at once (N = 2, 4, 8, 16 or 32): recursive halving with the pairing order of
// warp_sum (xor 16, 8, 4, 2, 1), so every total is bit-identical to warp_sum(v[i]).  N - 1 + log2(32/N) shuffles
// in 5 dependent rounds instead of 5 N shuffles in N serial chains.  The total of v[i] is returned in lanes
// (32/N) i .. (32/N) i + 32/N - 1; v is clobbered.
template <int N>
__device__ __forceinline__ float warp_multi_sum(float (&v)[N], int lane) {
  int o = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      const float send = hi ? v[j] : v[j + n / 2];
      const float keep = hi ? v[j + n / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float r = v[0];
#pragma unroll
  for (; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Counter-based generators.  A 64-bit KEY per (seed, step, global arm, stream) is derived by chained splitmix64
// (stream_key below: every input has its own mixing round, so no two (seed, step, arm, stream) tuples share a key short
// of a 64-bit collision); the per-element work is one keyed 32-bit finaliser over the element counter with the two key
// words injected before and between its multiply rounds.
//   stream 0: input dropout, ONE hash per 4 consecutive elements of x (a 16-byte chunk), 8 random bits per element;
//             keep <=> byte >= thresh, thresh = round(p * 256): exact for p = k/256 (the reference default p = 0.5 in
//             particular); other rates are quantised to 1/256.
//   stream 1: Gumbel uniforms U[cell][category];  stream 2: state noise E[cell][s]  (24 random bits, like torch.rand).
__device__ __forceinline__ uint32_t keyed_mix32(uint64_t key, uint64_t ctr) {
  uint32_t h = (uint32_t)ctr * 0x9E3779B1u ^ (uint32_t)(ctr >> 32) * 0x7FEB352Du ^ (uint32_t)key;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= (uint32_t)(key >> 32); h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t drop_bits4(uint64_t key, uint64_t chunk) { return keyed_mix32(key, chunk); }
// Used when the caller passes no noise tensors; the backward regenerates E.
__device__ __forceinline__ float noise_uniform(uint64_t key, uint64_t idx) {
  return (float)(keyed_mix32(key, idx) >> 8) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ bool drop_keep(uint64_t key, int64_t row, int64_t col, int64_t D, uint32_t thresh) {
  const uint64_t idx = (uint64_t)row * (uint64_t)D + (uint64_t)col;
  const uint32_t bits = (drop_bits4(key, idx >> 2) >> (8 * (uint32_t)(idx & 3))) & 0xFFu;
  return bits >= thresh;
}
#endif

struct DropSpec {
  const uint8_t* keep;  // injected mask [B][D] for this launch's arm 0, or nullptr
  int64_t keep_arm_stride;
  const uint64_t* keys; // mode == 2: device table of generator keys, one per LOCAL arm (Work::keys, stream 0)
  float scale;          // 1/(1-p)
  uint32_t thresh16;    // 8-bit threshold of the in-kernel generator (name kept): round(p * 256)
  int mode;             // 0: no dropout, 1: injected mask, 2: in-kernel generator
  int64_t D;            // genes per row (hash index)
  int64_t rows;         // rows of x (cells); tiles may overhang
};

}  // namespace mvae
