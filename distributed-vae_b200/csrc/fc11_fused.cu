// fc11_fused.cu — the last decoder layer fused with the reconstruction loss and its own backward.
//
// Reference ops (mmidas/nn_model.py): x_hat = relu(fc11(h10)) :287; 0.5*mse_sum/B + 0.5*BCE(bin(x_hat), bin(x))
// :542-546; autograd of both.  The gradient of the reconstruction term w.r.t. x_hat depends only on x_hat and
// x (dY = max(A-1,1)/B * (x_hat - x) * [x_hat > 0]; the BCE half acts on constants), so forward, loss and the
// first backward GEMM run in ONE pass over x and x_hat never reaches HBM as an activation:
//
//   fc11_rows_kernel ("row owner", CTA = 128 cells x a range of genes, loop over 32-gene tiles):
//     MMA1  X[128 x 32]   = h10[128 x H] . W11[32 x H]^T          (A resident in smem, accumulator in TMEM, 2 buffers)
//     epi   x_hat = relu(X + b11); sse += (x_hat-x)^2; mism += [x_hat>.1] != [x>.1]; dY -> overwrites the x tile
//     MMA2  G[128 x H]   += dY[128 x 32] . W11[32 x H]            (A = the dY tile in smem, B = W11 tile read MN-major)
//   G (= d loss / d h10) is written as split partials and summed in a fixed order.
//
//   fc11_genes_kernel = the same kernel with the roles of h10 and W11 swapped ("gene owner", CTA = 128 genes x
//   a range of cells, loop over 32-cell tiles): MMA1 X^T[128 genes x 32 cells] = W11 . h10^T, the epilogue
//   thread owns a gene (bias and d fc11.bias are per-thread scalars), reads the x tile transposed, writes
//   dY^T over it, MMA2 accumulates d fc11.weight[128 genes x H] += dY^T . h10 in TMEM across all cells.
//   x_hat is recomputed instead of storing dY: one more pass over x, no [B,D] intermediate at all.
//
// Warp roles as in gemm_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.  Three mbarrier
// rings: full/empty (TMA <-> MMA2), xhat_full/tmem_empty (MMA1 <-> epilogue), dy_ready (epilogue -> MMA2).
#include "gemm_tc.h"
#include "tc_common.cuh"

namespace mvae {

namespace {
using namespace tc;

constexpr int GN = 32;                 // genes per tile
constexpr int ROWS = 128;              // cells per CTA
constexpr int STAGES = 3;
constexpr int H10_BYTES = 4 * 16384;   // h10 block, K-major, 4 slabs of 128 rows x 128 B (H <= 128)
constexpr int WK_BYTES = 4 * 4096;     // W11 tile, K-major image: 4 h-slabs of 32 genes x 128 B
constexpr int WM_BYTES = 4 * 4096;     // W11 tile, MN-major image (32-byte-atom swizzle): 4 h-slabs of 32 genes x 128 B
constexpr int X_BYTES = 16384;         // x tile / dY tile: 128 rows x 128 B
constexpr int STAGE_BYTES = WK_BYTES + WM_BYTES + X_BYTES;
constexpr int FUSED_THREADS = 320;      // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quadrant)
constexpr int EPI_THREADS = 256;

struct RowsArgs {
  int B, D, H;
  int HN;                    // H rounded up to 16 (UMMA N of MMA2)
  int tiles_per_split, ntiles;
  int x_batched;
  float gscale;
  int want_grad;
  const float* bias; int64_t bias_arm_stride;          // fc11.bias
  float* dY; int64_t dy_arm_stride;                    // [A][B][D] (interim: feeds the dW11 GEMM) or nullptr
  float* x_rec; int64_t xrec_arm_stride;               // optional materialised reconstruction
  float* part; int64_t part_split_stride, part_arm_stride;   // d h10 partials [split][A][Bpad][128]
  double* recon_acc;                                   // acc_loss block
  float* db_part; int64_t db_split_stride, db_arm_stride;    // gene owner: d fc11.bias partials [split][A][Dpad]
};

template <bool GENE>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
fc11_fused_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWk,
                 const __grid_constant__ CUtensorMap tmWm, const __grid_constant__ CUtensorMap tmX, const RowsArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* h10s = smem;
  auto wk = [&](int s) { return smem + H10_BYTES + (size_t)s * STAGE_BYTES; };
  auto wm = [&](int s) { return smem + H10_BYTES + (size_t)s * STAGE_BYTES + WK_BYTES; };
  auto xs = [&](int s) { return smem + H10_BYTES + (size_t)s * STAGE_BYTES + WK_BYTES + WM_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + H10_BYTES + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full = bars;                   // [STAGES] TMA landed
  uint64_t* empty = bars + STAGES;         // [STAGES] MMA2 finished reading the stage
  uint64_t* dy_ready = bars + 2 * STAGES;  // [STAGES] epilogue wrote dY into the x tile
  uint64_t* xhat_full = bars + 3 * STAGES; // [2] MMA1 result in TMEM buffer b
  uint64_t* tmem_empty = xhat_full + 2;    // [2] epilogue drained TMEM buffer b
  uint64_t* h10_full = tmem_empty + 2;
  uint64_t* g_full = h10_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_full + 1);

  const int split = blockIdx.x, m0 = blockIdx.y * ROWS, arm = blockIdx.z;
  const int t0 = split * a.tiles_per_split;
  const int t1 = min(a.ntiles, t0 + a.tiles_per_split);
  const int nt = max(t1 - t0, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
      mbar_init(dy_ready + s, EPI_THREADS);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(xhat_full + b, 1);
      mbar_init(tmem_empty + b, EPI_THREADS);
    }
    mbar_init(h10_full, 1);
    mbar_init(g_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_g = tmem_base + 64;          // d h10 accumulator: columns [64, 64+HN)

  if (warp == 0) {
    // ===== TMA producer (whole warp runs the uniform loop; one elected lane issues) =====
    const int xb = a.x_batched ? arm : 0;
    if (elect_one()) {
      mbar_expect_tx(h10_full, H10_BYTES);
#pragma unroll
      for (int j = 0; j < 4; ++j) tma_load_3d(&tmH, h10_full, h10s + j * 16384, 32 * j, m0, arm);
    }
    __syncwarp();
    for (int i = 0; i < nt; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(empty + s, ph ^ 1);
      const int g0 = (t0 + i) * GN;    // first gene (row owner) / first cell (gene owner) of the tile
      if (elect_one()) {
        mbar_expect_tx(full + s, STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_3d(&tmWk, full + s, wk(s) + j * 4096, 32 * j, g0, arm);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_3d(&tmWm, full + s, wm(s) + j * 4096, 32 * j, g0, arm);
        if (!GENE) {
          tma_load_3d(&tmX, full + s, xs(s), g0, m0, xb);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(&tmX, full + s, xs(s) + j * 4096, m0 + 32 * j, g0, xb);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the uniform loop, one elected lane issues (operands stay in uniform registers) =====
    {
      const uint32_t idesc1 = make_idesc(128, GN, false, false);
      const uint32_t idesc2 = make_idesc(128, a.HN, false, true);
      const int ksteps1 = (a.H + 7) / 8;
      const uint32_t h10a = smem_u32(h10s);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t tg = tb + 64;
      mbar_wait(h10_full, 0);
      uint32_t gacc = 0;
      auto mma2 = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(dy_ready + s, (j / STAGES) & 1);
        tc_fence_after();
        const uint32_t xa = smem_u32(xs(s)), wma = smem_u32(wm(s));
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < GN / 8; ++ks)
            umma_tf32(tg, make_smem_desc(xa + ks * 32, 0, 1024, false), make_smem_desc(wma + ks * 1024, 4096, 512, true),
                      idesc2, (gacc | (uint32_t)ks) ? 1u : 0u);
          umma_commit(empty + s);
        }
        __syncwarp();
        gacc = 1;
      };
      auto release = [&](int j) {          // no second MMA: the stage is free once its epilogue is done
        const int sj = j % STAGES;
        mbar_wait(dy_ready + sj, (j / STAGES) & 1);
        if (elect_one()) umma_commit(empty + sj);
        __syncwarp();
      };
      for (int i = 0; i < nt; ++i) {
        const int s = i % STAGES, b = i & 1;
        mbar_wait(full + s, (i / STAGES) & 1);
        mbar_wait(tmem_empty + b, ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t wka = smem_u32(wk(s));
        if (elect_one()) {
          for (int ks = 0; ks < ksteps1; ++ks) {
            const uint32_t off = (uint32_t)(ks >> 2), sub = (uint32_t)(ks & 3) * 32;
            umma_tf32(tb + b * GN, make_smem_desc(h10a + off * 16384 + sub, 0, 1024, false),
                      make_smem_desc(wka + off * 4096 + sub, 0, 1024, false), idesc1, ks > 0 ? 1u : 0u);
          }
          umma_commit(xhat_full + b);
        }
        __syncwarp();
        if (i >= 1) {
          if (a.want_grad) mma2(i - 1); else release(i - 1);
        }
      }
      if (nt > 0) {
        if (a.want_grad) mma2(nt - 1); else release(nt - 1);
      }
      if (elect_one()) umma_commit(g_full);
      __syncwarp();
    }
  } else if (!GENE) {
    // ===== epilogue warps 2..9, row owner: thread = (cell, half of the 32 genes of a tile) =====
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;            // row within the CTA tile == TMEM lane
    const int row = m0 + r;
    const bool row_ok = row < a.B;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float* __restrict__ bias = a.bias + (int64_t)arm * a.bias_arm_stride;
    float* xrrow = a.x_rec ? a.x_rec + (int64_t)arm * a.xrec_arm_stride + (int64_t)row * a.D : nullptr;
    double sse = 0.0, mism = 0.0;
    for (int i = 0; i < nt; ++i) {
      const int s = i % STAGES, b = i & 1;
      const int g0 = (t0 + i) * GN + 16 * half;         // first gene of this thread's 16
      const bool full_tile = g0 + 16 <= a.D;
      // bias first: its latency overlaps the barrier waits
      float4 bvv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int g = g0 + 4 * c;
        bvv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full_tile) bvv[c] = __ldg(reinterpret_cast<const float4*>(bias + g));
        else {
          if (g < a.D) bvv[c].x = __ldg(bias + g);
          if (g + 1 < a.D) bvv[c].y = __ldg(bias + g + 1);
          if (g + 2 < a.D) bvv[c].z = __ldg(bias + g + 2);
          if (g + 3 < a.D) bvv[c].w = __ldg(bias + g + 3);
        }
      }
      mbar_wait(full + s, (i / STAGES) & 1);            // x tile visible to this thread
      mbar_wait(xhat_full + b, (i >> 1) & 1);
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld16(tmem_base + lane_addr + (uint32_t)(b * GN + 16 * half), acc);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tmem_empty + b);
      float4* xrow = reinterpret_cast<float4*>(xs(s) + r * 128);
      float fs = 0.f, fm = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int p = (4 * half + c) ^ (r & 7);           // SWIZZLE_128B: logical chunk lives at chunk p
        const float4 xv = xrow[p];
        const float xin[4] = {xv.x, xv.y, xv.z, xv.w};
        const float bb[4] = {bvv[c].x, bvv[c].y, bvv[c].z, bvv[c].w};
        float dy[4], xh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xh[e] = fmaxf(__uint_as_float(acc[4 * c + e]) + bb[e], 0.f);
          const float d = xh[e] - xin[e];
          fs = fmaf(d, d, fs);                            // masked below for partial tiles
          fm += ((xh[e] > 0.1f) != (xin[e] > 0.1f)) ? 1.f : 0.f;
          dy[e] = xh[e] > 0.f ? a.gscale * d : 0.f;
        }
        if (!(row_ok && full_tile)) {                     // slow path: overhanging rows / genes
          const int g = g0 + 4 * c;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (!(row_ok && g + e < a.D)) {
              const float d = xh[e] - xin[e];
              fs -= d * d;
              fm -= ((xh[e] > 0.1f) != (xin[e] > 0.1f)) ? 1.f : 0.f;
              dy[e] = 0.f;
            }
          }
        }
        xrow[p] = make_float4(dy[0], dy[1], dy[2], dy[3]);
        if (xrrow && row_ok) {
          const int g = g0 + 4 * c;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (g + e < a.D) xrrow[g + e] = xh[e];
        }
      }
      sse += (double)fs;
      mism += (double)fm;
      fence_proxy_async();
      mbar_arrive(dy_ready + s);
    }
    // ---- loss partial sums
    sse = warp_sum(sse);
    mism = warp_sum(mism);
    if (lane == 0 && a.recon_acc) {
      atomicAdd(a.recon_acc + accl_recon(arm), sse);
      atomicAdd(a.recon_acc + accl_recon(arm) + 1, mism);
    }
    // ---- d h10 partial: the two halves share the HN columns
    if (a.want_grad) {
      mbar_wait(g_full, 0);
      tc_fence_after();
      float* prow = a.part + (int64_t)split * a.part_split_stride + (int64_t)arm * a.part_arm_stride + (int64_t)row * 128;
      for (int c0 = 64 * half; c0 < min(a.HN, 64 * half + 64); c0 += 16) {
        uint32_t rr[16];
        if (nt > 0) {
          tmem_ld16(tmem_g + lane_addr + (uint32_t)c0, rr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) rr[e] = 0u;
        }
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            reinterpret_cast<float4*>(prow + c0)[e] = make_float4(__uint_as_float(rr[4 * e]), __uint_as_float(rr[4 * e + 1]),
                                                                  __uint_as_float(rr[4 * e + 2]), __uint_as_float(rr[4 * e + 3]));
        }
      }
    }
  } else {
    // ===== epilogue warps 2..9, gene owner: thread = (gene, half of the 32 cells of a tile) =====
    __shared__ float dbs[128];
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int gl = quad * 32 + lane;           // gene within the CTA block == TMEM lane
    const int gene = m0 + gl;
    const bool gene_ok = gene < a.D;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float bj = gene_ok ? __ldg(a.bias + (int64_t)arm * a.bias_arm_stride + gene) : 0.f;
    float dbsum = 0.f;
    for (int i = 0; i < nt; ++i) {
      const int s = i % STAGES, b = i & 1;
      const int r0 = (t0 + i) * GN + 16 * half;     // first cell of this thread's 16
      mbar_wait(full + s, (i / STAGES) & 1);
      mbar_wait(xhat_full + b, (i >> 1) & 1);
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld16(tmem_base + lane_addr + (uint32_t)(b * GN + 16 * half), acc);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tmem_empty + b);
      // x tile: 4 gene slabs of [32 cells x 128 B] (SWIZZLE_128B); this thread's gene is element `lane` of slab `quad`
      const uint8_t* xslab = xs(s) + quad * 4096;
      float xv[16];
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const int r = 16 * half + rr;
        xv[rr] = *reinterpret_cast<const float*>(xslab + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + (lane & 3) * 4);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");     // every epilogue thread has read the x tile
      float dy[16];
      const bool all_rows = gene_ok && (r0 + 16 <= a.B);
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const float xh = fmaxf(__uint_as_float(acc[rr]) + bj, 0.f);
        float v = xh > 0.f ? a.gscale * (xh - xv[rr]) : 0.f;
        if (!all_rows && !(gene_ok && r0 + rr < a.B)) v = 0.f;
        dy[rr] = v;
        dbsum += v;
      }
      float4* drow = reinterpret_cast<float4*>(xs(s) + gl * 128);   // dY^T row of this gene (K-major, SWIZZLE_128B)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        drow[(4 * half + c) ^ (gl & 7)] = make_float4(dy[4 * c], dy[4 * c + 1], dy[4 * c + 2], dy[4 * c + 3]);
      fence_proxy_async();
      mbar_arrive(dy_ready + s);
    }
    if (half == 1) dbs[gl] = dbsum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (half == 0 && gene_ok)
      a.db_part[(int64_t)split * a.db_split_stride + (int64_t)arm * a.db_arm_stride + gene] = dbsum + dbs[gl];
    mbar_wait(g_full, 0);
    tc_fence_after();
    float* prow = a.part + (int64_t)split * a.part_split_stride + (int64_t)arm * a.part_arm_stride + (int64_t)gene * 128;
    for (int c0 = 64 * half; c0 < min(a.HN, 64 * half + 64); c0 += 16) {
      uint32_t rr[16];
      if (nt > 0) {
        tmem_ld16(tmem_g + lane_addr + (uint32_t)c0, rr);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) rr[e] = 0u;
      }
      if (gene_ok) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          reinterpret_cast<float4*>(prow + c0)[e] = make_float4(__uint_as_float(rr[4 * e]), __uint_as_float(rr[4 * e + 1]),
                                                                __uint_as_float(rr[4 * e + 2]), __uint_as_float(rr[4 * e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

int choose_gene_split(int ctas_mn, int ntiles, int max_split) {
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_split && s <= ntiles; ++s) {
    const int ctas = ctas_mn * s;
    const int waves = (ctas + 147) / 148;
    const double eff = (double)ctas / (waves * 148.0);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  return best;
}

__global__ void __launch_bounds__(256) partial_sum2_kernel(const float* part, int64_t split_stride, int64_t batch_stride,
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

}  // namespace

// x_hat / loss / dY / d h10 in one pass.  dY_out (optional) receives dY [A][B][D]; x_rec (optional) x_hat.
int tc_fc11_rows(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale,
                 int want_grad, float* dY_out, float* x_rec, double* recon_acc, cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  float* work = st.work;
  CUtensorMap tmH, tmWk, tmWm, tmX;
  int rc = make_map(&tmH, work + w.d[4], H, B, H, A, (int64_t)B * H, 128, false);
  if (rc) return rc;
  rc = make_map(&tmWk, st.params + L.offset[FC11_W], H, D, H, A, L.arm_stride, GN, false);
  if (rc) return rc;
  rc = make_map(&tmWm, st.params + L.offset[FC11_W], H, D, H, A, L.arm_stride, GN, true);
  if (rc) return rc;
  rc = make_map(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, 128, false);
  if (rc) return rc;
  RowsArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.D = D; a.H = H; a.HN = (H + 15) / 16 * 16;
  a.ntiles = (D + GN - 1) / GN;
  const int mt = (B + ROWS - 1) / ROWS;
  const int nsplit = choose_gene_split(mt * A, a.ntiles, want_grad ? w.fc1_splitk : 8);
  a.tiles_per_split = (a.ntiles + nsplit - 1) / nsplit;
  a.x_batched = in.x_arm_stride > 0;
  a.gscale = gscale; a.want_grad = want_grad;
  a.bias = st.params + L.offset[FC11_B]; a.bias_arm_stride = L.arm_stride;
  a.dY = dY_out; a.dy_arm_stride = (int64_t)B * D;
  a.x_rec = x_rec; a.xrec_arm_stride = (int64_t)B * D;
  a.part = work + w.fc1_part; a.part_arm_stride = (int64_t)w.Bpad * 128; a.part_split_stride = (int64_t)A * a.part_arm_stride;
  a.recon_acc = recon_acc;
  const size_t smem = H10_BYTES + (size_t)STAGES * STAGE_BYTES + 32 * 8 + 1024;
  static bool attr = false;
  if (!attr) {
    MVAE_CUDA(cudaFuncSetAttribute(fc11_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  fc11_fused_kernel<false><<<dim3(nsplit, mt, A), FUSED_THREADS, smem, s>>>(tmH, tmWk, tmWm, tmX, a);
  MVAE_LAUNCH_CHECK();
  if (want_grad) {
    partial_sum2_kernel<<<dim3((H + 31) / 32, (B + 7) / 8, A), 256, 0, s>>>(a.part, a.part_split_stride, a.part_arm_stride,
                                                                            128, nsplit, work + w.g_d10, (int64_t)B * H, H, B, H);
    MVAE_LAUNCH_CHECK();
  }
  return 0;
}


// d fc11.weight and d fc11.bias by the gene-owner pass (x_hat recomputed, nothing [B,D]-sized is stored)
int tc_fc11_genes(const mvae_dims& d, const mvae_state& st, const mvae_inputs& in, const Work& w, float gscale,
                  cudaStream_t s) {
  mvae_layout L;
  compute_layout(d, &L);
  const int A = d.n_arm, B = d.batch, D = d.input_dim, H = d.fc_dim;
  float* work = st.work;
  CUtensorMap tmR, tmTk, tmTm, tmX;
  // resident operand: W11 block [128 genes x H]; tiles: h10 [32 cells x H] in both images; x: [32 cells x 32 genes] boxes
  int rc = make_map(&tmR, st.params + L.offset[FC11_W], H, D, H, A, L.arm_stride, 128, false);
  if (rc) return rc;
  rc = make_map(&tmTk, work + w.d[4], H, B, H, A, (int64_t)B * H, GN, false);
  if (rc) return rc;
  rc = make_map(&tmTm, work + w.d[4], H, B, H, A, (int64_t)B * H, GN, true);
  if (rc) return rc;
  rc = make_map(&tmX, in.x, D, B, in.x_row_stride, A, in.x_arm_stride, GN, false);
  if (rc) return rc;
  RowsArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.D = D; a.H = H; a.HN = (H + 15) / 16 * 16;
  a.ntiles = (B + GN - 1) / GN;
  const int mt = (D + ROWS - 1) / ROWS;
  const int nsplit = choose_gene_split(mt * A, a.ntiles, 8);
  a.tiles_per_split = (a.ntiles + nsplit - 1) / nsplit;
  a.x_batched = in.x_arm_stride > 0;
  a.gscale = gscale; a.want_grad = 1;
  a.bias = st.params + L.offset[FC11_B]; a.bias_arm_stride = L.arm_stride;
  a.part = work + w.fc1_part; a.part_arm_stride = (int64_t)w.Dpad * 128; a.part_split_stride = (int64_t)A * a.part_arm_stride;
  a.db_part = work + w.db_part; a.db_arm_stride = w.Dpad; a.db_split_stride = (int64_t)A * w.Dpad;
  const size_t smem = H10_BYTES + (size_t)STAGES * STAGE_BYTES + 32 * 8 + 1024;
  static bool attr = false;
  if (!attr) {
    MVAE_CUDA(cudaFuncSetAttribute(fc11_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  fc11_fused_kernel<true><<<dim3(nsplit, mt, A), FUSED_THREADS, smem, s>>>(tmR, tmTk, tmTm, tmX, a);
  MVAE_LAUNCH_CHECK();
  partial_sum2_kernel<<<dim3((H + 31) / 32, (D + 7) / 8, A), 256, 0, s>>>(a.part, a.part_split_stride, a.part_arm_stride, 128,
                                                                          nsplit, st.grads + L.offset[FC11_W], L.arm_stride, H, D, H);
  MVAE_LAUNCH_CHECK();
  // d fc11.bias: [split][A][Dpad] -> grads (treated as a [1 x D] matrix per arm)
  partial_sum2_kernel<<<dim3((D + 31) / 32, 1, A), 256, 0, s>>>(a.db_part, a.db_split_stride, a.db_arm_stride, 0, nsplit,
                                                                st.grads + L.offset[FC11_B], L.arm_stride, 0, 1, D);
  MVAE_LAUNCH_CHECK();
  return 0;
}

}  // namespace mvae
