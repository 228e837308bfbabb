// micro-benchmark: tcgen05.mma kind::tf32 issue/latency behaviour (dependent chain vs independent accumulators)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../distributed-vae_b200/csrc/tc_common.cuh"
using namespace mvae::tc;

// mode: 0 SS same acc, 1 SS alternating nacc accs, 2 TS same acc, 3 TS alternating
__global__ void __launch_bounds__(128, 1) k(int N, int nmma, int nacc, int ts, int commit_every, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(dummy + i, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc(128, N, false, false);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < nmma; ++i) {
          const uint32_t d = tb + (uint32_t)(i % nacc) * 128u * 0 + (uint32_t)(i % nacc) * (uint32_t)(N <= 128 ? 128 : 256);
          const uint64_t bd = make_smem_desc(sb + (i & 3) * 32, 0, 1024, false);
          if (ts) umma_tf32_ts(d, tb + 480u + (i & 3) * 8, bd, idesc, 1u);
          else umma_tf32(d, make_smem_desc(sa + (i & 3) * 32, 0, 1024, false), bd, idesc, 1u);
          if (commit_every && (i % commit_every) == commit_every - 1) umma_commit(dummy + ((i / commit_every) & 3));
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  const int Ns[] = {32, 64, 112, 128};
  for (int ts = 1; ts < 2; ++ts)
    for (int N : Ns)
      for (int nacc : {0, 1, 2, 4, 13}) {
        const int commit_every = nacc; nacc = 1;
        const int nmma = 260;
        k<<<148, 128, 40960>>>(N, nmma, nacc, ts, commit_every, d);
        long long h = 0; cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%s N=%3d commit_every=%d : %6.1f cycles/MMA (math floor %d) %s\n", ts ? "TS" : "SS", N, commit_every, (double)h / nmma, N / 2,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
