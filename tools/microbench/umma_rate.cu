// Microbenchmark: back-to-back tcgen05.mma dispatch/execution rate on sm_100a for the small shapes the attention
// kernels use (M=128, K=16 per instruction; N = 64/128/256; B K-major or MN-major; A from smem or from TMEM).
// One CTA per SM, one issuing thread; clk per MMA = (clock64 around `iters` x 8 MMAs + commit + wait) / count.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I recommendersystem_b200/csrc -o umma_rate tools/microbench/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace graft;

template <int N, int B_MN, int A_TMEM, int ROT>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_align1024(smem_raw);
  uint8_t* sA = base;            // 128 rows x 128 B (K-major, 64 bf16 of K): 16 KB
  uint8_t* sB = base + 16384;    // up to 256 rows x 128 B: 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (49152 / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, B_MN);
    const uint32_t a = smem_u32(sA), b = smem_u32(sB);
    uint32_t phase = 0;
    // warm-up
    umma_bf16(tmem, umma_desc_sw128(a, 0, 1024), umma_desc_sw128(b, 0, 1024), idesc, 0);
    umma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t d = tmem + (ROT ? ((i * 8 + k) & 1) * 256 : 0);   // ROT: alternate between two accumulators
        const uint64_t db = B_MN ? umma_desc_sw128(b + (k & 3) * 2048, 8192, 1024) : umma_desc_sw128(b + (k & 3) * 32, 0, 1024);
        if (A_TMEM) umma_bf16_ts(d, tmem + 480 + (k & 3) * 8, db, idesc, 1);
        else umma_bf16(d, umma_desc_sw128(a + (k & 3) * 32, 0, 1024), db, idesc, 1);
      }
    }
    const long long t1 = clock64();   // dispatch only
    umma_commit(&bar); mbar_wait(&bar, phase);
    const long long t2 = clock64();   // all retired
    cycles[blockIdx.x * 2] = t1 - t0;
    cycles[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int N, int B_MN, int A_TMEM, int ROT>
void run(const char* name, long long* d_cyc) {
  const int iters = 256, smem = 49152 + 1024;
  cudaFuncSetAttribute(umma_rate_kernel<N, B_MN, A_TMEM, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    umma_rate_kernel<N, B_MN, A_TMEM, ROT><<<grid, 128, smem>>>(iters, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s ERROR %s\n", name, cudaGetErrorString(e)); return; }
    long long h[2];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s grid %3d  dispatch %6.1f clk/MMA   retired %6.1f clk/MMA  (floor %d)\n", name, grid, h[0] / (iters * 8.0), h[1] / (iters * 8.0), N / 2);
  }
}

int main() {
  long long* d_cyc;
  cudaMalloc(&d_cyc, 148 * 2 * sizeof(long long));
  run<64, 0, 0, 0>("N=64  SS B K-major   one accumulator", d_cyc);
  run<64, 0, 0, 1>("N=64  SS B K-major   two accumulators", d_cyc);
  run<64, 1, 0, 0>("N=64  SS B MN-major  one accumulator", d_cyc);
  run<64, 1, 0, 1>("N=64  SS B MN-major  two accumulators", d_cyc);
  run<64, 0, 1, 0>("N=64  TS B K-major   one accumulator", d_cyc);
  run<64, 1, 1, 0>("N=64  TS B MN-major  one accumulator", d_cyc);
  run<64, 1, 1, 1>("N=64  TS B MN-major  two accumulators", d_cyc);
  run<128, 0, 0, 0>("N=128 SS B K-major   one accumulator", d_cyc);
  run<128, 0, 0, 1>("N=128 SS B K-major   two accumulators", d_cyc);
  run<128, 1, 0, 0>("N=128 SS B MN-major  one accumulator", d_cyc);
  run<128, 0, 1, 0>("N=128 TS B K-major   one accumulator", d_cyc);
  run<256, 0, 0, 0>("N=256 SS B K-major   one accumulator", d_cyc);
  run<256, 0, 0, 1>("N=256 SS B K-major   two accumulators", d_cyc);
  run<256, 0, 1, 0>("N=256 TS B K-major   one accumulator", d_cyc);
  return 0;
}
