// Micro-benchmark: tcgen05.mma dispatch rate on sm_100a for the shapes the GAT kernels use.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I magat_pathplanning_b200/csrc tools/mma_rate.cu -o /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace magat;

template <int TS, int N>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc<512>(&slot);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N);
    const uint32_t sa = tc::smem_u32(smem);
    const uint64_t a = tc::make_sw128_desc(sa), b = tc::make_sw128_desc(sa + 32768);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (TS) tc::umma_bf16_ts(tm + 256, tm + kk * 8, b + kk * 2, idesc, 1);
        else tc::umma_bf16(tm + 256, a + kk * 2, b + kk * 2, idesc, 1);
      }
    }
    t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc::tc_fence_after(); tc::tmem_dealloc<512>(tm); }
}

template <int TS, int N>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 2000;
  cudaFuncSetAttribute(k_rate<TS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_rate<TS, N><<<grid, 128, 100 * 1024>>>(iters, d);
  k_rate<TS, N><<<grid, 128, 100 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s grid %3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", name, grid, (double)h[0] / (iters * 4),
         (double)h[1] / (iters * 4), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<0, 64>("SS M128 N64", grid);
    run<0, 128>("SS M128 N128", grid);
    run<0, 256>("SS M128 N256", grid);
    run<1, 64>("TS M128 N64", grid);
    run<1, 128>("TS M128 N128", grid);
    run<1, 256>("TS M128 N256", grid);
  }
  return 0;
}
