// Microbenchmark: tcgen05.ld throughput per SM as a function of the number of reading warps (B200 TMEM read bandwidth).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int kShape>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cyc, uint32_t* sink, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      const uint32_t a = base + ((i * 32 + (warp >> 2) * 64) & 255);
      if (kShape == 32) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(a) : "memory");
      } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(a) : "memory");
#pragma unroll
        for (int j = 16; j < 32; ++j) v[j] = 0;
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc ^= v[j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}
int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 8 * 148); cudaMalloc(&sink, 4 * 148 * 512);
  const int iters = 4000;
  for (int shape : {32, 16})
    for (int nw : {1, 2, 4, 8, 16}) {
      if (shape == 32) k<32><<<148, 512>>>(iters, cyc, sink, nw); else k<16><<<148, 512>>>(iters, cyc, sink, nw);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * nw * 32 * shape * 4;
      printf("x%d  warps %2d: %lld cycles, %.1f B/clk/SM, %.1f cycles per ld per warp (%s)\n", shape, nw, h, bytes / h, double(h) / iters, cudaGetErrorString(e));
    }
  return 0;
}
