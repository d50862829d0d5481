// Read-only HBM bandwidth probe: how fast can a B200 stream bytes it never writes back?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/readbw scripts/micro/readbw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int U>
__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ p, size_t n, unsigned* sink) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}
__global__ void copy_kernel(const uint4* __restrict__ p, uint4* __restrict__ q, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) q[i] = p[i];
}
int main() {
  const size_t bytes = 4ull << 30;
  uint4 *p, *q; unsigned* sink;
  cudaMalloc(&p, bytes); cudaMalloc(&q, bytes); cudaMalloc(&sink, 4);
  cudaMemset(p, 1, bytes); cudaMemset(q, 2, bytes);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const size_t n = bytes / 16;
  for (int ctas_per_sm : {2, 4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) {
      float best = 1e9f;
      for (int it = 0; it < 6; ++it) {
        cudaEventRecord(a);
        read_kernel<8><<<148 * ctas_per_sm, 256>>>(p, n, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
      }
      if (rep) printf("read  U=8 ctas/sm=%2d: %.1f GB/s\n", ctas_per_sm, bytes / best / 1e6);
    }
  }
  {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(a);
      copy_kernel<<<148 * 16, 256>>>(p, q, n);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("copy (read+write bytes): %.1f GB/s\n", 2.0 * bytes / best / 1e6);
    best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(a);
      cudaMemcpyAsync(q, p, bytes, cudaMemcpyDeviceToDevice);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("cudaMemcpy D2D (read+write bytes): %.1f GB/s\n", 2.0 * bytes / best / 1e6);
  }
  // a 1.245 GB region (the lm_head's size) read repeatedly: what the lm_head GEMM could reach
  for (int it = 0; it < 3; ++it) {
    const size_t nb = 1244659712ull / 16;
    cudaEventRecord(a);
    read_kernel<8><<<148 * 8, 256>>>(p, nb, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("read 1.245 GB: %.1f us = %.1f GB/s\n", ms * 1e3, nb * 16 / ms / 1e6);
  }
  return 0;
}
