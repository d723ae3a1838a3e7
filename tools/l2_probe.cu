// L2 / HBM read-bandwidth probe: every CTA streams its share of a buffer with 16-byte ld.global.nc loads.
// A 32 MB buffer stays resident in the 126 MB L2 (L2 -> SM rate); a 1 GB buffer streams from HBM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_probe.bin tools/l2_probe.cu && tools/l2_probe.bin
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) read_kernel(const float4* __restrict__ p, size_t n4, int reps, float* sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 a, b, c, d;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p + i));
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + i + stride));
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(p + i + 2 * stride));
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(p + i + 3 * stride));
      acc += a.x + b.y + c.z + d.w;
    }
  }
  if (acc == 123.456f) *sink = acc;
}

int main() {
  const size_t big = (size_t)1 << 30;
  float4* buf; float* sink;
  cudaMalloc(&buf, big); cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, big);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t sizes[] = {(size_t)16 << 20, (size_t)32 << 20, (size_t)64 << 20, (size_t)96 << 20, (size_t)256 << 20, big};
  for (size_t sz : sizes) {
    const int reps = (int)((((size_t)8) << 30) / sz);
    for (int ctas_per_sm : {4, 8}) {
      read_kernel<<<148 * ctas_per_sm, 256>>>(buf, sz / 16, 2, sink);     // warm-up (fills L2 when it fits)
      cudaEventRecord(e0);
      read_kernel<<<148 * ctas_per_sm, 256>>>(buf, sz / 16, reps, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("{\"buffer_mb\": %zu, \"ctas_per_sm\": %d, \"read_gbs\": %.1f}\n", sz >> 20, ctas_per_sm, (double)sz * reps / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
