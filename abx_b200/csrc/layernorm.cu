// LayerNorm over the channel dimension of the pair / sequence activations (torch.nn.LayerNorm semantics,
// eps inside the square root; reference: every `LayerNorm(...)` of abx/model/seqformer.py and
// score_network.py:117-135).  One warp per row, the row held in registers (two-pass mean / variance),
// 16-byte loads and stores: the kernel is a pure HBM stream (read C floats, write C floats per row).
// The number of float4 per lane is a template parameter (KV = ceil(C / 128)): with the generic 8-deep predicated
// loops the C = 192 pair LayerNorm issued ~340 warp instructions per row and was issue-bound at 3.7 TB/s
// (ncu: issue slots 72 % busy, DRAM 44 %; profiles/r01_trunk_kernels_ncu.md).
//
// Optional output transposition of the two middle dimensions of a [B, n, n, C] tensor (row (b,i,j) is
// written to (b,j,i)), which replaces the `rearrange(pair_act, 'b i j c -> b j i c')` copy in front of the
// per-column triangle attention (seqformer.py:537-538).
#include "common.cuh"

namespace abx {

constexpr int kLnWarps = 8, kLnMaxV = 8;   // up to 8 float4 per lane: C <= 1024

template <int KV>
__global__ void __launch_bounds__(kLnWarps * 32) layernorm_kernel(
    long long rows, int C, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    float eps, int transpose_n, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = C >> 2;                        // float4 per row
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[KV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) {
      v[k] = xr[i];
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
  long long orow = row;
  if (transpose_n > 0) {
    if (rows <= 0xffffffffLL) {                  // 32-bit index arithmetic: 64-bit divisions cost ~100 instructions each
      const unsigned n = (unsigned)transpose_n, nn = n * n, r32 = (unsigned)row;
      const unsigned b = r32 / nn, r = r32 - b * nn, i = r / n, j = r - i * n;
      orow = (long long)b * nn + (long long)j * n + i;
    } else {
      const long long nn = (long long)transpose_n * transpose_n;
      const long long b = row / nn, r = row % nn;
      orow = b * nn + (r % transpose_n) * transpose_n + r / transpose_n;
    }
  }
  float4* yr = reinterpret_cast<float4*>(y + orow * C);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) {
      const float4 g = __ldg(g4 + i), b = __ldg(b4 + i);
      float4 o;
      o.x = (v[k].x - mean) * rstd * g.x + b.x;
      o.y = (v[k].y - mean) * rstd * g.y + b.y;
      o.z = (v[k].z - mean) * rstd * g.z + b.z;
      o.w = (v[k].w - mean) * rstd * g.w + b.w;
      yr[i] = o;
    }
  }
}

// LayerNorm over the channels of a CHANNEL-MAJOR tensor x [B, C, n, np] (np >= n: padded rows), written row-major
// y [B, n, n, C]: the 'b c i j -> b i j c' rearrange + final LayerNorm after the triangle-multiplication product
// (seqformer.py:500-502).  One CTA per (b, i, 32 values of j): coalesced 128-byte reads along j, transpose
// through shared memory, warp reductions over the C <= 128 channels, coalesced writes along c.
template <int C>
__global__ void __launch_bounds__(128) layernorm_cm_kernel(int n, int np, const float* __restrict__ x,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float eps, float* __restrict__ y) {
  __shared__ float t[C][33];
  const int j0 = blockIdx.x * 32, i = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    // compile-time trip count: all C / 4 loads of a thread are in flight before the first shared-memory store
    const int j = j0 + lane;
    const float* xp = x + (((size_t)b * C + warp) * n + i) * np + j;
    const size_t cstride = (size_t)4 * n * np;
    float v[C / 4];
#pragma unroll
    for (int k = 0; k < C / 4; ++k) v[k] = (j < n) ? __ldg(xp + k * cstride) : 0.f;
#pragma unroll
    for (int k = 0; k < C / 4; ++k) t[warp + 4 * k][lane] = v[k];
  }
  __syncthreads();
  constexpr int nc = C / 32;                     // channels per lane
  float g[nc], be[nc];
#pragma unroll
  for (int k = 0; k < nc; ++k) { g[k] = __ldg(gamma + lane + 32 * k); be[k] = __ldg(beta + lane + 32 * k); }
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int jl = warp * 8 + jj, j = j0 + jl;
    if (j >= n) break;
    float v[nc], sum = 0.f;
#pragma unroll
    for (int k = 0; k < nc; ++k) { v[k] = t[lane + 32 * k][jl]; sum += v[k]; }
    const float mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < nc; ++k) { const float d = v[k] - mean; sq += d * d; }
    const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
    float* yr = y + (((size_t)b * n + i) * n + j) * C;
#pragma unroll
    for (int k = 0; k < nc; ++k) yr[lane + 32 * k] = (v[k] - mean) * rstd * g[k] + be[k];
  }
}

}  // namespace abx

extern "C" int abx_layernorm_cm(void* stream, int B, int C, int n, int np, const float* x, const float* gamma,
                                const float* beta, float eps, float* y) {
  using namespace abx;
  ABX_REQUIRE(B > 0 && n > 0 && np >= n && x && gamma && beta && y, "abx_layernorm_cm: bad shape or null argument");
  ABX_REQUIRE(C % 32 == 0 && C <= 128, "abx_layernorm_cm: C must be a multiple of 32 and <= 128 (got %d)", C);
  ABX_REQUIRE(n <= 65535 && B <= 65535, "abx_layernorm_cm: n and B must be <= 65535");
  const dim3 grid((n + 31) / 32, n, B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 32: layernorm_cm_kernel<32><<<grid, 128, 0, st>>>(n, np, x, gamma, beta, eps, y); break;
    case 64: layernorm_cm_kernel<64><<<grid, 128, 0, st>>>(n, np, x, gamma, beta, eps, y); break;
    case 96: layernorm_cm_kernel<96><<<grid, 128, 0, st>>>(n, np, x, gamma, beta, eps, y); break;
    default: layernorm_cm_kernel<128><<<grid, 128, 0, st>>>(n, np, x, gamma, beta, eps, y); break;
  }
  count_launch();
  return check_launch("layernorm_cm_kernel");
}

extern "C" int abx_layernorm(void* stream, long long rows, int C, const float* x, const float* gamma, const float* beta,
                             float eps, int transpose_n, float* y) {
  using namespace abx;
  ABX_REQUIRE(rows > 0 && C > 0 && x && gamma && beta && y, "abx_layernorm: bad shape or null argument");
  ABX_REQUIRE(C % 4 == 0 && C <= 32 * 4 * kLnMaxV, "abx_layernorm: C must be a multiple of 4 and <= %d (got %d)", 32 * 4 * kLnMaxV, C);
  ABX_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0),
              "abx_layernorm: pointers must be 16-byte aligned");
  ABX_REQUIRE(transpose_n >= 0 && (transpose_n == 0 || rows % ((long long)transpose_n * transpose_n) == 0),
              "abx_layernorm: rows must be a multiple of transpose_n^2");
  ABX_REQUIRE(transpose_n == 0 || x != y, "abx_layernorm: the transposing form cannot run in place");
  const long long blocks = (rows + kLnWarps - 1) / kLnWarps;
  ABX_REQUIRE(blocks < 2147483647LL, "abx_layernorm: too many rows");
  const int kv = ((C >> 2) + 31) / 32;          // float4 per lane
  auto launch = [&](auto kern) { kern<<<(unsigned)blocks, kLnWarps * 32, 0, (cudaStream_t)stream>>>(rows, C, x, gamma, beta, eps, transpose_n, y); };
  switch (kv) {
    case 1: launch(layernorm_kernel<1>); break;
    case 2: launch(layernorm_kernel<2>); break;
    case 3: launch(layernorm_kernel<3>); break;
    case 4: launch(layernorm_kernel<4>); break;
    case 5: launch(layernorm_kernel<5>); break;
    case 6: launch(layernorm_kernel<6>); break;
    default: launch(layernorm_kernel<kLnMaxV>); break;
  }
  count_launch();
  return check_launch("layernorm_kernel");
}
