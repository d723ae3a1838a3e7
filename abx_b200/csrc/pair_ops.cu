// Small fused kernels for the pair-activation path of the trunk (reference abx/model/seqformer.py).
//
// pair_input_kernel — the pair input of EmbeddingAndSeqformer.forward (:193-222) in one pass:
//   pair_act[b,i,j,:] = concat(static[i,j,0:Cs], te[b], te[b]) + LayerNorm(prev_pair[b,i,j,:]) + emb[prev_pos[b,i,j], :]
// (static = step-invariant pair embedding of the complex, te = timestep embedding, emb = proj_prev_pos table).
// The reference spends cat + LayerNorm + add + embedding gather + add = ~9 passes over [B,N,N,192]; this is one
// read of prev_pair and one write.  Warp per row, row in registers, 16-byte accesses.
//
// outer_product_kernel — OuterProductMean features (:392-407):
//   out[b,i,j,0:C] = left[b,j,:] * right[b,i,:],  out[b,i,j,C:2C] = left[b,j,:] - right[b,i,:]
// written directly in the layout the output projection GEMM reads (no prod / diff / cat temporaries).
#include "common.cuh"

namespace abx {

constexpr int kPiWarps = 8, kPiMaxV = 2;    // up to 2 float4 per lane: C <= 256

__global__ void __launch_bounds__(kPiWarps * 32) pair_input_kernel(
    long long rows, int nn, int C, int Cs, int Ct, const float* __restrict__ stat, const float* __restrict__ te,
    const float* __restrict__ prev, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    const long long* __restrict__ prev_pos, const float* __restrict__ emb, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kPiWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = C >> 2;
  // 32-bit index arithmetic (the host checks rows < 2^31): a 64-bit division costs ~100 instructions, and this kernel is a
  // pure stream (one row = 768 bytes in, 768 out)
  const unsigned r32 = (unsigned)row, b = r32 / (unsigned)nn, ij = r32 - b * (unsigned)nn;
  float4 v[kPiMaxV];
  float sum = 0.f;
  const float4* pr = prev ? reinterpret_cast<const float4*>(prev + row * C) : nullptr;
#pragma unroll
  for (int k = 0; k < kPiMaxV; ++k) {
    const int i4 = lane + 32 * k;
    if (i4 < nv && pr) { v[k] = pr[i4]; sum += (v[k].x + v[k].y) + (v[k].z + v[k].w); }
  }
  float mean = 0.f, rstd = 0.f;
  if (pr) {
    mean = warp_sum(sum) / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < kPiMaxV; ++k) {
      const int i4 = lane + 32 * k;
      if (i4 < nv) { const float a = v[k].x - mean, b2 = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean; sq += (a * a + b2 * b2) + (c * c + d * d); }
    }
    rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
  }
  const float4* er = prev_pos ? reinterpret_cast<const float4*>(emb + __ldg(prev_pos + row) * C) : nullptr;
  float4* yr = reinterpret_cast<float4*>(y + row * C);
#pragma unroll
  for (int k = 0; k < kPiMaxV; ++k) {
    const int i4 = lane + 32 * k;
    if (i4 >= nv) continue;
    const int c0 = 4 * i4;
    float4 o;                                         // concat(static, te, te): channel blocks never straddle a float4
    if (c0 < Cs) o = __ldg(reinterpret_cast<const float4*>(stat + (size_t)ij * Cs + c0));
    else { const int t0 = c0 - Cs; o = __ldg(reinterpret_cast<const float4*>(te + (size_t)b * Ct + (t0 >= Ct ? t0 - Ct : t0))); }
    if (pr) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i4), be = __ldg(reinterpret_cast<const float4*>(beta) + i4);
      // (static + LN(prev)) + emb, in the reference's order of additions
      o.x += (v[k].x - mean) * rstd * g.x + be.x; o.y += (v[k].y - mean) * rstd * g.y + be.y;
      o.z += (v[k].z - mean) * rstd * g.z + be.z; o.w += (v[k].w - mean) * rstd * g.w + be.w;
    }
    if (er) { const float4 e = __ldg(er + i4); o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
    yr[i4] = o;
  }
}

// grid (ceil(N q / 256), N, B) with q = 2C/4 float4 per output row: thread = one float4 of row (b, i, j) — all index arithmetic
// in 32 bits (the flat 64-bit div / mod chain of the first version made this write-only stream issue-bound at 1.3 TB/s)
__global__ void __launch_bounds__(256) outer_product_kernel(int N, int C, const float* __restrict__ left,
                                                            const float* __restrict__ right, float* __restrict__ out) {
  const unsigned q = (unsigned)(2 * C) >> 2;
  const unsigned t = blockIdx.x * 256u + threadIdx.x;          // (j, f) flattened
  if (t >= (unsigned)N * q) return;
  const unsigned j = t / q, f = t - j * q;
  const unsigned i = blockIdx.y, b = blockIdx.z;
  const unsigned c0 = 4 * f >= (unsigned)C ? 4 * f - C : 4 * f;
  const float4 l = __ldg(reinterpret_cast<const float4*>(left + ((size_t)b * N + j) * C + c0));
  const float4 r = __ldg(reinterpret_cast<const float4*>(right + ((size_t)b * N + i) * C + c0));
  float4 o;
  if (4 * f < (unsigned)C) o = make_float4(l.x * r.x, l.y * r.y, l.z * r.z, l.w * r.w);
  else o = make_float4(l.x - r.x, l.y - r.y, l.z - r.z, l.w - r.w);
  reinterpret_cast<float4*>(out)[((size_t)b * N + i) * N * q + t] = o;
}

}  // namespace abx

extern "C" int abx_pair_input(void* stream, int B, int N, int C, int Cs, int Ct, const float* stat, const float* te,
                              const float* prev_pair, const float* gamma, const float* beta, float eps,
                              const int64_t* prev_pos, const float* emb, float* y) {
  using namespace abx;
  ABX_REQUIRE(B > 0 && N > 0 && stat && te && y, "abx_pair_input: bad shape or null argument");
  ABX_REQUIRE(C % 4 == 0 && C <= 128 * kPiMaxV && Cs % 4 == 0 && Ct % 4 == 0 && Cs + 2 * Ct == C,
              "abx_pair_input: need C = Cs + 2 Ct <= 256 with all widths multiples of 4 (C=%d Cs=%d Ct=%d)", C, Cs, Ct);
  ABX_REQUIRE(!prev_pair || (gamma && beta), "abx_pair_input: prev_pair needs LayerNorm parameters");
  ABX_REQUIRE(!prev_pos || emb, "abx_pair_input: prev_pos needs the embedding table");
  const long long rows = (long long)B * N * N;
  const long long blocks = (rows + kPiWarps - 1) / kPiWarps;
  ABX_REQUIRE(rows < 2147483647LL, "abx_pair_input: too many rows");
  pair_input_kernel<<<(unsigned)blocks, kPiWarps * 32, 0, (cudaStream_t)stream>>>(
      rows, N * N, C, Cs, Ct, stat, te, prev_pair, gamma, beta, eps, reinterpret_cast<const long long*>(prev_pos), emb, y);
  count_launch();
  return check_launch("pair_input_kernel");
}

extern "C" int abx_outer_product(void* stream, int B, int N, int C, const float* left, const float* right, float* out) {
  using namespace abx;
  ABX_REQUIRE(B > 0 && N > 0 && C > 0 && C % 4 == 0 && left && right && out, "abx_outer_product: bad shape or null argument");
  ABX_REQUIRE(N <= 65535 && B <= 65535 && (long long)N * (2 * C / 4) < 2147483647LL, "abx_outer_product: N, B must be <= 65535");
  outer_product_kernel<<<dim3((unsigned)((N * (2 * C / 4) + 255) / 256), N, B), 256, 0, (cudaStream_t)stream>>>(N, C, left, right, out);
  count_launch();
  return check_launch("outer_product_kernel");
}
