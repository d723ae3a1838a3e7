// Fused gated-attention core with pair bias for the trunk's triangle attention (reference:
// abx/model/seqformer.py:283-301 `Attention.forward` as called by `TriangleAttention` :506-550):
//
//   o[b,s,i,h,:] = softmax_j( q[b,s,i,h,:] . k[b,s,j,h,:] / sqrt(D) + bias[b,h,i,j], keys with mask 0 -> finfo.min ) v[b,s,j,h,:]
//
// The reference materialises logits and weights of shape [B,S,H,L,L] (2.7 GB at B=4, L=350); here they
// never leave the SM.  One CTA per (b,s,h) and 384 query rows: the L x D key and value slices are staged once in shared
// memory; lane = query row (q row and the output accumulator live in registers, no cross-lane reduction);
// keys are processed 32 at a time with an online softmax (exact row max / sum up to rounding); the bias tile
// of the warp's 32 rows x 32 keys is loaded coalesced and transposed through a per-warp shared-memory tile.
#include <float.h>

#include "common.cuh"

namespace abx {

constexpr int kAttChunk = 32, kAttMaxThreads = 384;   // 384 threads -> up to 170 registers each

template <int D>
__global__ void __launch_bounds__(kAttMaxThreads, 1) pair_attention_kernel(
    int L, int H, int S, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int ld,
    const float* __restrict__ bias, const float* __restrict__ key_mask, float scale, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                                  // [L][D]
  float* Vs = Ks + (size_t)L * D;                  // [L][D]
  float* Ms = Vs + (size_t)L * D;                  // [Lpad] additive mask flags (1 keep / 0 drop)
  const int Lpad = (L + 31) & ~31;
  float* Bt = Ms + Lpad;                           // per warp [32][33] bias tile
  const int h = blockIdx.x, bs = blockIdx.y, b = bs / S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t row0 = (size_t)bs * L;

  // stage K, V (rows of D floats, 16-byte copies) and the key mask
  constexpr int D4 = D / 4;
  for (int idx = threadIdx.x; idx < L * D4; idx += blockDim.x) {
    const int j = idx / D4, d4 = idx % D4;
    const size_t g = (row0 + j) * (size_t)ld + h * D + 4 * d4;
    reinterpret_cast<float4*>(Ks)[idx] = *reinterpret_cast<const float4*>(k + g);
    reinterpret_cast<float4*>(Vs)[idx] = *reinterpret_cast<const float4*>(v + g);
  }
  for (int j = threadIdx.x; j < Lpad; j += blockDim.x)
    Ms[j] = (j < L) ? (key_mask ? __ldg(key_mask + (size_t)b * L + j) : 1.f) : 0.f;
  __syncthreads();

  const int wrow0 = blockIdx.z * blockDim.x + warp * 32;   // first query row of this warp
  const int i = wrow0 + lane;                      // query row of this lane
  const bool row_ok = i < L;
  const int ic = row_ok ? i : L - 1;
  // packed fp32 pairs: the inner products and the value accumulation run on FFMA2 (fma.rn.f32x2, sm_100),
  // two FMAs per issued instruction
  float2 qr[D / 2], acc[D / 2];
  {
    const float4* qp = reinterpret_cast<const float4*>(q + (row0 + ic) * (size_t)ld + h * D);
#pragma unroll
    for (int d4 = 0; d4 < D4; ++d4) {
      const float4 t = __ldg(qp + d4);
      qr[2 * d4] = make_float2(t.x * scale, t.y * scale);
      qr[2 * d4 + 1] = make_float2(t.z * scale, t.w * scale);
    }
  }
#pragma unroll
  for (int d = 0; d < D / 2; ++d) acc[d] = make_float2(0.f, 0.f);
  float m = -FLT_MAX, l = 0.f;
  float* bt = Bt + warp * (32 * 33);
  const float* bias_bh = bias + ((size_t)b * H + h) * L * L;

  for (int j0 = 0; j0 < L; j0 += kAttChunk) {
    const int nj = min(kAttChunk, L - j0);
    // bias tile: coalesced along j, transposed so that lane = row can read its own 32 values
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const int ir = wrow0 + r;
      bt[r * 33 + lane] = (ir < L && lane < nj) ? __ldg(bias_bh + (size_t)ir * L + j0 + lane) : 0.f;
    }
    __syncwarp();
    float s[kAttChunk];
    float cmax = -FLT_MAX;
#pragma unroll
    for (int jj = 0; jj < kAttChunk; ++jj) {
      if (jj < nj) {
        const float4* kr = reinterpret_cast<const float4*>(Ks + (size_t)(j0 + jj) * D);
        float2 da = make_float2(0.f, 0.f), db = make_float2(0.f, 0.f);
#pragma unroll
        for (int d4 = 0; d4 < D4; ++d4) {
          const float4 kk = kr[d4];
          da = __ffma2_rn(qr[2 * d4], make_float2(kk.x, kk.y), da);
          db = __ffma2_rn(qr[2 * d4 + 1], make_float2(kk.z, kk.w), db);
        }
        float sv = ((da.x + da.y) + (db.x + db.y)) + bt[lane * 33 + jj];
        sv = (Ms[j0 + jj] != 0.f) ? sv : -FLT_MAX;                    // masked_fill(~k_mask, finfo.min)
        s[jj] = sv;
        cmax = fmaxf(cmax, sv);
      } else {
        s[jj] = -FLT_MAX;
      }
    }
    const float m_new = fmaxf(m, cmax);
    const float corr = expf(m - m_new);
    l *= corr;
#pragma unroll
    for (int d = 0; d < D / 2; ++d) { acc[d].x *= corr; acc[d].y *= corr; }
    m = m_new;
#pragma unroll
    for (int jj = 0; jj < kAttChunk; ++jj) {
      if (jj < nj) {
        const float p = expf(s[jj] - m);
        l += p;
        const float4* vr = reinterpret_cast<const float4*>(Vs + (size_t)(j0 + jj) * D);
#pragma unroll
        const float2 pp = make_float2(p, p);
#pragma unroll
        for (int d4 = 0; d4 < D4; ++d4) {
          const float4 vv = vr[d4];
          acc[2 * d4] = __ffma2_rn(pp, make_float2(vv.x, vv.y), acc[2 * d4]);
          acc[2 * d4 + 1] = __ffma2_rn(pp, make_float2(vv.z, vv.w), acc[2 * d4 + 1]);
        }
      }
    }
  }
  if (row_ok) {
    const float inv = 1.f / l;
    float4* op = reinterpret_cast<float4*>(out + (row0 + i) * (size_t)(H * D) + h * D);
#pragma unroll
    for (int d4 = 0; d4 < D4; ++d4)
      op[d4] = make_float4(acc[2 * d4].x * inv, acc[2 * d4].y * inv, acc[2 * d4 + 1].x * inv, acc[2 * d4 + 1].y * inv);
  }
}

static size_t attention_smem_bytes(int L, int D, int threads) {
  const int Lpad = (L + 31) & ~31;
  return ((size_t)2 * L * D + Lpad + (size_t)(threads / 32) * 32 * 33) * sizeof(float);
}

template <int D>
static int launch_attention(cudaStream_t st, int B, int S, int L, int H, const float* q, const float* k, const float* v, int ld,
                            const float* bias, const float* key_mask, float* out) {
  const int threads = min(((L + 31) / 32) * 32, kAttMaxThreads);
  const size_t smem = attention_smem_bytes(L, D, threads);
  ABX_REQUIRE(smem <= 227 * 1024, "abx_pair_attention: L=%d with head dim %d needs %zu bytes of shared memory (max 232448)", L, D, smem);
  ABX_CUDA(cudaFuncSetAttribute(pair_attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_attention_kernel<D><<<dim3(H, B * S, (L + threads - 1) / threads), threads, smem, st>>>(L, H, S, q, k, v, ld, bias, key_mask, 1.0f / sqrtf((float)D), out);
  count_launch();
  return check_launch("pair_attention_kernel");
}

}  // namespace abx

extern "C" int abx_pair_attention(void* stream, int B, int S, int L, int H, int D, const float* q, const float* k,
                                  const float* v, int ld, const float* bias, const float* key_mask, float* out) {
  using namespace abx;
  ABX_REQUIRE(B > 0 && S > 0 && L > 0 && H > 0 && q && k && v && bias && out, "abx_pair_attention: bad shape or null argument");
  ABX_REQUIRE(ld % 4 == 0 && ld >= H * D, "abx_pair_attention: ld must be a multiple of 4 and >= H*D");
  ABX_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0),
              "abx_pair_attention: q, k, v, out must be 16-byte aligned");
  ABX_REQUIRE((long long)B * S <= 65535, "abx_pair_attention: B*S exceeds 65535");
  cudaStream_t st = (cudaStream_t)stream;
  switch (D) {
    case 16: return launch_attention<16>(st, B, S, L, H, q, k, v, ld, bias, key_mask, out);
    case 32: return launch_attention<32>(st, B, S, L, H, q, k, v, ld, bias, key_mask, out);
    case 48: return launch_attention<48>(st, B, S, L, H, q, k, v, ld, bias, key_mask, out);
    case 64: return launch_attention<64>(st, B, S, L, H, q, k, v, ld, bias, key_mask, out);
  }
  set_error("abx_pair_attention: head dim %d not instantiated (16, 32, 48, 64)", D);
  return ABX_ERR_INVALID;
}
