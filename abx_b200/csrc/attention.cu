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

// =====================================================================================================
// Tensor-core version of the same attention core: S = Q K^T and O = P V on mma.sync.m16n8k8 TF32 with the
// 3xTF32 operand split (hi = x & ~0x1fff, lo = x - hi; hi*lo + lo*hi + hi*hi), i.e. fp32-level accuracy.
// tcgen05 is not used here on purpose: the tiles are 16 x 8 x 8 fragments of a 350 x 350 x 48 problem per
// (b,s,h) and live entirely in registers between the two products (FlashAttention-2 dataflow), which the
// TMEM/descriptor model of tcgen05 does not fit without staging P through shared memory.
//
// One CTA per (b,s,h) and a group of up to 12 query tiles of 16 rows (one warp each).  K and V of the
// (b,s,h) slice are staged once in shared memory with row stride D+4 floats, which makes every fragment
// load of a warp hit 32 distinct banks.  Per 32-key chunk a warp computes four 16x8 score tiles (chained
// in the tensor core over the D/8 k-steps), adds the pair bias, applies the key mask, updates the online
// softmax per row (quad shuffles), and multiplies the probabilities with V; the accumulator layout of S
// doubles as the A-operand layout of P V after relabelling the keys inside each group of 8
// (k-index t <-> key 2t, k-index t+4 <-> key 2t+1), so no shuffles are needed between the two products.
// =====================================================================================================
namespace abx {

constexpr int kMmaMaxWarps = 12;

template <int D>
__global__ void __launch_bounds__(kMmaMaxWarps * 32, 1) pair_attention_mma_kernel(
    int L, int H, int S, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int ld,
    const float* __restrict__ bias, const float* __restrict__ key_mask, const float* __restrict__ gate, float scale,
    float* __restrict__ out) {
  constexpr int KS = D + 4, D4 = D / 4, NK = D / 8;   // smem row stride, float4 per row, k-steps of Q K^T (= n-tiles of P V)
  extern __shared__ __align__(16) float sm[];
  const int Lpad = (L + 31) & ~31;
  float* Ks = sm;                                  // [Lpad][KS]
  float* Vs = Ks + (size_t)Lpad * KS;              // [Lpad][KS]
  float* Ms = Vs + (size_t)Lpad * KS;              // [Lpad] 1 keep / 0 masked
  const int h = blockIdx.x, bs = blockIdx.y, b = bs / S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const size_t row0 = (size_t)bs * L;

  for (int idx = threadIdx.x; idx < Lpad * D4; idx += blockDim.x) {
    const int j = idx / D4, d4 = idx % D4;
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
    if (j < L) {
      const size_t gi = (row0 + j) * (size_t)ld + h * D + 4 * d4;
      kk = *reinterpret_cast<const float4*>(k + gi);
      vv = *reinterpret_cast<const float4*>(v + gi);
    }
    *reinterpret_cast<float4*>(Ks + (size_t)j * KS + 4 * d4) = kk;
    *reinterpret_cast<float4*>(Vs + (size_t)j * KS + 4 * d4) = vv;
  }
  for (int j = threadIdx.x; j < Lpad; j += blockDim.x)
    Ms[j] = (j < L) ? (key_mask ? __ldg(key_mask + (size_t)b * L + j) : 1.f) : 0.f;
  __syncthreads();

  const int r0 = (blockIdx.z * (blockDim.x >> 5) + warp) * 16;   // first query row of this warp
  if (r0 >= L) return;
  const int i0 = min(r0 + g, L - 1), i1 = min(r0 + g + 8, L - 1);

  // Q fragments (scaled), split once
  uint32_t qhi[NK][4], qlo[NK][4];
  {
    const float* q0 = q + (row0 + i0) * (size_t)ld + h * D;
    const float* q1 = q + (row0 + i1) * (size_t)ld + h * D;
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
      split_tf32(__ldg(q0 + 8 * kk + t) * scale, qhi[kk][0], qlo[kk][0]);
      split_tf32(__ldg(q1 + 8 * kk + t) * scale, qhi[kk][1], qlo[kk][1]);
      split_tf32(__ldg(q0 + 8 * kk + t + 4) * scale, qhi[kk][2], qlo[kk][2]);
      split_tf32(__ldg(q1 + 8 * kk + t + 4) * scale, qhi[kk][3], qlo[kk][3]);
    }
  }
  float oacc[NK][4];
#pragma unroll
  for (int m = 0; m < NK; ++m) oacc[m][0] = oacc[m][1] = oacc[m][2] = oacc[m][3] = 0.f;
  float m0 = -FLT_MAX, m1 = -FLT_MAX, l0 = 0.f, l1 = 0.f;      // running max / partial sum of rows g and g+8
  const float* bias0 = bias + (((size_t)b * H + h) * L + i0) * L;
  const float* bias1 = bias + (((size_t)b * H + h) * L + i1) * L;

  for (int j0 = 0; j0 < Lpad; j0 += 32) {
    // ---- S = Q K^T for 32 keys: four 16x8 tiles, chained over the D/8 k-steps in the tensor core
    float s[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    // k-step outer, score tile inner: consecutive MMAs go to different accumulators (4 independent chains)
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
      uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float* kr = Ks + (size_t)(j0 + 8 * n + g) * KS + t + 8 * kk;
        split_tf32(kr[0], bh0[n], bl0[n]);
        split_tf32(kr[4], bh1[n], bl1[n]);
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) mma_tf32(s[n], qhi[kk], bl0[n], bl1[n]);
#pragma unroll
      for (int n = 0; n < 4; ++n) mma_tf32(s[n], qlo[kk], bh0[n], bh1[n]);
#pragma unroll
      for (int n = 0; n < 4; ++n) mma_tf32(s[n], qhi[kk], bh0[n], bh1[n]);
    }
    // ---- bias, key mask, tail keys; row maxima
    float cm0 = -FLT_MAX, cm1 = -FLT_MAX;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = j0 + 8 * n + 2 * t + e;
        const bool valid = j < L;
        const bool keep = Ms[j] != 0.f;
        const float b0v = valid ? __ldg(bias0 + j) : 0.f, b1v = valid ? __ldg(bias1 + j) : 0.f;
        float a0 = s[n][e] + b0v, a1 = s[n][2 + e] + b1v;
        a0 = valid ? (keep ? a0 : -FLT_MAX) : -INFINITY;       // masked_fill(finfo.min); padding keys contribute exactly 0
        a1 = valid ? (keep ? a1 : -FLT_MAX) : -INFINITY;
        s[n][e] = a0; s[n][2 + e] = a1;
        cm0 = fmaxf(cm0, a0); cm1 = fmaxf(cm1, a1);
      }
    }
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
    const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
    const float c0 = expf(m0 - mn0), c1 = expf(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int m = 0; m < NK; ++m) { oacc[m][0] *= c0; oacc[m][1] *= c0; oacc[m][2] *= c1; oacc[m][3] *= c1; }
    // ---- P = exp(S - m); O += P V (per chunk chained in the tensor core, then added in fp32)
    float pacc[NK][4];
#pragma unroll
    for (int m = 0; m < NK; ++m) pacc[m][0] = pacc[m][1] = pacc[m][2] = pacc[m][3] = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float p00 = expf(s[n][0] - m0), p01 = expf(s[n][1] - m0), p10 = expf(s[n][2] - m1), p11 = expf(s[n][3] - m1);
      l0 += p00 + p01; l1 += p10 + p11;
      // A fragment of P for k-step n: (row g, k t) = key 2t, (row g+8, k t), (row g, k t+4) = key 2t+1, (row g+8, k t+4)
      uint32_t phi[4], plo[4];
      split_tf32(p00, phi[0], plo[0]); split_tf32(p10, phi[1], plo[1]);
      split_tf32(p01, phi[2], plo[2]); split_tf32(p11, phi[3], plo[3]);
      const float* vr = Vs + (size_t)(j0 + 8 * n + 2 * t) * KS + g;
      uint32_t vh0[NK], vl0[NK], vh1[NK], vl1[NK];
#pragma unroll
      for (int m = 0; m < NK; ++m) {
        split_tf32(vr[8 * m], vh0[m], vl0[m]);             // V[key 2t][8m+g]
        split_tf32(vr[KS + 8 * m], vh1[m], vl1[m]);        // V[key 2t+1][8m+g]
      }
#pragma unroll
      for (int m = 0; m < NK; ++m) mma_tf32(pacc[m], phi, vl0[m], vl1[m]);
#pragma unroll
      for (int m = 0; m < NK; ++m) mma_tf32(pacc[m], plo, vh0[m], vh1[m]);
#pragma unroll
      for (int m = 0; m < NK; ++m) mma_tf32(pacc[m], phi, vh0[m], vh1[m]);
    }
#pragma unroll
    for (int m = 0; m < NK; ++m) { oacc[m][0] += pacc[m][0]; oacc[m][1] += pacc[m][1]; oacc[m][2] += pacc[m][2]; oacc[m][3] += pacc[m][3]; }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  // optional output gating (seqformer.py:296-299): out = sigmoid(gate) * weighted average, gate laid out like q
  const size_t HD = (size_t)H * D;
  auto store_row = [&](int i, float a0, float a1, int m) {
    float2 o2 = make_float2(a0, a1);
    if (gate) {
      const float2 gv = *reinterpret_cast<const float2*>(gate + (row0 + i) * (size_t)ld + h * D + 8 * m + 2 * t);
      o2.x *= 1.f / (1.f + expf(-gv.x));
      o2.y *= 1.f / (1.f + expf(-gv.y));
    }
    *reinterpret_cast<float2*>(out + (row0 + i) * HD + h * D + 8 * m + 2 * t) = o2;
  };
  if (r0 + g < L) {
#pragma unroll
    for (int m = 0; m < NK; ++m) store_row(r0 + g, oacc[m][0] * inv0, oacc[m][1] * inv0, m);
  }
  if (r0 + g + 8 < L) {
#pragma unroll
    for (int m = 0; m < NK; ++m) store_row(r0 + g + 8, oacc[m][2] * inv1, oacc[m][3] * inv1, m);
  }
}

template <int D>
static int launch_attention_mma(cudaStream_t st, int B, int S, int L, int H, const float* q, const float* k, const float* v,
                                int ld, const float* bias, const float* key_mask, const float* gate, float* out) {
  const int Lpad = (L + 31) & ~31;
  const size_t smem = ((size_t)2 * Lpad * (D + 4) + Lpad) * sizeof(float);
  ABX_REQUIRE(smem <= 227 * 1024, "abx_pair_attention: L=%d with head dim %d needs %zu bytes of shared memory (max 232448)", L, D, smem);
  ABX_CUDA(cudaFuncSetAttribute(pair_attention_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (L + 15) / 16, nz = (tiles + kMmaMaxWarps - 1) / kMmaMaxWarps, warps = (tiles + nz - 1) / nz;
  pair_attention_mma_kernel<D><<<dim3(H, B * S, nz), warps * 32, smem, st>>>(L, H, S, q, k, v, ld, bias, key_mask, gate,
                                                                               1.0f / sqrtf((float)D), out);
  count_launch();
  return check_launch("pair_attention_mma_kernel");
}

}  // namespace abx

namespace abx {
template <int D>
int launch_attention_tc5(cudaStream_t st, int B, int S, int L, int H, const float* q, const float* k, const float* v, int ld,
                         const float* bias_tiles, const float* gate, float* out);
template <int D> size_t attention_tc5_smem(int L);   // shared memory of the tcgen05 kernel (attention_tc5.cu) for key length L
static bool tc5_fits(int L, int D) {
  switch (D) {
    case 16: return attention_tc5_smem<16>(L) <= 227 * 1024;
    case 32: return attention_tc5_smem<32>(L) <= 227 * 1024;
    case 48: return attention_tc5_smem<48>(L) <= 227 * 1024;
  }
  return false;                                      // D = 64: the operand tiles of two query slots do not fit
}
}  // namespace abx

namespace abx { int attention_tc5_profile(unsigned long long* out32); }
// Diagnostics: the 32 per-role cycle counters CTA (0,0,0) of the last tcgen05 attention launch recorded (library built with
// -DABX_ATTN_PROFILE=1; layout in attention_tc5.cu).
extern "C" int abx_attention_profile(unsigned long long* out16) {
  ABX_REQUIRE(out16 != nullptr, "abx_attention_profile: null output");
  return abx::attention_tc5_profile(out16);
}

// impl: 0 or 2 = mma.sync kernel, 1 = SIMT kernel (the tcgen05 kernel takes the bias transposed: abx_pair_attention_tc5)
extern "C" int abx_pair_attention_impl(void* stream, int impl, int B, int S, int L, int H, int D, const float* q, const float* k,
                                       const float* v, int ld, const float* bias, const float* key_mask, const float* gate,
                                       float* out) {
  using namespace abx;
  if (impl == 1) {
    ABX_REQUIRE(gate == nullptr, "abx_pair_attention_impl: output gating is only fused in the tensor-core kernel (impl 0)");
    return abx_pair_attention(stream, B, S, L, H, D, q, k, v, ld, bias, key_mask, out);
  }
  ABX_REQUIRE(gate == nullptr || (uintptr_t)gate % 8 == 0, "abx_pair_attention_impl: gate must be 8-byte aligned");
  ABX_REQUIRE(B > 0 && S > 0 && L > 0 && H > 0 && q && k && v && bias && out, "abx_pair_attention: bad shape or null argument");
  ABX_REQUIRE(ld % 4 == 0 && ld >= H * D, "abx_pair_attention: ld must be a multiple of 4 and >= H*D");
  ABX_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0),
              "abx_pair_attention: q, k, v, out must be 16-byte aligned");
  ABX_REQUIRE((long long)B * S <= 65535, "abx_pair_attention: B*S exceeds 65535");
  cudaStream_t st = (cudaStream_t)stream;
  switch (D) {
    case 16: return launch_attention_mma<16>(st, B, S, L, H, q, k, v, ld, bias, key_mask, gate, out);
    case 32: return launch_attention_mma<32>(st, B, S, L, H, q, k, v, ld, bias, key_mask, gate, out);
    case 48: return launch_attention_mma<48>(st, B, S, L, H, q, k, v, ld, bias, key_mask, gate, out);
    case 64: return launch_attention_mma<64>(st, B, S, L, H, q, k, v, ld, bias, key_mask, gate, out);
  }
  set_error("abx_pair_attention: head dim %d not instantiated (16, 32, 48, 64)", D);
  return ABX_ERR_INVALID;
}

// tcgen05 kernel (attention_tc5.cu).  bias_tiles: the pair bias pre-tiled, pre-scaled and with the key mask folded in
// (layout in include/abx_b200.h).
extern "C" int abx_pair_attention_tc5(void* stream, int B, int S, int L, int H, int D, const float* q, const float* k, const float* v,
                                      int ld, const float* bias_tiles, const float* gate, float* out) {
  using namespace abx;
  ABX_REQUIRE(B > 0 && S > 0 && L > 0 && H > 0 && q && k && v && bias_tiles && out, "abx_pair_attention_tc5: bad shape or null argument");
  ABX_REQUIRE(ld % 4 == 0 && ld >= H * D, "abx_pair_attention_tc5: ld must be a multiple of 4 and >= H*D");
  ABX_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                  (gate == nullptr || (uintptr_t)gate % 16 == 0),
              "abx_pair_attention_tc5: q, k, v, gate, out must be 16-byte aligned");
  ABX_REQUIRE((long long)B * S <= 65535, "abx_pair_attention_tc5: B*S exceeds 65535");
  ABX_REQUIRE(tc5_fits(L, D), "abx_pair_attention_tc5: head dim %d not supported (D in {16,32,48})", D);
  cudaStream_t st = (cudaStream_t)stream;
  switch (D) {
    case 16: return launch_attention_tc5<16>(st, B, S, L, H, q, k, v, ld, bias_tiles, gate, out);
    case 32: return launch_attention_tc5<32>(st, B, S, L, H, q, k, v, ld, bias_tiles, gate, out);
    default: return launch_attention_tc5<48>(st, B, S, L, H, q, k, v, ld, bias_tiles, gate, out);
  }
}

extern "C" int abx_pair_attention_tc5_supported(int L, int D) { return abx::tc5_fits(L, D) ? 1 : 0; }
