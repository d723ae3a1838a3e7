// Invariant Point Attention (abx/model/folding.py:47-132) on sm_100a — host pipeline of one layer-call:
//
//   node GEMM (gemm_tf32x3.cu)   x -> [q_scalar | kv_scalar | q_point_local | kv_point_local]        :69-86
//   ipa_pack_nodes_kernel        rigid transform of the points into the global frame, per-residue packing :89-93
//   ipa_fused_kernel             logits (scalar + point distance + pair bias), mask, online softmax, attention
//   (ipa_fused.cu)               over scalar / point values and over the pair activations — z read ONCE — inverse
//                                rigid transform, norms: the 2112-wide feature row                   :79-128
//   node GEMM (split-K)          final_proj over the feature row (+ bias + residual)                 :130-132
//
//   ipa_pair_bias_chunked_kernel sqrt(1/3) (z W_pair^T + b) in the chunked key-major layout the fused kernel streams
//   (ipa_fused.cu)               ([B, ceil(N/8), N, 100]): depends on z and the weights only, so IpaScore evaluates
//                                it once for its 8 weight-shared iterations                          :101-104
//
// The round-1 two-kernel core (tensor-core attention writing log-2 logits + pair-aggregation stream) is kept
// behind ABX_IPA_FUSED=0 for A/B measurements.
// Data layout (all fp32): z [B,N,N,128] row-major as the reference holds it.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"

namespace abx {

int launch_linear_f32(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w,
                      const float* bias, const float* residual, int relu, float* y, int ldy);
bool gemm_tf32x3_supported(int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw);
int launch_gemm_tf32x3_splitk(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                              int* splits_io, float* partials, int tile_n);
int gemm_backend();

constexpr int kH = ABX_IPA_H, kC = ABX_IPA_C, kCz = ABX_IPA_CZ, kFeat = ABX_IPA_FEAT;
constexpr int kSqk = 16, kSv = 16, kPqk = 4, kPv = 8;
constexpr int kQK = kSqk + 3 * kPqk;     // 28 floats per (residue, head) on the query/key side
constexpr int kVD = kSv + 3 * kPv;       // 40 floats per (residue, head) on the value side
constexpr int kProj = kH * (kSqk + kSqk + kSv) + 3 * kH * (kPqk + kPqk + kPv);   // 1152
constexpr int kOffKV = kH * kSqk;                   // 192: kv_scalar columns
constexpr int kOffQP = kOffKV + kH * (kSqk + kSv);  // 576: q_point_local columns
constexpr int kOffKVP = kOffQP + 3 * kH * kPqk;     // 720: kv_point_local columns
// feature row: [o_scalar 192 | o_point_local (r n) 288 | o_point_norm 96 | o_pair 1536]
constexpr int kFeatPt = kH * kSv, kFeatNorm = kFeatPt + 3 * kH * kPv, kFeatPair = kFeatNorm + kH * kPv;
static_assert(kProj == 1152 && kFeatPair + kH * kCz == kFeat, "IPA geometry");

// bulk L2 prefetch of kPfChunk bytes (16-byte aligned address): no destination, the lines just land in L2
constexpr unsigned kPfChunk = 4096;
__device__ __forceinline__ void prefetch_l2_chunk(const char* p) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(kPfChunk) : "memory");
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {   // streaming read: keep out of L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------------------
// pack: proj row -> Qdat/Kdat/Vdat with the points moved to the global frame (r3.rigids_apply, r3.py:9-16) and
// the query scalars pre-multiplied by sqrt(1/(3*16)) (folding.py:59,79).  One thread per (b, n, h, item):
// items 0-3 / 4-7 / 8-11 = float4 groups of the q / k / v scalars, 12-15 / 16-19 / 20-27 = q / k / v points.
// ---------------------------------------------------------------------------------------------------
constexpr int kPackItems = 3 * (kSqk / 4) + 2 * kPqk + kPv;   // 28

__global__ void __launch_bounds__(256) ipa_pack_kernel(int B, int N, const float* __restrict__ proj,
                                                       const float* __restrict__ rots, const float* __restrict__ trans,
                                                       float* __restrict__ Qdat, float* __restrict__ Kdat,
                                                       float* __restrict__ Vdat) {
  griddep_wait();                                    // proj comes from the node GEMM launched just before
  griddep_launch_dependents();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N * kH * kPackItems) return;
  const int item = idx % kPackItems, rest = idx / kPackItems;
  const int h = rest % kH, bn = rest / kH, b = bn / N, n = bn % N;
  const float* row = proj + (size_t)bn * kProj;
  const size_t o = ((size_t)(b * kH + h) * N + n);
  if (item < 12) {                                   // scalar channels, 4 at a time
    const int grp = item >> 2, c = 4 * (item & 3);
    if (grp == 0) {
      const float w_scalar = sqrtf(1.0f / (3.0f * kSqk));
      const float4 v = *reinterpret_cast<const float4*>(row + h * kSqk + c);
      *reinterpret_cast<float4*>(Qdat + o * kQK + c) = make_float4(v.x * w_scalar, v.y * w_scalar, v.z * w_scalar, v.w * w_scalar);
    } else if (grp == 1) {
      *reinterpret_cast<float4*>(Kdat + o * kQK + c) = *reinterpret_cast<const float4*>(row + kOffKV + h * (kSqk + kSv) + c);
    } else {
      *reinterpret_cast<float4*>(Vdat + o * kVD + c) = *reinterpret_cast<const float4*>(row + kOffKV + h * (kSqk + kSv) + kSqk + c);
    }
    return;
  }
  // one point: local coordinates are channel-major '(r n)', n = (h p)   folding.py:82,91,93
  const float* l;
  int stride;
  float* dst;
  if (item < 16) {
    const int p = item - 12;
    l = row + kOffQP + h * kPqk + p; stride = kH * kPqk; dst = Qdat + o * kQK + kSqk + 3 * p;
  } else {
    const int p = item - 16;                         // per head: 4 key points then 8 value points
    l = row + kOffKVP + h * (kPqk + kPv) + p; stride = kH * (kPqk + kPv);
    dst = (p < kPqk) ? (Kdat + o * kQK + kSqk + 3 * p) : (Vdat + o * kVD + kSv + 3 * (p - kPqk));
  }
  const float lx = l[0], ly = l[stride], lz = l[2 * stride];
  const float* R = rots + (size_t)bn * 9;
  const float* t = trans + (size_t)bn * 3;
  dst[0] = __ldg(t + 0) + (__ldg(R + 0) * lx + __ldg(R + 1) * ly + __ldg(R + 2) * lz);
  dst[1] = __ldg(t + 1) + (__ldg(R + 3) * lx + __ldg(R + 4) * ly + __ldg(R + 5) * lz);
  dst[2] = __ldg(t + 2) + (__ldg(R + 6) * lx + __ldg(R + 7) * ly + __ldg(R + 8) * lz);
}

// ---------------------------------------------------------------------------------------------------
// pair bias: one CTA per (b, i, 64-wide j tile); the z tile is staged in shared memory (row stride 132
// floats: conflict-free float4 reads with lanes on consecutive j), thread = (j, group of 3 heads).
// ---------------------------------------------------------------------------------------------------
constexpr int kBiasJ = 64, kZld = kCz + 4;

__global__ void __launch_bounds__(256) ipa_pair_bias_kernel(int N, const float* __restrict__ z,
                                                            const float* __restrict__ w_pair,
                                                            const float* __restrict__ b_pair, float* __restrict__ bias) {
  __shared__ __align__(16) float zs[kBiasJ * kZld];
  __shared__ __align__(16) float ws[kH * kCz];
  const int j0 = blockIdx.x * kBiasJ, i = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  for (int k = tid; k < kH * kCz / 4; k += 256)
    reinterpret_cast<float4*>(ws)[k] = __ldg(reinterpret_cast<const float4*>(w_pair) + k);
  const float4* zrow = reinterpret_cast<const float4*>(z + (((size_t)b * N + i) * N + j0) * kCz);
  const int nj = min(kBiasJ, N - j0);
  for (int k = tid; k < nj * (kCz / 4); k += 256) {
    int jj = k / (kCz / 4), c4 = k % (kCz / 4);
    *reinterpret_cast<float4*>(&zs[jj * kZld + 4 * c4]) = ldg_stream(zrow + k);
  }
  __syncthreads();
  const int jj = tid % kBiasJ, hg = tid / kBiasJ;     // heads 3*hg .. 3*hg+2
  if (jj >= nj) return;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll 8
  for (int c = 0; c < kCz; c += 4) {
    float4 zv = *reinterpret_cast<const float4*>(&zs[jj * kZld + c]);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      float4 wv = *reinterpret_cast<const float4*>(&ws[(3 * hg + u) * kCz + c]);
      acc[u] = fmaf(zv.x, wv.x, acc[u]); acc[u] = fmaf(zv.y, wv.y, acc[u]);
      acc[u] = fmaf(zv.z, wv.z, acc[u]); acc[u] = fmaf(zv.w, wv.w, acc[u]);
    }
  }
  const float w_pair_scale = sqrtf(1.0f / 3.0f);
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    int h = 3 * hg + u;
    bias[(((size_t)b * kH + h) * N + i) * N + j0 + jj] = w_pair_scale * (acc[u] + __ldg(b_pair + h));
  }
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------
// attention on the tensor cores (default): one warp per 16 query rows, FlashAttention-2 style fragments
// (mma.sync m16n8k8 TF32, 3xTF32 operand split = fp32-level accuracy).
//   logit = q_s.k_s + coef |Q-K|^2 + bias = [q_s, -2 coef Q] . [k_s, K] + coef |Q|^2 + coef |K|^2 + bias
// i.e. one 28-long (padded to 32) inner product plus a row term and a key term, so the O(N^2) part is an MMA.
// One pass over the keys with an online softmax: the log-2 logits are stored for the pair aggregation kernel
// (which normalises them with the final row max / sum), the probabilities multiply the 40-wide value rows.  The key
// and value rows of the (b,h) slice are staged once per CTA in shared memory (row strides 36 / 44 floats:
// conflict-free fragment loads).
// ---------------------------------------------------------------------------------------------------
constexpr int kMW = 8;                     // warps (16-row tiles) per CTA
constexpr int kKS = 36, kVS = 44;          // smem row strides of the key (32 used) and value (40 used) rows

__host__ __device__ inline size_t attn_mma_smem_floats(int N) {
  const int Np = (N + 31) & ~31;
  return (size_t)Np * (kKS + kVS) + 2 * (size_t)Np + (size_t)kMW * 16 * (kVD + 1);
}

__global__ void __launch_bounds__(kMW * 32, 2) ipa_attention_mma_kernel(
    int N, const float* __restrict__ Qdat, const float* __restrict__ Kdat, const float* __restrict__ Vdat,
    const float* __restrict__ bias, const float* __restrict__ mask, const float* __restrict__ rots,
    const float* __restrict__ trans, const float* __restrict__ point_weights, float* __restrict__ probs,
    float* __restrict__ stats, float* __restrict__ feats, const char* __restrict__ pf_base, unsigned pf_chunks) {
  extern __shared__ __align__(16) float sm[];
  griddep_wait();                            // Qdat / Kdat / Vdat come from the pack kernel launched just before
  griddep_launch_dependents();
  // L2 warm-up for the pair aggregation kernel that follows: DRAM is mostly idle while this kernel runs on the
  // tensor pipe, so lane 0 of every warp pulls one 4 KB chunk of [pf_base, pf_base + 4096 pf_chunks) — the head of
  // z, which the aggregation kernel reads first — into L2 per key chunk (warp w takes chunks w, w + #warps, ...)
  const unsigned pf_stride = gridDim.x * gridDim.y * gridDim.z * kMW;
  unsigned pf_idx = ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kMW + (threadIdx.x >> 5);
  const bool pf_lane = (threadIdx.x & 31) == 0;
  const int Np = (N + 31) & ~31;
  float* Ks = sm;                            // [Np][36]: k_s (16), K points (12), zeros (8)
  float* Vs = Ks + (size_t)Np * kKS;         // [Np][44]: v_s (16), V points (24), pad
  float* Ck = Vs + (size_t)Np * kVS;         // [Np] coef |K_j|^2
  float* Ms = Ck + Np;                       // [Np] mask_j
  float* Ot = Ms + Np;                       // per warp [16][41] output tile
  const int h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const size_t bh = (size_t)b * kH + h;
  const float pw = __ldg(point_weights + h);
  const float gamma = (pw > 20.f) ? pw : log1pf(expf(pw));                        // F.softplus  folding.py:96
  const float coef = -0.5f * sqrtf(1.0f / (3.0f * kPqk * 9.0f / 2.0f)) * gamma;   // -1/2 w_point gamma  :97-99

  // stage the packed key / value rows with cp.async (all copies in flight at once); pad columns and rows are zeroed
  for (int idx = threadIdx.x; idx < Np * (kKS / 4); idx += blockDim.x) {
    const int j = idx / (kKS / 4), c4 = idx % (kKS / 4);
    float* dst = Ks + (size_t)j * kKS + 4 * c4;
    if (j < N && c4 < kQK / 4) cp_async16(dst, Kdat + (bh * N + j) * kQK + 4 * c4);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int idx = threadIdx.x; idx < Np * (kVS / 4); idx += blockDim.x) {
    const int j = idx / (kVS / 4), c4 = idx % (kVS / 4);
    float* dst = Vs + (size_t)j * kVS + 4 * c4;
    if (j < N && c4 < kVD / 4) cp_async16(dst, Vdat + (bh * N + j) * kVD + 4 * c4);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  for (int j = threadIdx.x; j < Np; j += blockDim.x) {
    float s2 = 0.f;
    if (j < N) {
#pragma unroll
      for (int c = kSqk; c < kQK; ++c) { const float kv = Ks[(size_t)j * kKS + c]; s2 = fmaf(kv, kv, s2); }
    }
    Ck[j] = coef * s2;
    Ms[j] = (j < N) ? __ldg(mask + (size_t)b * N + j) : 0.f;
  }
  __syncthreads();

  const int r0 = (blockIdx.x * kMW + warp) * 16;
  if (r0 >= N) {                             // no query rows for this warp: issue its share of the prefetches and leave
    if (pf_lane)
      for (; pf_idx < pf_chunks; pf_idx += pf_stride) prefetch_l2_chunk(pf_base + (size_t)pf_idx * kPfChunk);
    return;
  }
  const int i0 = min(r0 + g, N - 1), i1 = min(r0 + g + 8, N - 1);
  // A fragments of [q_s, -2 coef Q, 0]: 4 k-steps of 8; row terms coef |Q_i|^2
  uint32_t qhi[4][4], qlo[4][4];
  float cq0 = 0.f, cq1 = 0.f;
  {
    const float* q0 = Qdat + (bh * N + i0) * kQK;
    const float* q1 = Qdat + (bh * N + i1) * kQK;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 8 * kk + t + 4 * e;
        float a0 = 0.f, a1 = 0.f;
        if (c < kQK) { a0 = __ldg(q0 + c); a1 = __ldg(q1 + c); }
        if (c >= kSqk && c < kQK) { a0 *= -2.f * coef; a1 *= -2.f * coef; }
        split_tf32(a0, qhi[kk][2 * e], qlo[kk][2 * e]);
        split_tf32(a1, qhi[kk][2 * e + 1], qlo[kk][2 * e + 1]);
      }
    }
    for (int c = kSqk; c < kQK; ++c) { const float a0 = __ldg(q0 + c), a1 = __ldg(q1 + c); cq0 = fmaf(a0, a0, cq0); cq1 = fmaf(a1, a1, cq1); }
    cq0 *= coef; cq1 *= coef;
  }
  const float mi0 = __ldg(mask + (size_t)b * N + i0), mi1 = __ldg(mask + (size_t)b * N + i1);
  const float* bias0 = bias + (bh * N + i0) * N;
  const float* bias1 = bias + (bh * N + i1) * N;

  // bias values of a 32-key chunk in fragment order (row g: [n][e], row g+8: [n][2+e]); loaded one chunk ahead
  auto load_bias = [&](int j0, float (&bv)[4][4]) {
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = j0 + 8 * n + 2 * t + e;
        const bool valid = j < N;
        bv[n][e] = valid ? __ldg(bias0 + j) : 0.f;
        bv[n][2 + e] = valid ? __ldg(bias1 + j) : 0.f;
      }
  };
  constexpr float kLog2e = 1.4426950408889634f;
  auto scores = [&](int j0, const float (&bv)[4][4], float (&s)[4][4]) {
#pragma unroll
    for (int n = 0; n < 4; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    // k-step outer, score tile inner: consecutive MMAs go to different accumulators (4 independent chains)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float* kr = Ks + (size_t)(j0 + 8 * n + g) * kKS + t + 8 * kk;
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
#pragma unroll
    for (int n = 0; n < 4; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = j0 + 8 * n + 2 * t + e;
        const bool valid = j < N;
        const float ck = Ck[j], mj = Ms[j];
        float a0 = (((s[n][e] + cq0) + ck) + bv[n][e]) * kLog2e;          // logits in units of log 2: exp(x) = exp2(x log2 e)
        float a1 = (((s[n][2 + e] + cq1) + ck) + bv[n][2 + e]) * kLog2e;
        a0 = valid ? ((mi0 * mj != 0.f) ? a0 : -FLT_MAX) : -INFINITY;     // mask_2d  folding.py:106-109
        a1 = valid ? ((mi1 * mj != 0.f) ? a1 : -FLT_MAX) : -INFINITY;
        s[n][e] = a0; s[n][2 + e] = a1;
      }
    }
  };

  // ---- single pass over the keys: logits (log-2 units) go to the `probs` buffer, the softmax is online
  //      (running max / sum per row, accumulator rescaled per 32-key chunk); the pair-aggregation kernel
  //      normalises the stored logits with the final (max, 1/sum) written to `stats`.
  float m0 = -FLT_MAX, m1 = -FLT_MAX, l0 = 0.f, l1 = 0.f;
  float oacc[5][4];
#pragma unroll
  for (int m = 0; m < 5; ++m) oacc[m][0] = oacc[m][1] = oacc[m][2] = oacc[m][3] = 0.f;
  float* pr0 = probs + (bh * N + i0) * N;
  float* pr1 = probs + (bh * N + i1) * N;
  const bool w0 = r0 + g < N, w1 = r0 + g + 8 < N, vec2 = (N % 2) == 0;
  float bnext[4][4];
  load_bias(0, bnext);
  for (int j0 = 0; j0 < Np; j0 += 32) {
    float s[4][4], bcur[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { bcur[n][0] = bnext[n][0]; bcur[n][1] = bnext[n][1]; bcur[n][2] = bnext[n][2]; bcur[n][3] = bnext[n][3]; }
    if (j0 + 32 < Np) load_bias(j0 + 32, bnext);
    if (pf_lane && pf_idx < pf_chunks) { prefetch_l2_chunk(pf_base + (size_t)pf_idx * kPfChunk); pf_idx += pf_stride; }
    scores(j0, bcur, s);
    float c0 = -FLT_MAX, c1 = -FLT_MAX;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      c0 = fmaxf(c0, fmaxf(s[n][0], s[n][1])); c1 = fmaxf(c1, fmaxf(s[n][2], s[n][3]));
      const int j = j0 + 8 * n + 2 * t;
      if (vec2 && j + 1 < N) {
        if (w0) *reinterpret_cast<float2*>(pr0 + j) = make_float2(s[n][0], s[n][1]);
        if (w1) *reinterpret_cast<float2*>(pr1 + j) = make_float2(s[n][2], s[n][3]);
      } else {
        if (j < N) { if (w0) pr0[j] = s[n][0]; if (w1) pr1[j] = s[n][2]; }
        if (j + 1 < N) { if (w0) pr0[j + 1] = s[n][1]; if (w1) pr1[j + 1] = s[n][3]; }
      }
    }
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1)); c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
    c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
    const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);
    const float k0 = exp2f(m0 - n0), k1 = exp2f(m1 - n1);
    m0 = n0; m1 = n1;
    l0 *= k0; l1 *= k1;
#pragma unroll
    for (int m = 0; m < 5; ++m) { oacc[m][0] *= k0; oacc[m][1] *= k0; oacc[m][2] *= k1; oacc[m][3] *= k1; }
    float pacc[5][4];
#pragma unroll
    for (int m = 0; m < 5; ++m) pacc[m][0] = pacc[m][1] = pacc[m][2] = pacc[m][3] = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float p00 = exp2f(s[n][0] - m0), p01 = exp2f(s[n][1] - m0);
      const float p10 = exp2f(s[n][2] - m1), p11 = exp2f(s[n][3] - m1);
      l0 += p00 + p01; l1 += p10 + p11;
      uint32_t phi[4], plo[4];                       // A fragment: k-index t <-> key 2t, t+4 <-> key 2t+1
      split_tf32(p00, phi[0], plo[0]); split_tf32(p10, phi[1], plo[1]);
      split_tf32(p01, phi[2], plo[2]); split_tf32(p11, phi[3], plo[3]);
      const float* vr = Vs + (size_t)(j0 + 8 * n + 2 * t) * kVS + g;
      uint32_t vh0[5], vl0[5], vh1[5], vl1[5];
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        split_tf32(vr[8 * m], vh0[m], vl0[m]);
        split_tf32(vr[kVS + 8 * m], vh1[m], vl1[m]);
      }
#pragma unroll
      for (int m = 0; m < 5; ++m) mma_tf32(pacc[m], phi, vl0[m], vl1[m]);
#pragma unroll
      for (int m = 0; m < 5; ++m) mma_tf32(pacc[m], plo, vh0[m], vh1[m]);
#pragma unroll
      for (int m = 0; m < 5; ++m) mma_tf32(pacc[m], phi, vh0[m], vh1[m]);
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) { oacc[m][0] += pacc[m][0]; oacc[m][1] += pacc[m][1]; oacc[m][2] += pacc[m][2]; oacc[m][3] += pacc[m][3]; }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
#pragma unroll
  for (int m = 0; m < 5; ++m) { oacc[m][0] *= inv0; oacc[m][1] *= inv0; oacc[m][2] *= inv1; oacc[m][3] *= inv1; }
  if (t == 0) {
    if (w0) *reinterpret_cast<float2*>(stats + (bh * N + r0 + g) * 2) = make_float2(m0, inv0);
    if (w1) *reinterpret_cast<float2*>(stats + (bh * N + r0 + g + 8) * 2) = make_float2(m1, inv1);
  }

  // ---- node features of these 16 rows (as in the SIMT kernel): via a per-warp tile [16][41]
  float* O = Ot + warp * 16 * (kVD + 1);
#pragma unroll
  for (int m = 0; m < 5; ++m) {
    O[g * (kVD + 1) + 8 * m + 2 * t] = oacc[m][0]; O[g * (kVD + 1) + 8 * m + 2 * t + 1] = oacc[m][1];
    O[(g + 8) * (kVD + 1) + 8 * m + 2 * t] = oacc[m][2]; O[(g + 8) * (kVD + 1) + 8 * m + 2 * t + 1] = oacc[m][3];
  }
  __syncwarp();
  for (int o = lane; o < 16 * kSv; o += 32) {
    const int r = o / kSv, c = o % kSv, i = r0 + r;
    if (i < N) feats[((size_t)b * N + i) * kFeat + h * kSv + c] = O[r * (kVD + 1) + c];      // 'b i h c -> b i (h c)'
  }
  for (int o = lane; o < 16 * kPv; o += 32) {
    const int r = o / kPv, p = o % kPv, i = r0 + r;
    if (i >= N) continue;
    const size_t bn = (size_t)b * N + i;
    float R[9], tr[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = __ldg(rots + bn * 9 + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) tr[k] = __ldg(trans + bn * 3 + k);
    float it[3], gp[3], l[3];                        // invert_rigids (r3.py:54-59) then rigids_apply  folding.py:121
#pragma unroll
    for (int k = 0; k < 3; ++k) it[k] = -(R[k] * tr[0] + R[3 + k] * tr[1] + R[6 + k] * tr[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) gp[k] = O[r * (kVD + 1) + kSv + 3 * p + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) l[k] = it[k] + (R[k] * gp[0] + R[3 + k] * gp[1] + R[6 + k] * gp[2]);
    float* f = feats + bn * kFeat;
#pragma unroll
    for (int k = 0; k < 3; ++k) f[kFeatPt + k * (kH * kPv) + h * kPv + p] = l[k];              // '(r n)'  folding.py:122
    f[kFeatNorm + h * kPv + p] = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2] + 1e-8f);       // :123
  }
}

// ---------------------------------------------------------------------------------------------------
// pair aggregation: o_pair[b,i,h,:] = sum_j a[b,h,i,j] z[b,i,j,:].  One CTA per (b, i): the 12 x N
// probabilities of the row are staged (transposed to [j][12]) in shared memory, each warp streams
// every 4th z[i,j,:] row straight from HBM into registers (one 16-byte load per lane covers the 512 B
// row), 48 FMAs per load; the four partial sums are reduced through shared memory.
// ---------------------------------------------------------------------------------------------------
constexpr int kAggThreads = 128, kAggWarps = kAggThreads / 32, kAggUnroll = 8;

__global__ void __launch_bounds__(kAggThreads) ipa_pair_aggregate_kernel(int N, const float* __restrict__ z,
                                                                         const float* __restrict__ probs,
                                                                         const float* __restrict__ stats,
                                                                         float* __restrict__ feats) {
  extern __shared__ __align__(16) float A[];        // [N][12] probabilities, then reused for the reduction
  griddep_wait();                                   // probs / stats come from the attention kernel launched just before
  griddep_launch_dependents();
  const int i = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int h = wid; h < kH; h += kAggWarps) {
    const float* pr = probs + (((size_t)b * kH + h) * N + i) * N;
    if (stats) {                                   // tensor-core attention path: log-2 logits + (row max, 1 / row sum)
      const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + ((size_t)b * kH + h) * N + i);
      for (int j = lane; j < N; j += 32) A[j * kH + h] = exp2f(__ldg(pr + j) - st.x) * st.y;
    } else {
      for (int j = lane; j < N; j += 32) A[j * kH + h] = __ldg(pr + j);
    }
  }
  __syncthreads();

  // packed fp32 pairs: 24 FFMA2 (fma.rn.f32x2) instead of 48 FFMA per 16-byte z load
  float2 acc[kH][2];
#pragma unroll
  for (int h = 0; h < kH; ++h) acc[h][0] = acc[h][1] = make_float2(0.f, 0.f);
  const float4* zrow = reinterpret_cast<const float4*>(z + ((size_t)b * N + i) * N * kCz) + lane;
  for (int j0 = wid; j0 < N; j0 += kAggWarps * kAggUnroll) {
    float4 zv[kAggUnroll];
#pragma unroll
    for (int u = 0; u < kAggUnroll; ++u) {
      int j = j0 + u * kAggWarps;
      zv[u] = (j < N) ? ldg_stream(zrow + (size_t)j * (kCz / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kAggUnroll; ++u) {
      int j = j0 + u * kAggWarps;
      if (j < N) {
        const float4* ap = reinterpret_cast<const float4*>(A + j * kH);
        float a[kH];
#pragma unroll
        for (int k = 0; k < kH / 4; ++k) { float4 v = ap[k]; a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w; }
        const float2 zlo = make_float2(zv[u].x, zv[u].y), zhi = make_float2(zv[u].z, zv[u].w);
#pragma unroll
        for (int h = 0; h < kH; ++h) {
          const float2 aa = make_float2(a[h], a[h]);
          acc[h][0] = __ffma2_rn(aa, zlo, acc[h][0]);
          acc[h][1] = __ffma2_rn(aa, zhi, acc[h][1]);
        }
      }
    }
  }
  __syncthreads();                                   // done with A; reuse as red[warp][h][128]
  float4* red = reinterpret_cast<float4*>(A);
#pragma unroll
  for (int h = 0; h < kH; ++h)
    red[(wid * kH + h) * (kCz / 4) + lane] = make_float4(acc[h][0].x, acc[h][0].y, acc[h][1].x, acc[h][1].y);
  __syncthreads();
  float4* out = reinterpret_cast<float4*>(feats + ((size_t)b * N + i) * kFeat + kFeatPair);   // 'b i h c -> b i (h c)'
  for (int o = threadIdx.x; o < kH * kCz / 4; o += kAggThreads) {
    float4 s = red[o];
#pragma unroll
    for (int w = 1; w < kAggWarps; ++w) {
      float4 v = red[w * kH * (kCz / 4) + o];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[o] = s;
  }
}

__host__ inline size_t agg_smem_bytes(int N) {
  size_t a = (size_t)N * kH, b = (size_t)kAggWarps * kH * kCz;
  return (a > b ? a : b) * sizeof(float);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

constexpr int kMaxSplits = 8;   // split-K factor of the final projection (2112 -> 256) when B*N is small

// fused path (ipa_fused.cu)
size_t ipa_fused_qp_floats(int B, int N);
size_t ipa_fused_kvp_floats(int B, int N);
size_t ipa_pair_bias_floats(int B, int N);
int ipa_watchdog_read(unsigned long long* out);
int ipa_prof(int enable, unsigned long long* out64);
int launch_ipa_pair_bias(cudaStream_t s, int B, int N, const float* z, const float* w_pair, const float* b_pair, float* bias);
int launch_ipa_pack_nodes(cudaStream_t s, int B, int N, const float* proj, const float* rots, const float* trans, float* Qp,
                          float* KVp);
int launch_ipa_fused(cudaStream_t s, int B, int N, const float* Qp, const float* KVp, const float* bias, const float* mask,
                     const float* rots, const float* trans, const float* point_weights, const float* z, float* feats);

struct IpaWorkspace {
  float *proj, *Qdat, *Kdat, *Vdat, *probs, *stats, *feats, *bias, *partials;
  size_t total;
};

// The GEMM main loop costs about the same per 32-wide k-slab whatever the tile width, so the final projection
// uses the widest tiles (128 columns: 2 per row block) and splits K so that every SM gets at most one tile.
static int final_proj_splits(int M) {
  const int tiles = ceil_div(M, 128) * (kC / 128);
  int s = 148 / tiles;
  return s < 1 ? 1 : (s > kMaxSplits ? kMaxSplits : s);
}

// ABX_IPA_FUSED=0 selects the round-1 two-kernel pipeline (tensor-core attention writing log-2 logits, then the
// pair-aggregation stream) for A/B measurements at N <= 640; default: the fused kernel of ipa_fused.cu.
static bool ipa_fused() {
  static bool v = [] { const char* e = getenv("ABX_IPA_FUSED"); return !(e && e[0] == '0'); }();
  return v;
}

static IpaWorkspace carve(void* base, int B, int N, bool with_bias, bool with_feats) {
  IpaWorkspace w;
  size_t off = 0;
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(p + off); off += align_up(floats * sizeof(float)); return r; };
  const size_t bn = (size_t)B * N;
  w.proj = take(bn * kProj);
  w.Qdat = take(bn * kH * kQK);                          // fused path: packed queries [B,N,12,28]
  w.Kdat = take(bn * kH * (kQK + kVD));                  // fused path: packed keys + values [B,N,816]; else keys [B,H,N,28]
  w.Vdat = take(bn * kH * kVD);
  w.probs = ipa_fused() ? nullptr : take(bn * kH * N);
  w.stats = take(bn * kH * 2);
  w.feats = with_feats ? take(bn * kFeat) : nullptr;
  w.partials = (with_feats && final_proj_splits((int)bn) > 1) ? take((size_t)final_proj_splits((int)bn) * bn * kC) : nullptr;
  {
    const size_t head_major = bn * kH * N, chunked = ipa_pair_bias_floats(B, N);
    w.bias = with_bias ? take(head_major > chunked ? head_major : chunked) : nullptr;
  }
  w.total = off;
  return w;
}

// ABX_IPA_PDL=0 launches the kernels of a layer-call the ordinary way; default: programmatic dependent launch,
// every kernel of the chain (node GEMM, pack, fused attention, final projection, finalize) may start
// while its predecessor drains and executes griddep_wait() before touching memory.
static bool ipa_pdl() {
  static bool v = [] { const char* e = getenv("ABX_IPA_PDL"); return !(e && e[0] == '0'); }();
  return v;
}

struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(pdl_scope_active()) { pdl_scope_set(on); }
  ~PdlScope() { pdl_scope_set(prev); }
};

#define ABX_LAUNCH(name, kernel, grid, block, smem, s, ...)                                  \
  do {                                                                                       \
    const cudaError_t le__ = launch_kernel(kernel, grid, block, smem, s, __VA_ARGS__);       \
    count_launch();                                                                          \
    if (le__ != cudaSuccess) {                                                               \
      set_error("launch of " name " failed: %s", cudaGetErrorString(le__));                  \
      return ABX_ERR_CUDA;                                                                   \
    }                                                                                        \
    if (int rc__ = check_launch(name)) return rc__;                                          \
  } while (0)

static int ipa_features(cudaStream_t s, int B, int N, const float* x, const float* z, const float* mask,
                        const float* rots, const float* trans, const abx_ipa_weights* w, const float* pair_bias,
                        float* feats, const IpaWorkspace& ws) {
  const int M = B * N;
  int rc;
  PdlScope pdl(ipa_pdl());
  // node projections (folding.py:69-86): four Linear layers into one [M, 1152] buffer — one GEMM when the
  // caller provides the row-concatenated weights
  if (w->w_proj_cat) {
    if ((rc = launch_linear_f32(s, M, kProj, kC, x, kC, w->w_proj_cat, w->b_proj_cat, nullptr, 0, ws.proj, kProj))) return rc;
  } else {
    if ((rc = launch_linear_f32(s, M, kH * kSqk, kC, x, kC, w->w_q_scalar, w->b_q_scalar, nullptr, 0, ws.proj, kProj))) return rc;
    if ((rc = launch_linear_f32(s, M, kH * (kSqk + kSv), kC, x, kC, w->w_kv_scalar, w->b_kv_scalar, nullptr, 0, ws.proj + kOffKV, kProj))) return rc;
    if ((rc = launch_linear_f32(s, M, 3 * kH * kPqk, kC, x, kC, w->w_q_point, w->b_q_point, nullptr, 0, ws.proj + kOffQP, kProj))) return rc;
    if ((rc = launch_linear_f32(s, M, 3 * kH * (kPqk + kPv), kC, x, kC, w->w_kv_point, w->b_kv_point, nullptr, 0, ws.proj + kOffKVP, kProj))) return rc;
  }
  const size_t msmem = attn_mma_smem_floats(N) * sizeof(float);
  if (ipa_fused() || msmem > 227 * 1024) {
    if (pair_bias == nullptr) {
      ABX_REQUIRE(ws.bias != nullptr, "ipa: no pair bias and no workspace room for it");
      if ((rc = launch_ipa_pair_bias(s, B, N, z, w->w_pair, w->b_pair, ws.bias))) return rc;
      pair_bias = ws.bias;
    }
    if ((rc = launch_ipa_pack_nodes(s, B, N, ws.proj, rots, trans, ws.Qdat, ws.Kdat))) return rc;
    return launch_ipa_fused(s, B, N, ws.Qdat, ws.Kdat, pair_bias, mask, rots, trans, w->point_weights, z, feats);
  }

  // round-1 two-kernel path (A/B only): it reads a head-major bias, which it always evaluates itself
  ABX_REQUIRE(ws.probs != nullptr && ws.bias != nullptr, "ipa: the two-kernel path needs the probability and bias workspace");
  {
    PdlScope plain(false);
    ipa_pair_bias_kernel<<<dim3(ceil_div(N, kBiasJ), N, B), 256, 0, s>>>(N, z, w->w_pair, w->b_pair, ws.bias);
    count_launch();
    if ((rc = check_launch("ipa_pair_bias_kernel"))) return rc;
    pair_bias = ws.bias;
  }
  ABX_LAUNCH("ipa_pack_kernel", ipa_pack_kernel, dim3(ceil_div(M * kH * kPackItems, 256)), dim3(256), 0, s, B, N, ws.proj, rots,
             trans, ws.Qdat, ws.Kdat, ws.Vdat);
  ABX_CUDA(cudaFuncSetAttribute(ipa_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  ABX_LAUNCH("ipa_attention_mma_kernel", ipa_attention_mma_kernel, dim3(ceil_div(N, 16 * kMW), kH, B), dim3(kMW * 32), msmem, s, N,
             ws.Qdat, ws.Kdat, ws.Vdat, pair_bias, mask, rots, trans, w->point_weights, ws.probs, ws.stats, feats,
             reinterpret_cast<const char*>(z), 0u);
  const size_t gsmem = agg_smem_bytes(N);
  ABX_CUDA(cudaFuncSetAttribute(ipa_pair_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
  ABX_LAUNCH("ipa_pair_aggregate_kernel", ipa_pair_aggregate_kernel, dim3(N, B), dim3(kAggThreads), gsmem, s, N, z, ws.probs, ws.stats, feats);
  return ABX_OK;
}

// out = sum_s partials[s] + bias (+ residual): the reduction of the split-K final projection
__global__ void __launch_bounds__(256) ipa_finalize_kernel(int MN4, int splits, const float4* __restrict__ partials,
                                                           const float4* __restrict__ bias, const float4* __restrict__ residual,
                                                           float4* __restrict__ out) {
  griddep_wait();                                   // partials come from the split-K GEMM launched just before
  griddep_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN4) return;
  float4 a = partials[i];
  for (int s = 1; s < splits; ++s) {
    const float4 p = partials[(size_t)s * MN4 + i];
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
  }
  if (bias) { const float4 b = __ldg(bias + (i % (kC / 4))); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  if (residual) { const float4 r = residual[i]; a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w; }
  out[i] = a;
}

static int ipa_check(const char* fn, int B, int N, const void* x, const void* z, const void* mask, const void* rots,
                     const void* trans, const abx_ipa_weights* w) {
  ABX_REQUIRE(B > 0 && N > 0 && x && z && mask && rots && trans && w, "%s: bad shape or null argument", fn);
  ABX_REQUIRE(N <= 1536, "%s: N=%d exceeds the supported maximum of 1536 residues", fn, N);
  ABX_REQUIRE(w->w_q_scalar && w->w_kv_scalar && w->w_q_point && w->w_kv_point && w->w_pair && w->b_pair &&
                  w->point_weights && w->w_final, "%s: null weight pointer", fn);
  ABX_REQUIRE((uintptr_t)z % 16 == 0 && (uintptr_t)x % 16 == 0, "%s: x and z must be 16-byte aligned", fn);
  return ABX_OK;
}

}  // namespace abx

using namespace abx;

extern "C" size_t abx_ipa_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return carve(nullptr, B, N, true, true).total;
}

extern "C" int abx_ipa_watchdog_read(unsigned long long* out8) {
  ABX_REQUIRE(out8 != nullptr, "abx_ipa_watchdog_read: null argument");
  return ipa_watchdog_read(out8);
}

extern "C" int abx_ipa_profile(int enable, unsigned long long* out64) { return ipa_prof(enable, out64); }

extern "C" size_t abx_ipa_pair_bias_floats(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return ipa_pair_bias_floats(B, N);
}

extern "C" int abx_ipa_pair_bias(void* stream, int B, int N, const float* z, const float* w_pair, const float* b_pair,
                                 float* pair_bias) {
  ABX_REQUIRE(B > 0 && N > 0 && z && w_pair && b_pair && pair_bias, "abx_ipa_pair_bias: bad shape or null argument");
  ABX_REQUIRE((uintptr_t)z % 16 == 0 && (uintptr_t)w_pair % 16 == 0 && (uintptr_t)pair_bias % 16 == 0,
              "abx_ipa_pair_bias: z, w_pair and pair_bias must be 16-byte aligned");
  return launch_ipa_pair_bias((cudaStream_t)stream, B, N, z, w_pair, b_pair, pair_bias);
}

extern "C" int abx_ipa_attention_features(void* stream, int B, int N, const float* x, const float* z, const float* mask,
                                          const float* rots, const float* trans, const abx_ipa_weights* w,
                                          const float* pair_bias, float* feats, void* workspace, size_t workspace_bytes) {
  int rc = ipa_check("abx_ipa_attention_features", B, N, x, z, mask, rots, trans, w);
  if (rc) return rc;
  ABX_REQUIRE(feats && workspace, "abx_ipa_attention_features: null output or workspace");
  IpaWorkspace ws = carve(workspace, B, N, pair_bias == nullptr || !ipa_fused(), false);
  ABX_REQUIRE(workspace_bytes >= ws.total, "abx_ipa_attention_features: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
  return ipa_features((cudaStream_t)stream, B, N, x, z, mask, rots, trans, w, pair_bias, feats, ws);
}

extern "C" int abx_ipa_forward(void* stream, int B, int N, const float* x, const float* z, const float* mask,
                               const float* rots, const float* trans, const abx_ipa_weights* w, const float* pair_bias,
                               const float* residual, float* out, void* workspace, size_t workspace_bytes) {
  int rc = ipa_check("abx_ipa_forward", B, N, x, z, mask, rots, trans, w);
  if (rc) return rc;
  ABX_REQUIRE(out && workspace, "abx_ipa_forward: null output or workspace");
  IpaWorkspace ws = carve(workspace, B, N, pair_bias == nullptr || !ipa_fused(), true);
  ABX_REQUIRE(workspace_bytes >= ws.total, "abx_ipa_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
  cudaStream_t s = (cudaStream_t)stream;
  PdlScope pdl(ipa_pdl());
  if ((rc = ipa_features(s, B, N, x, z, mask, rots, trans, w, pair_bias, ws.feats, ws))) return rc;
  // final_proj (folding.py:130-132) + the residual of score_network.py:128.  With few rows the 2112-long
  // reduction is split across CTAs (split-K) and summed by a small kernel; otherwise one GEMM with fused epilogue.
  const int M = B * N;
  int splits = final_proj_splits(M);
  const bool vec = ((uintptr_t)out % 16 == 0) && (!residual || (uintptr_t)residual % 16 == 0) && (!w->b_final || (uintptr_t)w->b_final % 16 == 0);
  if (splits > 1 && ws.partials && vec && gemm_backend() != 1 && gemm_tf32x3_supported(M, kC, kFeat, ws.feats, kFeat, w->w_final, kFeat)) {
    if ((rc = launch_gemm_tf32x3_splitk(s, M, kC, kFeat, ws.feats, kFeat, w->w_final, kFeat, &splits, ws.partials, 128))) return rc;
    const int MN4 = M * kC / 4;
    ABX_LAUNCH("ipa_finalize_kernel", ipa_finalize_kernel, dim3(ceil_div(MN4, 256)), dim3(256), 0, s, MN4, splits,
               reinterpret_cast<const float4*>(ws.partials), reinterpret_cast<const float4*>(w->b_final),
               reinterpret_cast<const float4*>(residual), reinterpret_cast<float4*>(out));
    return ABX_OK;
  }
  return launch_linear_f32(s, M, kC, kFeat, ws.feats, kFeat, w->w_final, w->b_final, residual, 0, out, kC);
}
