// Fused Invariant Point Attention core (abx/model/folding.py:79-128) — ONE kernel for
//   logits (scalar q.k + point distances + pair bias, mask)  :79-109
//   softmax over the keys                                     :110
//   attention over scalar / point values, inverse rigid transform, point norms   :114-123
//   attention over the pair activations  o_pair[i,h,:] = sum_j a[h,i,j] z[i,j,:]  :126-127
// with the pair tensor z [B,N,N,128] read exactly once and no [B,H,N,N] logits / probability tensor.
//
// Work decomposition.  One CTA = a tile of R <= 20 consecutive query rows of one batch element, one warp per
// query row, all 12 heads; the CTA walks the keys in chunks of 8.  Everything the loop consumes arrives through
// the bulk-copy engine (cp.async.bulk + mbarrier transaction counts), so no global load sits on a warp's
// critical path and the bytes in flight do not occupy registers:
//   * z[b,i,:,:] (N x 512 B, contiguous): 2 KB pieces into a warp-private 3-slot ring (4 KB in flight per warp,
//     80 KB per SM), no cross-warp synchronisation;
//   * the pair bias of row i, stored key-major [B,N,N,12] by ipa_pair_bias_kernel: one 384-byte copy per chunk
//     into a warp-private double buffer, which then receives the chunk's probabilities in place;
//   * the packed key / value operands of the 8 keys (per key 12 x (28 + 40) floats, the 12 key-side logit
//     constants and the key mask: 3344 B), shared by the R rows of the tile: a double-buffered chunk filled by
//     warp 0 and handed back through a count-R mbarrier.  The 126 MB L2 serves these re-reads at > 30 TB/s
//     (tools/l2_probe.cu), so they do not compete with the HBM stream.
// Per chunk each warp computes its row's 12 x 8 logits with lanes on (key, 3 heads) in exact fp32 SIMT arithmetic
//   logit = [q_s, -2c Q] . [k_s, K] + c|Q|^2 + c|K|^2 + bias      (28-long packed FFMA2 dot product, c = -gamma_h w_point / 2)
// runs an online softmax (running max / sum per head, accumulators rescaled only when a maximum moves), leaves the
// 8 x 12 probabilities in shared memory, accumulates the 480-wide value row (lane = float4 slices) and then the
// 12 x 128 pair row (lane = 4 channels, 24 FFMA2 per 16 bytes of z).
// The tile height R is chosen on the host so that B * ceil(N / R) tiles fill the 148 SMs in whole rounds
// (B = 8, N = 350: R = 20 -> 144 CTAs, one per SM).
#include <float.h>

#include "common.cuh"

namespace abx {

namespace {

constexpr int kH = ABX_IPA_H, kCz = ABX_IPA_CZ, kFeat = ABX_IPA_FEAT;
constexpr int kSqk = 16, kSv = 16, kPqk = 4, kPv = 8;
constexpr int kQK = kSqk + 3 * kPqk;          // 28: query / key operand per head
constexpr int kVD = kSv + 3 * kPv;            // 40: value operand per head
constexpr int kQRow = kH * kQK;               // 336 floats per residue: packed queries
constexpr int kVOff = kH * kQK;               // values follow the keys inside a packed key/value row
constexpr int kKVRow = kH * (kQK + kVD);      // 816 floats per residue: packed keys + values
constexpr int kKVStride = kKVRow + 4;         // shared-memory row stride: conflict-free float4 reads with lanes on keys
constexpr int kChunk = 8;                     // keys per key/value chunk
constexpr int kZKeys = 4, kZSlots = 3;        // z ring: 3 slots of 4 keys (2 KB) per warp
constexpr int kZSlotFloats = kZKeys * kCz;
constexpr int kMaxRows = 20;                  // query rows (= warps) per CTA
constexpr int kStatFloats = 32;               // per warp: alpha[12], 1/sum[12], pad
constexpr int kFeatPt = kH * kSv, kFeatNorm = kFeatPt + 3 * kH * kPv, kFeatPair = kFeatNorm + kH * kPv;
constexpr float kLog2e = 1.4426950408889634f;

__host__ __device__ inline size_t fused_smem_floats(int R) {
  return (size_t)2 * kChunk * kKVStride + (size_t)R * (kZSlots * kZSlotFloats + kChunk * kH + kQRow + kStatFloats);
}
__host__ inline size_t fused_smem_bytes(int R) { return fused_smem_floats(R) * sizeof(float) + (4 + (size_t)R * kZSlots) * sizeof(uint64_t); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared bulk copy (bytes and both addresses multiples of 16) completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(float* smem_dst, const float* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// orders this thread's earlier generic-proxy shared-memory accesses before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

}  // namespace

// ---------------------------------------------------------------------------------------------------
// pack: row of the fused node projection [q_scalar 192 | kv_scalar 384 | q_point_local 144 | kv_point_local 432]
// (folding.py:69-86) -> packed queries Qp [B,N,12,28] = (q_s * sqrt(1/48), 4 query points in the global frame) and
// packed keys / values KVp [B,N,816] = 12 x (k_s 16, 4 key points) then 12 x (v_s 16, 8 value points), points
// moved to the global frame (r3.rigids_apply, r3.py:9-16).  One thread per (b, n, h, item): items 0-3 / 4-7 / 8-11 =
// float4 groups of the q / k / v scalars, 12-15 / 16-19 / 20-27 = q / k / v points.
// ---------------------------------------------------------------------------------------------------
constexpr int kProj = 1152, kOffKV = kH * kSqk, kOffQP = kOffKV + kH * (kSqk + kSv), kOffKVP = kOffQP + 3 * kH * kPqk;
constexpr int kPackItems = 3 * (kSqk / 4) + 2 * kPqk + kPv;   // 28

__global__ void __launch_bounds__(256) ipa_pack_nodes_kernel(int B, int N, const float* __restrict__ proj,
                                                             const float* __restrict__ rots, const float* __restrict__ trans,
                                                             float* __restrict__ Qp, float* __restrict__ KVp) {
  griddep_wait();                                    // proj comes from the node GEMM launched just before
  griddep_launch_dependents();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N * kH * kPackItems) return;
  const int item = idx % kPackItems, rest = idx / kPackItems;
  const int h = rest % kH, bn = rest / kH;
  const float* row = proj + (size_t)bn * kProj;
  float* q = Qp + (size_t)bn * kQRow + h * kQK;
  float* k = KVp + (size_t)bn * kKVRow + h * kQK;
  float* v = KVp + (size_t)bn * kKVRow + kVOff + h * kVD;
  if (item < 12) {                                   // scalar channels, 4 at a time
    const int grp = item >> 2, c = 4 * (item & 3);
    if (grp == 0) {
      const float w_scalar = sqrtf(1.0f / (3.0f * kSqk));                       // folding.py:59,79
      const float4 s = *reinterpret_cast<const float4*>(row + h * kSqk + c);
      *reinterpret_cast<float4*>(q + c) = make_float4(s.x * w_scalar, s.y * w_scalar, s.z * w_scalar, s.w * w_scalar);
    } else if (grp == 1) {
      *reinterpret_cast<float4*>(k + c) = *reinterpret_cast<const float4*>(row + kOffKV + h * (kSqk + kSv) + c);
    } else {
      *reinterpret_cast<float4*>(v + c) = *reinterpret_cast<const float4*>(row + kOffKV + h * (kSqk + kSv) + kSqk + c);
    }
    return;
  }
  // one point: local coordinates are channel-major '(r n)', n = (h p)   folding.py:82,91,93
  const float* l;
  int stride;
  float* dst;
  if (item < 16) {
    const int p = item - 12;
    l = row + kOffQP + h * kPqk + p; stride = kH * kPqk; dst = q + kSqk + 3 * p;
  } else {
    const int p = item - 16;                         // per head: 4 key points then 8 value points
    l = row + kOffKVP + h * (kPqk + kPv) + p; stride = kH * (kPqk + kPv);
    dst = (p < kPqk) ? (k + kSqk + 3 * p) : (v + kSv + 3 * (p - kPqk));
  }
  const float lx = l[0], ly = l[stride], lz = l[2 * stride];
  const float* Rm = rots + (size_t)bn * 9;
  const float* t = trans + (size_t)bn * 3;
  dst[0] = __ldg(t + 0) + (__ldg(Rm + 0) * lx + __ldg(Rm + 1) * ly + __ldg(Rm + 2) * lz);
  dst[1] = __ldg(t + 1) + (__ldg(Rm + 3) * lx + __ldg(Rm + 4) * ly + __ldg(Rm + 5) * lz);
  dst[2] = __ldg(t + 2) + (__ldg(Rm + 6) * lx + __ldg(Rm + 7) * ly + __ldg(Rm + 8) * lz);
}

// ---------------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxRows * 32, 1)
ipa_fused_kernel(int N, int R, int tiles_per_b, const float* __restrict__ Qp, const float* __restrict__ KVp,
                 const float* __restrict__ bias, const float* __restrict__ mask, const float* __restrict__ rots,
                 const float* __restrict__ trans, const float* __restrict__ point_weights, const float* __restrict__ z,
                 float* __restrict__ feats) {
  extern __shared__ __align__(128) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x / tiles_per_b, i0 = (blockIdx.x % tiles_per_b) * R;
  const int nvalid = min(R, N - i0);                // rows (warps) of this tile that exist
  float* KV = sm;                                               // [2][8][820]
  float* ZR = KV + 2 * kChunk * kKVStride + (size_t)warp * kZSlots * kZSlotFloats;   // this warp's z ring [3][512]
  float* PS = sm + 2 * kChunk * kKVStride + (size_t)R * kZSlots * kZSlotFloats + (size_t)warp * kChunk * kH;   // [8][12]
  float* QS = sm + 2 * kChunk * kKVStride + (size_t)R * (kZSlots * kZSlotFloats + kChunk * kH) + (size_t)warp * kQRow;
  float* ST = sm + 2 * kChunk * kKVStride + (size_t)R * (kZSlots * kZSlotFloats + kChunk * kH + kQRow) + (size_t)warp * kStatFloats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + fused_smem_floats(R));
  uint64_t* kvfull = bars;                          // [2] bulk copies of a key/value chunk have landed
  uint64_t* kvempty = bars + 2;                     // [2] every row warp is done with the chunk
  uint64_t* zfull = bars + 4 + warp * kZSlots;      // [3] this warp's z slots

  if (threadIdx.x == 0) {
    mbar_init(kvfull, 1); mbar_init(kvfull + 1, 1);
    mbar_init(kvempty, nvalid); mbar_init(kvempty + 1, nvalid);
  }
  if (lane == 0) { mbar_init(zfull, 1); mbar_init(zfull + 1, 1); mbar_init(zfull + 2, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (warp >= nvalid) return;                       // no CTA-wide barrier below this line

  const int i = i0 + warp;
  const size_t bn = (size_t)b * N + i;
  const int nchunks = (N + kChunk - 1) / kChunk, nq = (N + kZKeys - 1) / kZKeys;
  const float* zrow = z + bn * (size_t)N * kCz;
  auto issue_z = [&](int q) {                       // lane 0 only: keys 4q .. 4q+3 of this row into slot q % 3
    const int slot = q % kZSlots;
    const uint32_t bytes = (uint32_t)min(kZKeys, N - q * kZKeys) * kCz * sizeof(float);
    mbar_expect_tx(zfull + slot, bytes);
    bulk_g2s(ZR + slot * kZSlotFloats, zrow + (size_t)q * kZSlotFloats, bytes, zfull + slot);
  };
  // z is an input of the whole layer (not produced by the preceding kernels of the chain): start streaming now
  if (lane == 0)
    for (int q = 0; q < kZSlots && q < nq; ++q) issue_z(q);

  griddep_wait();                                   // Qp / KVp come from the kernels launched just before
  griddep_launch_dependents();

  auto issue_kv = [&](int c) {                      // warp 0, all lanes: chunk c -> buffer c & 1
    const int buf = c & 1, nk = min(kChunk, N - c * kChunk);
    if (lane == 0) mbar_expect_tx(kvfull + buf, (uint32_t)nk * kKVRow * sizeof(float));
    __syncwarp();
    if (lane < nk)
      bulk_g2s(KV + (size_t)(buf * kChunk + lane) * kKVStride, KVp + ((size_t)b * N + c * kChunk + lane) * kKVRow,
               kKVRow * sizeof(float), kvfull + buf);
  };
  if (warp == 0) {
    issue_kv(0);
    if (nchunks > 1) issue_kv(1);
  }
  // this row's packed queries -> shared memory (read as warp-wide broadcasts below)
  {
    const float4* src = reinterpret_cast<const float4*>(Qp + bn * kQRow);
    for (int k = lane; k < kQRow / 4; k += 32) reinterpret_cast<float4*>(QS)[k] = __ldg(src + k);
  }
  __syncwarp();

  // lane roles. logits: (key kk = lane & 7, heads 3 hq .. 3 hq + 2); values: float4 slices lane + 32 u of the 480-wide
  // value row (120 slices: lanes 24-31 hold three); pair row: channels 4 lane .. 4 lane + 3
  const int kk = lane & 7, hq = lane >> 3;
  float coef[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const float pw = __ldg(point_weights + 3 * hq + t);
    const float gamma = (pw > 20.f) ? pw : log1pf(expf(pw));                       // F.softplus  folding.py:96
    coef[t] = -0.5f * sqrtf(1.0f / (3.0f * kPqk * 9.0f / 2.0f)) * gamma;           // -1/2 w_point gamma  :97-99
  }
  const float mi = __ldg(mask + (size_t)b * N + i);
  const float* mrow = mask + (size_t)b * N;
  const float* brow = bias + (((size_t)b * kH + 3 * hq) * N + i) * N;             // head 3 hq + t at brow + t N N
  const size_t bstride = (size_t)N * N;
  int hd[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) hd[u] = min((lane + 32 * u) / (kVD / 4), kH - 1);
  const bool has3 = lane + 96 < kH * kVD / 4;

  float m[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, lsum[3] = {0.f, 0.f, 0.f};
  float4 oval[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) oval[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 acc[kH][2];
#pragma unroll
  for (int h = 0; h < kH; ++h) acc[h][0] = acc[h][1] = make_float2(0.f, 0.f);

  float bnext[3], mnext;
  {
    const bool v0 = kk < N;
#pragma unroll
    for (int t = 0; t < 3; ++t) bnext[t] = v0 ? __ldg(brow + t * bstride + kk) : 0.f;
    mnext = v0 ? __ldg(mrow + kk) : 0.f;
  }

  for (int c = 0; c < nchunks; ++c) {
    const int j0 = c * kChunk, nk = min(kChunk, N - j0), buf = c & 1;
    const float* kvp = KV + (size_t)buf * kChunk * kKVStride;
    const bool valid = j0 + kk < N;
    float bcur[3] = {bnext[0], bnext[1], bnext[2]};
    const float mj = mnext;
    {                                               // bias / mask of the next chunk: in flight during this one
      const int jn = j0 + kChunk + kk;
      const bool vn = jn < N;
#pragma unroll
      for (int t = 0; t < 3; ++t) bnext[t] = vn ? __ldg(brow + t * bstride + jn) : 0.f;
      mnext = vn ? __ldg(mrow + jn) : 0.f;
    }
    mbar_wait(kvfull + buf, (c >> 1) & 1);

    // ---- logits of (key kk, heads 3 hq + t), in units of log 2 ----
    float s[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const float4* kp = reinterpret_cast<const float4*>(kvp + kk * kKVStride + (3 * hq + t) * kQK);
      const float4* qp = reinterpret_cast<const float4*>(QS + (3 * hq + t) * kQK);
      float2 dot = make_float2(0.f, 0.f), dd = make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kSqk / 4; ++u) {
        const float4 kv = kp[u], qv = qp[u];
        dot = ffma2(make_float2(qv.x, qv.y), make_float2(kv.x, kv.y), dot);
        dot = ffma2(make_float2(qv.z, qv.w), make_float2(kv.z, kv.w), dot);
      }
#pragma unroll
      for (int u = kSqk / 4; u < kQK / 4; ++u) {
        const float4 kv = kp[u], qv = qp[u];
        const float2 d0 = make_float2(qv.x - kv.x, qv.y - kv.y), d1 = make_float2(qv.z - kv.z, qv.w - kv.w);
        dd = ffma2(d0, d0, dd);
        dd = ffma2(d1, d1, dd);
      }
      const float lg = (((dot.x + dot.y) + coef[t] * (dd.x + dd.y)) + bcur[t]) * kLog2e;
      s[t] = valid ? ((mi * mj != 0.f) ? lg : -FLT_MAX) : -INFINITY;               // mask_2d  folding.py:106-109
    }
    // ---- online softmax: running max per head over the 8 lanes that share hq ----
    float alpha[3];
    bool moved = false;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      float cm = s[t];
      cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
      cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
      cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, 4));
      const float mn = fmaxf(m[t], cm);
      alpha[t] = exp2f(m[t] - mn);
      moved |= (mn != m[t]);
      m[t] = mn;
      const float p = exp2f(s[t] - mn);
      lsum[t] = fmaf(lsum[t], alpha[t], p);
      PS[kk * kH + 3 * hq + t] = p;
      if (kk == 0) ST[3 * hq + t] = alpha[t];
    }
    const bool rescale = __any_sync(0xffffffffu, moved);
    __syncwarp();

    // ---- values: oval += p[key, head] * V[key, head, :] ----
    if (rescale) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float a = ST[hd[u]];
        oval[u].x *= a; oval[u].y *= a; oval[u].z *= a; oval[u].w *= a;
      }
    }
    for (int k8 = 0; k8 < nk; ++k8) {
      const float4* vp = reinterpret_cast<const float4*>(kvp + k8 * kKVStride + kVOff) + lane;
      const float* pr = PS + k8 * kH;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (u < 3 || has3) {
          const float4 v = vp[32 * u];
          const float p = pr[hd[u]];
          oval[u].x = fmaf(p, v.x, oval[u].x); oval[u].y = fmaf(p, v.y, oval[u].y);
          oval[u].z = fmaf(p, v.z, oval[u].z); oval[u].w = fmaf(p, v.w, oval[u].w);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(kvempty + buf);       // this row is done with the chunk's keys / values
    if (warp == 0 && c + 2 < nchunks) {              // refill the buffer once every row has released it
      mbar_wait(kvempty + buf, (c >> 1) & 1);
      fence_proxy_async();
      issue_kv(c + 2);
    }

    // ---- pair row: acc[h] += p[key, h] * z[i, key, :] ----
    if (rescale) {
#pragma unroll
      for (int h = 0; h < kH; ++h) {
        const float a = ST[h];
        acc[h][0].x *= a; acc[h][0].y *= a; acc[h][1].x *= a; acc[h][1].y *= a;
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int q = 2 * c + half;
      if (q < nq) {
        const int slot = q % kZSlots, nz = min(kZKeys, N - q * kZKeys);
        mbar_wait(zfull + slot, (q / kZSlots) & 1);
        const float4* zs = reinterpret_cast<const float4*>(ZR + slot * kZSlotFloats) + lane;
        for (int k4 = 0; k4 < nz; ++k4) {
          const float4 zv = zs[k4 * (kCz / 4)];
          const float4* ap = reinterpret_cast<const float4*>(PS + (half * kZKeys + k4) * kH);
          float a[kH];
#pragma unroll
          for (int u = 0; u < kH / 4; ++u) { const float4 v = ap[u]; a[4 * u] = v.x; a[4 * u + 1] = v.y; a[4 * u + 2] = v.z; a[4 * u + 3] = v.w; }
          const float2 zlo = make_float2(zv.x, zv.y), zhi = make_float2(zv.z, zv.w);
#pragma unroll
          for (int h = 0; h < kH; ++h) {
            const float2 aa = make_float2(a[h], a[h]);
            acc[h][0] = ffma2(aa, zlo, acc[h][0]);
            acc[h][1] = ffma2(aa, zhi, acc[h][1]);
          }
        }
        __syncwarp();                                // every lane has read the slot
        if (lane == 0 && q + kZSlots < nq) { fence_proxy_async(); issue_z(q + kZSlots); }
      }
    }
    __syncwarp();                                    // PS / ST are rewritten by the next chunk
  }

  // ---- normalise and write the 2112-wide feature row of residue i ----
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    float l = lsum[t];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    l += __shfl_xor_sync(0xffffffffu, l, 4);
    if (kk == 0) ST[16 + 3 * hq + t] = 1.f / l;
  }
  __syncwarp();
  float* frow = feats + bn * kFeat;
#pragma unroll
  for (int h = 0; h < kH; ++h) {                     // 'b i h c -> b i (h c)'  folding.py:126-127
    const float inv = ST[16 + h];
    *reinterpret_cast<float4*>(frow + kFeatPair + h * kCz + 4 * lane) =
        make_float4(acc[h][0].x * inv, acc[h][0].y * inv, acc[h][1].x * inv, acc[h][1].y * inv);
  }
  float* OV = ZR;                                    // the z ring is idle now: stage the 480 value outputs
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (u < 3 || has3) {
      const float inv = ST[16 + hd[u]];
      reinterpret_cast<float4*>(OV)[lane + 32 * u] = make_float4(oval[u].x * inv, oval[u].y * inv, oval[u].z * inv, oval[u].w * inv);
    }
  }
  __syncwarp();
  for (int e = lane; e < kH * kSv; e += 32) frow[e] = OV[(e / kSv) * kVD + (e % kSv)];       // 'b i h c -> b i (h c)'  :115
  {
    float Rm[9], tr[3], it[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rm[k] = __ldg(rots + bn * 9 + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) tr[k] = __ldg(trans + bn * 3 + k);
    // invert_rigids (r3.py:54-59): R^T, -(R^T t); then rigids_apply  folding.py:121
#pragma unroll
    for (int k = 0; k < 3; ++k) it[k] = -(Rm[k] * tr[0] + Rm[3 + k] * tr[1] + Rm[6 + k] * tr[2]);
    for (int pi = lane; pi < kH * kPv; pi += 32) {
      const int h = pi / kPv, p = pi % kPv;
      const float* g = OV + h * kVD + kSv + 3 * p;
      float l[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) l[k] = it[k] + (Rm[k] * g[0] + Rm[3 + k] * g[1] + Rm[6 + k] * g[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) frow[kFeatPt + k * (kH * kPv) + pi] = l[k];                  // '(r n)'  folding.py:122
      frow[kFeatNorm + pi] = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2] + 1e-8f);           // :123
    }
  }
}

// Tile height: time ~ rounds * (R + 6.4) — R rows of z per tile plus the tile's key / value chunks (3264 B per key
// against 512 B of z per key and row), rounds = ceil(tiles / SMs).
static int choose_rows(int B, int N, int sms) {
  int best = 1;
  double best_cost = 1e30;
  for (int R = 1; R <= kMaxRows && R <= N; ++R) {
    const long tiles = (long)B * ceil_div(N, R);
    const long rounds = (tiles + sms - 1) / sms;
    const double cost = (double)rounds * (R + 6.4);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = R; }
  }
  return best;
}

static int fused_sm_count() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
  }();
  return n;
}

size_t ipa_fused_qp_floats(int B, int N) { return (size_t)B * N * kQRow; }
size_t ipa_fused_kvp_floats(int B, int N) { return (size_t)B * N * kKVRow; }

int launch_ipa_pack_nodes(cudaStream_t s, int B, int N, const float* proj, const float* rots, const float* trans,
                          float* Qp, float* KVp) {
  const cudaError_t le = launch_kernel(ipa_pack_nodes_kernel, dim3(ceil_div(B * N * kH * kPackItems, 256)), dim3(256), 0, s, B, N,
                                       proj, rots, trans, Qp, KVp);
  count_launch();
  if (le != cudaSuccess) { set_error("launch of ipa_pack_nodes_kernel failed: %s", cudaGetErrorString(le)); return ABX_ERR_CUDA; }
  return check_launch("ipa_pack_nodes_kernel");
}

int launch_ipa_fused(cudaStream_t s, int B, int N, const float* Qp, const float* KVp, const float* bias, const float* mask,
                     const float* rots, const float* trans, const float* point_weights, const float* z, float* feats) {
  const int R = choose_rows(B, N, fused_sm_count());
  const int tiles_per_b = ceil_div(N, R);
  const size_t smem = fused_smem_bytes(R);
  ABX_CUDA(cudaFuncSetAttribute(ipa_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes(kMaxRows)));
  const cudaError_t le = launch_kernel(ipa_fused_kernel, dim3(B * tiles_per_b), dim3(32 * R), smem, s, N, R, tiles_per_b, Qp, KVp,
                                       bias, mask, rots, trans, point_weights, z, feats);
  count_launch();
  if (le != cudaSuccess) { set_error("launch of ipa_fused_kernel failed: %s", cudaGetErrorString(le)); return ABX_ERR_CUDA; }
  return check_launch("ipa_fused_kernel");
}

}  // namespace abx
