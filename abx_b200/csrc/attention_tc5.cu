// Triangle-attention core on the 5th-generation tensor cores (reference: abx/model/seqformer.py:283-301 `Attention.forward`
// as called by `TriangleAttention` :506-550):
//
//   o[b,s,i,h,:] = softmax_j( q[b,s,i,h,:] . k[b,s,j,h,:] / sqrt(D) + bias[b,h,i,j], keys with mask 0 -> finfo.min ) v[b,s,j,h,:]
//                  (* sigmoid(gate[b,s,i,h,:]) when the output gate is fused)
//
// One CTA = 128 query rows of one (b, s, h) slice; it walks the keys in tiles of 64.  Both products run as 3xTF32
// tcgen05.mma (hi = the raw fp32 word, lo = x - hi: fp32-level accuracy) with fp32 accumulators in tensor memory:
//   S  = Q K^T    M = 128 queries, N = 64 keys, K = D        A = Q (shared memory, K-major), B = K tile (shared memory)
//   O' = P V      M = 128 queries, N = D,       K = 64 keys  A = P IN TENSOR MEMORY (written back by the softmax warps),
//                                                            B = V^T tile (shared memory, K-major = keys contiguous per channel)
// The logits never leave the SM: the softmax warps (thread = query row = tensor-memory lane) read S with tcgen05.ld, add the pair
// bias (prefetched into registers while the tile is staged), apply the key mask, run the online softmax in base 2 and write the
// probabilities (hi / lo) back to tensor memory; the per-tile product O' is drained into register accumulators (rescaled in
// registers, added with round-to-nearest: the tensor core's own accumulation truncates).  K / V tiles are split and laid out by the
// loader warps (no-swizzle K-major core matrices with a 16-byte skew per 16-byte column block: conflict-free stores).
// The phases of a tile are sequential inside a CTA; two CTAs per SM (113 KB of shared memory, 256 tensor-memory columns each)
// fill each other's bubbles.
#include <float.h>

#include "common.cuh"

namespace abx {

namespace {

constexpr int kTQ = 128, kTK = 64, kThreads = 256, kSoftThreads = 128;
constexpr int kRegsSoft = 184, kRegsLoad = 64;    // 2 CTAs/SM: 2 * 128 * (184 + 64) = 63488 registers
constexpr uint32_t kTmemCols = 256;               // S [0,64) | P hi [64,128) | P lo [128,192) | O' [192, 192 + D)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// K-major operand without swizzle: core matrices of 8 rows x 16 bytes (rows 16 B apart); 16-byte column blocks `lbo` bytes
// apart, 8-row groups 128 bytes apart.  Descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int N) {   // D f32, A/B tf32, K-major, M = 128
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

template <int D> struct Tc5Smem {
  static constexpr uint32_t kLboQ = kTQ * 16 + 16, kLboK = kTK * 16 + 16, kLboV = D * 16 + 16;   // skewed column-block strides
  static constexpr uint32_t kQBytes = (D / 4) * kLboQ, kKBytes = (D / 4) * kLboK, kVBytes = (kTK / 4) * kLboV;
  static constexpr uint32_t q_hi = 0, q_lo = q_hi + kQBytes, k_hi = q_lo + kQBytes, k_lo = k_hi + kKBytes, v_hi = k_lo + kKBytes,
                            v_lo = v_hi + kVBytes, mask = v_lo + kVBytes;
  __host__ __device__ static uint32_t total(int L) { return mask + (uint32_t)((L + kTK - 1) / kTK) * kTK * 4 + 64; }
};

}  // namespace

template <int D>
__global__ void __launch_bounds__(kThreads, 2) pair_attention_tc5_kernel(
    int L, int H, int S, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int ld,
    const float* __restrict__ bias, const float* __restrict__ key_mask, const float* __restrict__ gate, float scale,
    float* __restrict__ out) {
  using SM = Tc5Smem<D>;
  constexpr int D4 = D / 4;
  extern __shared__ __align__(128) uint8_t sm[];
  float* Ms = reinterpret_cast<float*>(sm + SM::mask);
  const int nkt = (L + kTK - 1) / kTK;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + SM::mask + (size_t)nkt * kTK * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int h = blockIdx.x, bs = blockIdx.y, b = bs / S, q0 = blockIdx.z * kTQ;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t row0 = (size_t)bs * L;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = threadIdx.x; j < nkt * kTK; j += kThreads) Ms[j] = (j < L) ? (key_mask ? __ldg(key_mask + (size_t)b * L + j) : 1.f) : 0.f;
  // Q tile, pre-multiplied by log2(e) / sqrt(D): hi / lo, K-major core matrices
  for (int idx = threadIdx.x; idx < kTQ * D4; idx += kThreads) {
    const int i = idx / D4, c4 = idx % D4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + i < L) x = *reinterpret_cast<const float4*>(q + (row0 + q0 + i) * (size_t)ld + h * D + 4 * c4);
    x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
    hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
    lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
    const uint32_t off = c4 * SM::kLboQ + i * 16;
    *reinterpret_cast<float4*>(sm + SM::q_hi + off) = hi;
    *reinterpret_cast<float4*>(sm + SM::q_lo + off) = lo;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================= softmax / accumulator warps: thread = query row = tensor-memory lane =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoft));
    const int i = q0 + threadIdx.x;                  // query row of this thread
    const bool row_ok = i < L;
    const float* brow = bias + (((size_t)b * H + h) * L + (row_ok ? i : L - 1)) * L;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * warp) << 16);
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    float m = -FLT_MAX, l = 0.f;
    uint32_t ph = 0;
    const bool vec_bias = (L % 4) == 0;              // 16-byte bias loads need 16-byte aligned rows
    for (int kt = 0; kt < nkt; ++kt) {
      const int j0 = kt * kTK;
      // pair bias of this row for the tile's 64 keys: in flight while the loader warps stage the tile and S is computed
      float bv[kTK];
      if (vec_bias && j0 + kTK <= L) {
#pragma unroll
        for (int u = 0; u < kTK / 4; ++u) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(brow + j0) + u);
          bv[4 * u] = t4.x; bv[4 * u + 1] = t4.y; bv[4 * u + 2] = t4.z; bv[4 * u + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < kTK; ++u) bv[u] = (j0 + u < L) ? __ldg(brow + j0 + u) : 0.f;
      }
      __syncthreads();                               // (A) tile staged
      mbar_wait(bar, ph); ph ^= 1;                   // S = Q K^T complete
      tc_fence_after();
      float cm = -FLT_MAX;
      float sv[kTK];
#pragma unroll
      for (int c0 = 0; c0 < kTK; c0 += 16) {
        float t16[16];
        tmem_ld16(lane_base + c0, t16);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = j0 + c0 + u;
          float a = fmaf(bv[c0 + u], kLog2e, t16[u]);                   // base-2 logits
          a = (j < L) ? ((Ms[j] != 0.f) ? a : -FLT_MAX) : -INFINITY;    // masked_fill(finfo.min); padding keys contribute exactly 0
          sv[c0 + u] = a;
          cm = fmaxf(cm, a);
        }
      }
      const float mn = fmaxf(m, cm);
      const float alpha = exp2f(m - mn);
      m = mn;
      float psum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < kTK; c0 += 16) {
        uint32_t phi[16], plo[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float p = exp2f(sv[c0 + u] - mn);
          psum += p;
          phi[u] = __float_as_uint(p) & 0xffffe000u;
          plo[u] = __float_as_uint(p - __uint_as_float(phi[u]));
        }
        tmem_st16(lane_base + 64 + c0, phi);
        tmem_st16(lane_base + 128 + c0, plo);
      }
      l = fmaf(l, alpha, psum);
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] *= alpha;
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncthreads();                               // (B) P in tensor memory
      mbar_wait(bar, ph); ph ^= 1;                   // O' = P V complete
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 16) {
        float t16[16];
        tmem_ld16(lane_base + 192 + c0, t16);
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[c0 + u] += t16[u];
      }
      tc_fence_before();
      __syncthreads();                               // (C) tile consumed: operands and accumulators may be overwritten
    }
    if (row_ok) {
      const float inv = 1.f / l;
      const size_t HD = (size_t)H * D;
      float* orow = out + (row0 + i) * HD + h * D;
      const float* grow = gate ? gate + (row0 + i) * (size_t)ld + h * D : nullptr;
#pragma unroll
      for (int d = 0; d < D; d += 4) {
        float4 o = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
        if (grow) {                                  // out = sigmoid(gate) * attention  (seqformer.py:296-299)
          const float4 g4 = *reinterpret_cast<const float4*>(grow + d);
          o.x *= 1.f / (1.f + expf(-g4.x)); o.y *= 1.f / (1.f + expf(-g4.y));
          o.z *= 1.f / (1.f + expf(-g4.z)); o.w *= 1.f / (1.f + expf(-g4.w));
        }
        *reinterpret_cast<float4*>(orow + d) = o;
      }
    }
  } else {
    // ================= loader warps (+ the MMA issuer: warp 4) =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLoad));
    const int t = threadIdx.x - kSoftThreads;        // 0..127
    constexpr int kPer = kTK * D4 / kSoftThreads;    // float4 per thread and matrix (D = 48: 6)
    static_assert(kTK * D4 % kSoftThreads == 0, "tile must divide over the loader threads");
    float4 kreg[kPer], vreg[kPer];
    // K: thread <-> (key, 16-byte block) with the block index fastest: coalesced global reads, conflict-free 16-byte stores.
    // V: thread <-> key (fastest) x a run of kPer consecutive 16-byte blocks of that key's row: the transposing scalar stores of
    // a warp then go to 32 consecutive keys of one channel row = 32 distinct banks (each thread still reads whole 32-byte sectors).
    const int vkey = t % kTK, vc0 = (t / kTK) * kPer;
    static_assert(kSoftThreads / kTK * kPer == D4, "V mapping must cover the row");
    auto fetch = [&](int kt) {                       // global -> registers (the next tile's loads are in flight during the math)
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int idx = t + u * kSoftThreads, j = kt * kTK + idx / D4, c4 = idx % D4;
        kreg[u] = vreg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < L) kreg[u] = *reinterpret_cast<const float4*>(k + (row0 + j) * (size_t)ld + h * D + 4 * c4);
        const int jv = kt * kTK + vkey;
        if (jv < L) vreg[u] = *reinterpret_cast<const float4*>(v + (row0 + jv) * (size_t)ld + h * D + 4 * (vc0 + u));
      }
    };
    auto split4 = [](const float4& x, float4& hi, float4& lo) {
      hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
      lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
    };
    auto stage = [&]() {                             // registers -> hi / lo operand tiles
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int idx = t + u * kSoftThreads, jj = idx / D4, c4 = idx % D4;
        float4 hi, lo;
        split4(kreg[u], hi, lo);                     // K tile: rows = keys, 16-byte column block c4
        const uint32_t ko = c4 * SM::kLboK + jj * 16;
        *reinterpret_cast<float4*>(sm + SM::k_hi + ko) = hi;
        *reinterpret_cast<float4*>(sm + SM::k_lo + ko) = lo;
        split4(vreg[u], hi, lo);                     // V^T tile: rows = channels 4 c .. 4 c + 3, column block vkey / 4, element vkey % 4
        const uint32_t vo = (vkey >> 2) * SM::kLboV + (4 * (vc0 + u)) * 16 + (vkey & 3) * 4;
        *reinterpret_cast<float*>(sm + SM::v_hi + vo) = hi.x; *reinterpret_cast<float*>(sm + SM::v_hi + vo + 16) = hi.y;
        *reinterpret_cast<float*>(sm + SM::v_hi + vo + 32) = hi.z; *reinterpret_cast<float*>(sm + SM::v_hi + vo + 48) = hi.w;
        *reinterpret_cast<float*>(sm + SM::v_lo + vo) = lo.x; *reinterpret_cast<float*>(sm + SM::v_lo + vo + 16) = lo.y;
        *reinterpret_cast<float*>(sm + SM::v_lo + vo + 32) = lo.z; *reinterpret_cast<float*>(sm + SM::v_lo + vo + 48) = lo.w;
      }
    };
    const uint32_t smb = smem_u32(sm);
    constexpr uint32_t idS = idesc_tf32(kTK), idO = idesc_tf32(D);
    fetch(0);
    for (int kt = 0; kt < nkt; ++kt) {
      stage();
      fence_proxy_async();
      __syncthreads();                               // (A)
      if (warp == 4) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 8; ++ks) {       // S = Q K^T: small cross terms first, then hi * hi
            const uint32_t qa = 2 * ks * SM::kLboQ, ka = 2 * ks * SM::kLboK;
            umma_ss(tmem_base, umma_desc(smb + SM::q_hi + qa, SM::kLboQ), umma_desc(smb + SM::k_lo + ka, SM::kLboK), idS, ks > 0 ? 1u : 0u);
            umma_ss(tmem_base, umma_desc(smb + SM::q_lo + qa, SM::kLboQ), umma_desc(smb + SM::k_hi + ka, SM::kLboK), idS, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < D / 8; ++ks) {
            const uint32_t qa = 2 * ks * SM::kLboQ, ka = 2 * ks * SM::kLboK;
            umma_ss(tmem_base, umma_desc(smb + SM::q_hi + qa, SM::kLboQ), umma_desc(smb + SM::k_hi + ka, SM::kLboK), idS, 1u);
          }
          umma_commit(bar);
        }
        __syncwarp();
      }
      if (kt + 1 < nkt) fetch(kt + 1);
      __syncthreads();                               // (B)
      if (warp == 4) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kTK / 8; ++ks) {     // O' = P V over the tile's 64 keys
            const uint32_t va = 2 * ks * SM::kLboV;
            umma_ts(tmem_base + 192, tmem_base + 64 + 8 * ks, umma_desc(smb + SM::v_lo + va, SM::kLboV), idO, ks > 0 ? 1u : 0u);
            umma_ts(tmem_base + 192, tmem_base + 128 + 8 * ks, umma_desc(smb + SM::v_hi + va, SM::kLboV), idO, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < kTK / 8; ++ks)
            umma_ts(tmem_base + 192, tmem_base + 64 + 8 * ks, umma_desc(smb + SM::v_hi + 2 * ks * SM::kLboV, SM::kLboV), idO, 1u);
          umma_commit(bar);
        }
        __syncwarp();
      }
      __syncthreads();                               // (C)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <int D>
int launch_attention_tc5(cudaStream_t st, int B, int S, int L, int H, const float* q, const float* k, const float* v, int ld,
                         const float* bias, const float* key_mask, const float* gate, float* out) {
  const size_t smem = Tc5Smem<D>::total(L);
  ABX_REQUIRE(smem <= 113 * 1024, "abx_pair_attention: L=%d with head dim %d needs %zu bytes of shared memory per CTA", L, D, smem);
  ABX_CUDA(cudaFuncSetAttribute(pair_attention_tc5_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_attention_tc5_kernel<D><<<dim3(H, B * S, (L + kTQ - 1) / kTQ), kThreads, smem, st>>>(
      L, H, S, q, k, v, ld, bias, key_mask, gate, kLog2e / sqrtf((float)D), out);
  count_launch();
  return check_launch("pair_attention_tc5_kernel");
}

template int launch_attention_tc5<16>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, const float*, float*);
template int launch_attention_tc5<32>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, const float*, float*);
template int launch_attention_tc5<48>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, const float*, float*);
template int launch_attention_tc5<64>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, const float*, float*);

}  // namespace abx
