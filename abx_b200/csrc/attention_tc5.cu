// Triangle-attention core on the 5th-generation tensor cores (reference: abx/model/seqformer.py:283-301 `Attention.forward`
// as called by `TriangleAttention` :506-550):
//
//   o[b,s,i,h,:] = softmax_j( q[b,s,i,h,:] . k[b,s,j,h,:] / sqrt(D) + bias[b,h,i,j], keys with mask 0 -> finfo.min ) v[b,s,j,h,:]
//                  (* sigmoid(gate[b,s,i,h,:]) when the output gate is fused)
//
// One CTA = up to two 128-row query tiles ("slots") of one (b, s, h) slice, walking the keys in tiles of 64; the key / value
// tiles are staged once for both slots.  Both products run as 3xTF32 tcgen05.mma (hi = x & ~0x1fff, lo = x - hi: fp32-level
// accuracy) with fp32 accumulators in tensor memory:
//   S  = Q K^T    M = 128 queries, N = 64 keys, K = D        A = Q (shared memory, K-major), B = K tile (shared memory)
//   O' = P V      M = 128 queries, N = D,       K = 64 keys  A = P IN TENSOR MEMORY (written back by the softmax warps),
//                                                            B = V^T tile (shared memory, K-major = keys contiguous per channel)
// Warp-specialised, one CTA per SM, 16 warps (registers re-balanced with setmaxnreg):
//   warps 0-3 / 4-7   softmax warpgroup of slot 0 / 1: thread = query row = tensor-memory lane.  Per key tile: add the previous
//                     tile's O' to the register accumulators (round-to-nearest adds: the tensor core's own accumulation
//                     truncates), read S with tcgen05.ld, add the pair bias (prefetched two 16-key chunks ahead), apply the key
//                     mask, online softmax in base 2, write P (hi / lo) back to tensor memory, arrive `p_full`
//   warps 8-11        loaders: K / V tile -> registers -> hi / lo operand tiles (no-swizzle K-major core matrices with a 16-byte
//                     skew per 16-byte column block: conflict-free stores), two stages
//   warp 12           MMA issuer.  Issue order  PV_0(kt), S_0(kt+1), PV_1(kt), S_1(kt+1), ...: while one slot's warpgroup runs
//                     its softmax, the tensor core works on the other slot's products (ping-pong); `p_full` of a slot implies
//                     that its S and O' columns have been read, so no further hand-shake is needed
// The logits never leave the SM.
#include <float.h>

#include "common.cuh"

namespace abx {

namespace {

constexpr int kTQ = 128, kTK = 64, kThreads = 512, kSoftThreads = 128, kLoadThreads = 128, kSlots = 2, kStagesKV = 2;
constexpr int kRegsSoft = 192, kRegsLoad = 80, kRegsCtl = 40;   // 256 * 192 + 128 * (80 + 40) = 65536 - 1024 registers
constexpr uint32_t kSlotCols = 240;               // per query-tile slot: S [0,64) | P hi [64,128) | P lo [128,192) | O' [192, 192 + D)
constexpr uint32_t kTmemCols = 512;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// A wait that cannot hang the device: a synchronisation bug ends the kernel with a trap (launch error) after ~1 s.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0, tries = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (!ok && ++tries > (1u << 24)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 2^x for x <= 0 (softmax numerators): the bare MUFU.EX2 — results below 2^-126 flush to zero, -inf gives 0
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {   // the caller issues tmem_ld_wait() before reading v
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}


template <int D> struct Tc5Smem {
  static constexpr uint32_t kLboQ = kTQ * 16 + 16, kLboK = kTK * 16 + 16, kLboV = D * 16 + 16;   // skewed column-block strides
  static constexpr uint32_t kQBytes = (D / 4) * kLboQ, kKBytes = (D / 4) * kLboK, kVBytes = (kTK / 4) * kLboV;
  static constexpr uint32_t kSlotBytes = 2 * kQBytes;                          // Q hi | Q lo
  static constexpr uint32_t kStageBytes = 2 * kKBytes + 2 * kVBytes;           // K hi | K lo | V^T hi | V^T lo
  static constexpr uint32_t q_base = 0, kv_base = kSlots * kSlotBytes, mask = kv_base + kStagesKV * kStageBytes;
  __host__ __device__ static uint32_t total(int L) {                           // + key mask, tile flags, 10 mbarriers, TMEM slot
    const uint32_t nkt = (uint32_t)((L + kTK - 1) / kTK);
    return mask + nkt * kTK * 4 + ((nkt + 15) / 16) * 16 + 10 * 8 + 16;
  }
};

}  // namespace

template <int D>
__global__ void __launch_bounds__(kThreads, 1) pair_attention_tc5_kernel(
    int L, int H, int S, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int ld,
    const float* __restrict__ bias, const float* __restrict__ key_mask, const float* __restrict__ gate, float scale,
    float* __restrict__ out) {
  using SM = Tc5Smem<D>;
  constexpr int D4 = D / 4;
  extern __shared__ __align__(128) uint8_t sm[];
  const int nkt = (L + kTK - 1) / kTK, nqt = (L + kTQ - 1) / kTQ;
  float* Ms = reinterpret_cast<float*>(sm + SM::mask);
  uint8_t* tile_plain = sm + SM::mask + (size_t)nkt * kTK * 4;     // 1: every key of the tile is inside L and unmasked
  uint64_t* bars = reinterpret_cast<uint64_t*>(tile_plain + ((nkt + 15) / 16) * 16);
  uint64_t* kv_full = bars;            // [2] loaders -> issuer
  uint64_t* kv_empty = bars + 2;       // [2] issuer (commit) -> loaders
  uint64_t* s_full = bars + 4;         // [2 slots] issuer (commit) -> softmax warpgroup
  uint64_t* p_full = bars + 6;         // [2 slots] softmax warpgroup -> issuer
  uint64_t* o_full = bars + 8;         // [2 slots] issuer (commit) -> softmax warpgroup
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int h = blockIdx.x, bs = blockIdx.y, b = bs / S, q0 = blockIdx.z * (kSlots * kTQ);
  const int nslots = min(kSlots, nqt - (int)blockIdx.z * kSlots);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t row0 = (size_t)bs * L;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(kv_full + i, kLoadThreads);
      mbar_init(kv_empty + i, 1);
      mbar_init(s_full + i, 1);
      mbar_init(p_full + i, kSoftThreads);
      mbar_init(o_full + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = threadIdx.x; j < nkt * kTK; j += kThreads) Ms[j] = (j < L) ? (key_mask ? __ldg(key_mask + (size_t)b * L + j) : 1.f) : 0.f;
  // Q tiles of the slots, pre-multiplied by log2(e) / sqrt(D): hi / lo, K-major core matrices
  for (int idx = threadIdx.x; idx < nslots * kTQ * D4; idx += kThreads) {
    const int i = idx / D4, c4 = idx % D4, slot = i / kTQ, il = i % kTQ;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + i < L) x = *reinterpret_cast<const float4*>(q + (row0 + q0 + i) * (size_t)ld + h * D + 4 * c4);
    x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
    hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
    lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
    const uint32_t off = SM::q_base + slot * SM::kSlotBytes + c4 * SM::kLboQ + il * 16;
    *reinterpret_cast<float4*>(sm + off) = hi;
    *reinterpret_cast<float4*>(sm + off + SM::kQBytes) = lo;
  }
  __syncthreads();
  if (threadIdx.x < nkt) {
    bool plain = (threadIdx.x + 1) * kTK <= L;
    for (int u = 0; u < kTK && plain; ++u) plain = Ms[threadIdx.x * kTK + u] != 0.f;
    tile_plain[threadIdx.x] = plain ? 1 : 0;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ================= softmax / accumulator warpgroups: thread = query row = tensor-memory lane =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoft));
    const int slot = warp >> 2;
    if (slot < nslots) {
      const int i = q0 + slot * kTQ + (threadIdx.x & (kSoftThreads - 1));   // query row of this thread
      const bool row_ok = i < L;
      const float* brow = bias + (((size_t)b * H + h) * L + (row_ok ? i : L - 1)) * L;
      const uint32_t lane_base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + slot * kSlotCols;
      const bool vec_bias = (L % 4) == 0;              // 16-byte bias loads need 16-byte aligned rows
      auto load_bias = [&](int j, float (&dst)[16]) {  // pair bias of this row for keys j .. j + 15 (0 beyond L)
        if (vec_bias && j + 16 <= L) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(brow + j) + u);
            dst[4 * u] = t4.x; dst[4 * u + 1] = t4.y; dst[4 * u + 2] = t4.z; dst[4 * u + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 16; ++u) dst[u] = (j + u < L) ? __ldg(brow + j + u) : 0.f;
        }
      };
      float acc[D];
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = 0.f;
      float m = -FLT_MAX, l = 0.f;
      float bv0[16], bv1[16], bv2[16], bv3[16];
      load_bias(0, bv0);
      load_bias(16, bv1);
      for (int kt = 0; kt < nkt; ++kt) {
        const int j0 = kt * kTK;
        const uint32_t par = kt & 1;
        if (kt > 0) {                                  // O' of the previous tile (relative to the same reference point as acc)
          mbar_wait(o_full + slot, par ^ 1);
          tc_fence_after();
          float t[D];
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 16) tmem_ld16_nowait(lane_base + 192 + c0, t + c0);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < D; ++d) acc[d] += t[d];
        }
        load_bias(j0 + 32, bv2);
        load_bias(j0 + 48, bv3);
        mbar_wait(s_full + slot, par);                 // S = Q K^T of this tile
        tc_fence_after();
        float sv[kTK];
#pragma unroll
        for (int c0 = 0; c0 < kTK; c0 += 16) tmem_ld16_nowait(lane_base + c0, sv + c0);
        tmem_ld_wait();
        const bool plain = tile_plain[kt] != 0;
        float cm = -FLT_MAX;
        auto chunk = [&](int c0, const float (&bv)[16]) {   // base-2 logits of keys j0 + c0 .. + 15
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float a = fmaf(bv[u], kLog2e, sv[c0 + u]);
            if (!plain) {                                   // masked_fill(finfo.min); padding keys contribute exactly 0
              const int j = j0 + c0 + u;
              a = (j < L) ? ((Ms[j] != 0.f) ? a : -FLT_MAX) : -INFINITY;
            }
            sv[c0 + u] = a;
            cm = fmaxf(cm, a);
          }
        };
        chunk(0, bv0);
        chunk(16, bv1);
        if (kt + 1 < nkt) {                            // first half of the next tile's bias: in flight during the rest of this tile
          load_bias(j0 + kTK, bv0);
          load_bias(j0 + kTK + 16, bv1);
        }
        chunk(32, bv2);
        chunk(48, bv3);
        const float mn = fmaxf(m, cm);
        const float alpha = ex2_ftz(m - mn);
        m = mn;
        float psum = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < kTK; c0 += 16) {
          uint32_t phi[16], plo[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float p = ex2_ftz(sv[c0 + u] - mn);
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
        mbar_arrive(p_full + slot);                    // P written; S and O' of this slot have been read
      }
      {
        mbar_wait(o_full + slot, (nkt - 1) & 1);
        tc_fence_after();
        float t[D];
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 16) tmem_ld16_nowait(lane_base + 192 + c0, t + c0);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] += t[d];
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
    }
  } else if (warp < 12) {
    // ================= loader warps: K / V tile -> hi / lo operand tiles, two stages =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLoad));
    const int t = threadIdx.x - 2 * kSoftThreads;    // 0..127
    constexpr int kPer = kTK * D4 / kLoadThreads;    // float4 per thread and matrix (D = 48: 6)
    static_assert(kTK * D4 % kLoadThreads == 0, "tile must divide over the loader threads");
    float4 kreg[kPer], vreg[kPer];
    // K: thread <-> (key, 16-byte block) with the block index fastest: coalesced global reads, conflict-free 16-byte stores.
    // V: thread <-> key (fastest) x a run of kPer consecutive 16-byte blocks of that key's row: the transposing scalar stores of
    // a warp then go to 32 consecutive keys of one channel row = 32 distinct banks (each thread still reads whole 32-byte sectors).
    const int vkey = t % kTK, vc0 = (t / kTK) * kPer;
    static_assert(kLoadThreads / kTK * kPer == D4, "V mapping must cover the row");
    auto fetch = [&](int kt) {                       // global -> registers (the next tile's loads are in flight during the math)
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int idx = t + u * kLoadThreads, j = kt * kTK + idx / D4, c4 = idx % D4;
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
    auto stage = [&](uint8_t* st) {                  // registers -> hi / lo operand tiles of one stage
      uint8_t* k_hi = st, *k_lo = st + SM::kKBytes, *v_hi = st + 2 * SM::kKBytes, *v_lo = v_hi + SM::kVBytes;
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int idx = t + u * kLoadThreads, jj = idx / D4, c4 = idx % D4;
        float4 hi, lo;
        split4(kreg[u], hi, lo);                     // K tile: rows = keys, 16-byte column block c4
        const uint32_t ko = c4 * SM::kLboK + jj * 16;
        *reinterpret_cast<float4*>(k_hi + ko) = hi;
        *reinterpret_cast<float4*>(k_lo + ko) = lo;
        split4(vreg[u], hi, lo);                     // V^T tile: rows = channels 4 c .. 4 c + 3, column block vkey / 4, element vkey % 4
        const uint32_t vo = (vkey >> 2) * SM::kLboV + (4 * (vc0 + u)) * 16 + (vkey & 3) * 4;
        *reinterpret_cast<float*>(v_hi + vo) = hi.x; *reinterpret_cast<float*>(v_hi + vo + 16) = hi.y;
        *reinterpret_cast<float*>(v_hi + vo + 32) = hi.z; *reinterpret_cast<float*>(v_hi + vo + 48) = hi.w;
        *reinterpret_cast<float*>(v_lo + vo) = lo.x; *reinterpret_cast<float*>(v_lo + vo + 16) = lo.y;
        *reinterpret_cast<float*>(v_lo + vo + 32) = lo.z; *reinterpret_cast<float*>(v_lo + vo + 48) = lo.w;
      }
    };
    fetch(0);
    for (int kt = 0; kt < nkt; ++kt) {
      const int st = kt & 1;
      if (kt >= kStagesKV) mbar_wait(kv_empty + st, ((kt >> 1) - 1) & 1);   // the products of tile kt - 2 have read the stage
      stage(sm + SM::kv_base + st * SM::kStageBytes);
      fence_proxy_async();
      mbar_arrive(kv_full + st);
      if (kt + 1 < nkt) fetch(kt + 1);
    }
  } else {
    // ================= MMA issuer (warp 12); warps 13-15 only hold the tensor-memory allocation =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
    if (warp == 12) {
      const uint32_t smb = smem_u32(sm);
      constexpr uint32_t idS = idesc_tf32(kTK), idO = idesc_tf32(D);
      auto issue_s = [&](int slot, int st) {           // S = Q K^T: small cross terms first, then hi * hi
        const uint32_t qh = smb + SM::q_base + slot * SM::kSlotBytes, ql = qh + SM::kQBytes;
        const uint32_t kh = smb + SM::kv_base + st * SM::kStageBytes, kl = kh + SM::kKBytes;
        const uint32_t d = tmem_base + slot * kSlotCols;
#pragma unroll
        for (int ks = 0; ks < D / 8; ++ks) {
          const uint32_t qa = 2 * ks * SM::kLboQ, ka = 2 * ks * SM::kLboK;
          umma_ss(d, umma_desc(qh + qa, SM::kLboQ), umma_desc(kl + ka, SM::kLboK), idS, ks > 0 ? 1u : 0u);
          umma_ss(d, umma_desc(ql + qa, SM::kLboQ), umma_desc(kh + ka, SM::kLboK), idS, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < D / 8; ++ks) {
          const uint32_t qa = 2 * ks * SM::kLboQ, ka = 2 * ks * SM::kLboK;
          umma_ss(d, umma_desc(qh + qa, SM::kLboQ), umma_desc(kh + ka, SM::kLboK), idS, 1u);
        }
        umma_commit(s_full + slot);
      };
      auto issue_pv = [&](int slot, int st) {          // O' = P V over the tile's 64 keys
        const uint32_t vh = smb + SM::kv_base + st * SM::kStageBytes + 2 * SM::kKBytes, vl = vh + SM::kVBytes;
        const uint32_t tb = tmem_base + slot * kSlotCols;
#pragma unroll
        for (int ks = 0; ks < kTK / 8; ++ks) {
          const uint32_t va = 2 * ks * SM::kLboV;
          umma_ts(tb + 192, tb + 64 + 8 * ks, umma_desc(vl + va, SM::kLboV), idO, ks > 0 ? 1u : 0u);
          umma_ts(tb + 192, tb + 128 + 8 * ks, umma_desc(vh + va, SM::kLboV), idO, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < kTK / 8; ++ks)
          umma_ts(tb + 192, tb + 64 + 8 * ks, umma_desc(vh + 2 * ks * SM::kLboV, SM::kLboV), idO, 1u);
        umma_commit(o_full + slot);
      };
      if (elect_one()) {
        mbar_wait(kv_full + 0, 0);
        tc_fence_after();
        for (int slot = 0; slot < nslots; ++slot) issue_s(slot, 0);
        for (int kt = 0; kt < nkt; ++kt) {
          const int st = kt & 1;
          for (int slot = 0; slot < nslots; ++slot) {
            mbar_wait(p_full + slot, kt & 1);
            tc_fence_after();
            issue_pv(slot, st);
            if (slot == nslots - 1) umma_commit(kv_empty + st);     // every product reading stage st has been issued
            if (kt + 1 < nkt) {
              if (slot == 0) {
                mbar_wait(kv_full + (st ^ 1), ((kt + 1) >> 1) & 1);
                tc_fence_after();
              }
              issue_s(slot, st ^ 1);
            }
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <int D>
int launch_attention_tc5(cudaStream_t st, int B, int S, int L, int H, const float* q, const float* k, const float* v, int ld,
                         const float* bias, const float* key_mask, const float* gate, float* out) {
  const size_t smem = Tc5Smem<D>::total(L);
  ABX_REQUIRE(smem <= 227 * 1024, "abx_pair_attention: L=%d with head dim %d needs %zu bytes of shared memory per CTA", L, D, smem);
  ABX_CUDA(cudaFuncSetAttribute(pair_attention_tc5_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nqt = (L + kTQ - 1) / kTQ;
  pair_attention_tc5_kernel<D><<<dim3(H, B * S, (nqt + kSlots - 1) / kSlots), kThreads, smem, st>>>(
      L, H, S, q, k, v, ld, bias, key_mask, gate, kLog2e / sqrtf((float)D), out);
  count_launch();
  return check_launch("pair_attention_tc5_kernel");
}

// shared memory the kernel needs for key length L (the dispatcher falls back to the mma.sync kernel beyond 227 KB: D = 64)
template <int D> size_t attention_tc5_smem(int L) { return Tc5Smem<D>::total(L); }
template size_t attention_tc5_smem<16>(int);
template size_t attention_tc5_smem<32>(int);
template size_t attention_tc5_smem<48>(int);

template int launch_attention_tc5<16>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, const float*, float*);
template int launch_attention_tc5<32>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, const float*, float*);
template int launch_attention_tc5<48>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, const float*, float*);

}  // namespace abx
