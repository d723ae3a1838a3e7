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
// The logits never leave the SM.  The pair bias arrives PRE-TILED (abx_pair_attention_tc5 in include/abx_b200.h):
//   bias_tiles[b][h][kt][it][j][i] = log2(e) * bias[b,h, 32 it + i, 64 kt + j]      (32 query rows x 64 keys per block),
// with -FLT_MAX at masked keys (the reference's masked_fill(finfo.min): a softmax over masked keys only stays uniform) and -inf
// at padding keys beyond L, so the kernel needs no mask logic and a warp's load of one key is a single 128-byte line at an
// immediate offset (row-per-lane loads of the reference layout — rows of 1400 bytes at L = 350 — kept the LSU busy ~14k clk per
// key tile).
#include <float.h>

#include "common.cuh"

namespace abx {

namespace {

constexpr int kTQ = 128, kTK = 64, kThreads = 512, kSoftThreads = 128, kLoadThreads = 128, kSlots = 2, kStagesKV = 2;
constexpr int kRegsSoft = 184, kRegsLoad = 80, kRegsCtl = 56;   // 256 * 184 + 128 * (80 + 56) = 65536 - 1024 registers
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


// Per-role wait / work cycle counters of CTA (0,0,0), compiled in with -DABX_ATTN_PROFILE=1 (read with abx_attention_profile):
// [0] softmax WG0 wait O', [1] wait S, [2] total; [4] loader wait kv_empty, [5] total; [8] issuer wait kv_full, [9] wait P,
// [10] issue (blocking tcgen05.mma / commit issue), [11] total; [12] key tiles.
#ifndef ABX_ATTN_PROFILE
#define ABX_ATTN_PROFILE 0
#endif
__device__ unsigned long long g_attn_prof[32];   // [16..21]: softmax WG0 phases: O' drain, S load, logits + max, exp + P store, store wait + arrive, bias issue
#define ABX_ATTN_PWAIT(slot, bar, par)                                                  \
  do {                                                                                  \
    if (prof) { const long long t0__ = clock64(); mbar_wait(bar, par); pw[slot] += clock64() - t0__; } \
    else mbar_wait(bar, par);                                                           \
  } while (0)

template <int D> struct Tc5Smem {
  // Column-block strides.  Q and K are NOT skewed: their 8-row x 16-byte core matrices must stay 128-byte aligned — a core
  // matrix that straddles two 128-byte rows of shared memory costs the SS-mode MMA two wavefronts instead of one (measured:
  // S = Q K^T at ~96 clk per MMA instead of 48).  V^T keeps a 16-byte skew (conflict-free transposing stores; its 12 core
  // matrices per MMA stay below the MMA's own time even at two wavefronts each).
  static constexpr uint32_t kLboQ = kTQ * 16, kLboK = kTK * 16, kLboV = D * 16 + 16;
  static constexpr uint32_t kQBytes = (D / 4) * kLboQ, kKBytes = (D / 4) * kLboK, kVBytes = (kTK / 4) * kLboV;
  static constexpr uint32_t kSlotBytes = 2 * kQBytes;                          // Q hi | Q lo
  static constexpr uint32_t kStageBytes = 2 * kKBytes + 2 * kVBytes;           // K hi | K lo | V^T hi | V^T lo
  static constexpr uint32_t q_base = 0, kv_base = kSlots * kSlotBytes, mask = kv_base + kStagesKV * kStageBytes;
  __host__ __device__ static uint32_t total(int) { return mask + 10 * 8 + 16; }   // + 10 mbarriers, TMEM slot
};

}  // namespace

template <int D>
__global__ void __launch_bounds__(kThreads, 1) pair_attention_tc5_kernel(
    int L, int H, int S, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int ld,
    const float* __restrict__ bias_tiles, const float* __restrict__ gate, float scale, float* __restrict__ out) {
  using SM = Tc5Smem<D>;
  constexpr int D4 = D / 4;
  extern __shared__ __align__(128) uint8_t sm[];
  const long long tentry = clock64();
  const int nkt = (L + kTK - 1) / kTK, nqt = (L + kTQ - 1) / kTQ;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM::mask);
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
  // Loader state lives at function scope so that the loader warps can issue the first K / V tile's loads BEFORE the Q staging
  // (two HBM round trips of the prologue overlap instead of following each other).
  const int t = threadIdx.x - 2 * kSoftThreads;      // 0..127 in the loader warps
  constexpr int kPer = kTK * D4 / kLoadThreads;      // float4 per thread and matrix (D = 48: 6)
  static_assert(kTK * D4 % kLoadThreads == 0, "tile must divide over the loader threads");
  float4 kreg[kPer], vreg[kPer];
  // K and V: thread <-> key (fastest) x a run of kPer consecutive 16-byte blocks of that key's row (whole 32-byte sectors per
  // thread).  K stores: a warp writes one block of 32 consecutive keys = 512 contiguous bytes; V stores (transposing, scalar):
  // 32 consecutive keys of one channel row = 32 distinct banks.
  const int vkey = t & (kTK - 1), vc0 = (t / kTK) * kPer;
  static_assert(kLoadThreads / kTK * kPer == D4, "mapping must cover the row");
  auto fetch = [&](int kt) {                         // global -> registers (the next tile's loads are in flight during the math)
    const int jv = kt * kTK + vkey;
    const size_t ro = (row0 + (jv < L ? jv : 0)) * (size_t)ld + h * D + 4 * vc0;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      kreg[u] = vreg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (jv < L) {
        kreg[u] = *reinterpret_cast<const float4*>(k + ro + 4 * u);
        vreg[u] = *reinterpret_cast<const float4*>(v + ro + 4 * u);
      }
    }
  };
  if (warp >= 8 && warp < 12) fetch(0);
  // Q tiles of the slots, pre-multiplied by log2(e) / sqrt(D): hi / lo, K-major core matrices.  Thread <-> (row, half of the
  // row's 16-byte blocks) with the row fastest: a warp's stores of one block are 32 consecutive rows = 512 contiguous bytes.
  {
    constexpr int kRun = D4 / 2;
    const int rows = nslots * kTQ;
    for (int item = threadIdx.x; item < 2 * rows; item += kThreads) {
      const int i = item % rows, run = item / rows, slot = i / kTQ, il = i % kTQ;
      const bool ok = q0 + i < L;
      const float* src = q + (row0 + q0 + (ok ? i : 0)) * (size_t)ld + h * D + 4 * run * kRun;
      float4 x[kRun];
#pragma unroll
      for (int u = 0; u < kRun; ++u) x[u] = ok ? *reinterpret_cast<const float4*>(src + 4 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kRun; ++u) {
        float4 v = x[u], hi, lo;
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
        const uint32_t off = SM::q_base + slot * SM::kSlotBytes + (run * kRun + u) * SM::kLboQ + il * 16;
        *reinterpret_cast<float4*>(sm + off) = hi;
        *reinterpret_cast<float4*>(sm + off + SM::kQBytes) = lo;
      }
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool prof = ABX_ATTN_PROFILE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  unsigned long long pw[3] = {0, 0, 0};
  const long long tstart = clock64();
  if (prof && threadIdx.x == 0) g_attn_prof[13] = tstart - tentry;     // prologue: barriers, TMEM allocation, mask, Q staging

  if (warp < 8) {
    // ================= softmax / accumulator warpgroups: thread = query row = tensor-memory lane =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoft));
    const int slot = warp >> 2;
    if (slot < nslots) {
      const int i = q0 + slot * kTQ + (threadIdx.x & (kSoftThreads - 1));   // query row of this thread
      const bool row_ok = i < L;
      const int nit = (L + 31) / 32;                   // 32-row blocks of the tiled bias; rows beyond them read the last block
      const size_t tile_stride = (size_t)nit * (kTK * 32);
      const float* bt = bias_tiles + ((size_t)(b * H + h) * nkt * nit + min(i >> 5, nit - 1)) * (kTK * 32) + lane;
      const uint32_t lane_base = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + slot * kSlotCols;
      auto load_bias = [&](const float* tile, int c0, float (&dst)[16]) {   // keys c0 .. c0 + 15 of a key tile, this lane's row
#pragma unroll
        for (int u = 0; u < 16; ++u) dst[u] = __ldg(tile + (c0 + u) * 32);
      };
      float acc[D];
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = 0.f;
      float m = -FLT_MAX, l = 0.f;
      float bv0[16], bv1[16], bv2[16], bv3[16];
      unsigned long long ph[6] = {0, 0, 0, 0, 0, 0};
      long long tp = 0;
#define ABX_PH(k) do { if (prof) { const long long n__ = clock64(); ph[k] += n__ - tp; tp = n__; } } while (0)
      load_bias(bt, 0, bv0);
      load_bias(bt, 16, bv1);
      for (int kt = 0; kt < nkt; ++kt) {
        const uint32_t par = kt & 1;
        if (kt > 0) {                                  // O' of the previous tile (relative to the same reference point as acc)
          ABX_ATTN_PWAIT(0, o_full + slot, par ^ 1);
          if (prof) tp = clock64();
          tc_fence_after();
          float t[D];
#pragma unroll
          for (int c0 = 0; c0 < D; c0 += 16) tmem_ld16_nowait(lane_base + 192 + c0, t + c0);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < D; ++d) acc[d] += t[d];
          ABX_PH(0);
        }
        if (prof) tp = clock64();
        load_bias(bt, 32, bv2);
        load_bias(bt, 48, bv3);
        ABX_PH(5);
        ABX_ATTN_PWAIT(1, s_full + slot, par);         // S = Q K^T of this tile
        tc_fence_after();
        if (prof) tp = clock64();
        // S is read and biased in two halves of 32 keys: never more than acc (D) + 64 logits + 32 bias values live — with all
        // 64 bias values and 64 logits in flight at once the loop spilled, and loaded bias values went straight to local memory
        float sv[kTK];
        float cm = -FLT_MAX;
        auto chunk = [&](int c0, const float (&bv)[16]) {   // base-2 logits of keys j0 + c0 .. + 15 (mask and padding ride in the bias)
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float a = sv[c0 + u] + bv[u];
            sv[c0 + u] = a;
            cm = fmaxf(cm, a);
          }
        };
        tmem_ld16_nowait(lane_base, sv);
        tmem_ld16_nowait(lane_base + 16, sv + 16);
        tmem_ld_wait();
        ABX_PH(1);
        chunk(0, bv0);
        chunk(16, bv1);
        tmem_ld16_nowait(lane_base + 32, sv + 32);
        tmem_ld16_nowait(lane_base + 48, sv + 48);
        tmem_ld_wait();
        chunk(32, bv2);
        chunk(48, bv3);
        bt += tile_stride;
        ABX_PH(2);
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
        if (kt + 1 < nkt) {                            // first half of the next tile's bias: lands during the waits for O' and S
          load_bias(bt, 0, bv0);
          load_bias(bt, 16, bv1);
        }
        ABX_PH(3);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(p_full + slot);                    // P written; S and O' of this slot have been read
        ABX_PH(4);
      }
      if (prof && threadIdx.x == 0)
        for (int k2 = 0; k2 < 6; ++k2) g_attn_prof[16 + k2] = ph[k2];
      {
        ABX_ATTN_PWAIT(0, o_full + slot, (nkt - 1) & 1);
        tc_fence_after();
        float t[D];
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 16) tmem_ld16_nowait(lane_base + 192 + c0, t + c0);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] += t[d];
      }
      if (prof && threadIdx.x == 0) { g_attn_prof[0] = pw[0]; g_attn_prof[1] = pw[1]; g_attn_prof[2] = clock64() - tstart; g_attn_prof[12] = nkt; }
      const long long tstore = clock64();
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
            o.x *= sigmoid_fast(g4.x); o.y *= sigmoid_fast(g4.y);
            o.z *= sigmoid_fast(g4.z); o.w *= sigmoid_fast(g4.w);
          }
          *reinterpret_cast<float4*>(orow + d) = o;
        }
      }
      if (prof && threadIdx.x == 0) g_attn_prof[14] = clock64() - tstore;   // output (gate loads + stores issued)
    }
  } else if (warp < 12) {
    // ================= loader warps: K / V tile -> hi / lo operand tiles, two stages =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLoad));
    auto split4 = [](const float4& x, float4& hi, float4& lo) {
      hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
      lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
    };
    auto stage = [&](uint8_t* st) {                  // registers -> hi / lo operand tiles of one stage
      uint8_t* k_hi = st, *k_lo = st + SM::kKBytes, *v_hi = st + 2 * SM::kKBytes, *v_lo = v_hi + SM::kVBytes;
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        float4 hi, lo;
        split4(kreg[u], hi, lo);                     // K tile: rows = keys, 16-byte column block vc0 + u
        const uint32_t ko = (vc0 + u) * SM::kLboK + vkey * 16;
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
    for (int kt = 0; kt < nkt; ++kt) {                 // tile 0 was fetched in the prologue
      const int st = kt & 1;
      if (kt >= kStagesKV) ABX_ATTN_PWAIT(0, kv_empty + st, ((kt >> 1) - 1) & 1);   // the products of tile kt - 2 have read the stage
      stage(sm + SM::kv_base + st * SM::kStageBytes);
      fence_proxy_async();
      mbar_arrive(kv_full + st);
      if (kt + 1 < nkt) fetch(kt + 1);
    }
    if (prof && t == 0) { g_attn_prof[4] = pw[0]; g_attn_prof[5] = clock64() - tstart; }
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
        ABX_ATTN_PWAIT(0, kv_full + 0, 0);
        tc_fence_after();
        for (int slot = 0; slot < nslots; ++slot) issue_s(slot, 0);
        for (int kt = 0; kt < nkt; ++kt) {
          const int st = kt & 1;
          for (int slot = 0; slot < nslots; ++slot) {
            ABX_ATTN_PWAIT(1, p_full + slot, kt & 1);
            const long long ti = prof ? clock64() : 0;
            tc_fence_after();
            issue_pv(slot, st);
            if (slot == nslots - 1) umma_commit(kv_empty + st);     // every product reading stage st has been issued
            if (kt + 1 < nkt) {
              if (slot == 0) {
                const long long tw = prof ? clock64() : 0;
                mbar_wait(kv_full + (st ^ 1), ((kt + 1) >> 1) & 1);
                if (prof) { const long long d = clock64() - tw; pw[0] += d; pw[2] -= d; }
                tc_fence_after();
              }
              issue_s(slot, st ^ 1);
            }
            if (prof) pw[2] += clock64() - ti;
          }
        }
        if (prof) { g_attn_prof[8] = pw[0]; g_attn_prof[9] = pw[1]; g_attn_prof[10] = pw[2]; g_attn_prof[11] = clock64() - tstart; }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) g_attn_prof[15] = clock64() - tentry;       // whole CTA up to the final barrier
  if (warp == 13) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <int D>
int launch_attention_tc5(cudaStream_t st, int B, int S, int L, int H, const float* q, const float* k, const float* v, int ld,
                         const float* bias_tiles, const float* gate, float* out) {
  const size_t smem = Tc5Smem<D>::total(L);
  ABX_REQUIRE(smem <= 227 * 1024, "abx_pair_attention: L=%d with head dim %d needs %zu bytes of shared memory per CTA", L, D, smem);
  ABX_CUDA(cudaFuncSetAttribute(pair_attention_tc5_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nqt = (L + kTQ - 1) / kTQ;
  pair_attention_tc5_kernel<D><<<dim3(H, B * S, (nqt + kSlots - 1) / kSlots), kThreads, smem, st>>>(
      L, H, S, q, k, v, ld, bias_tiles, gate, kLog2e / sqrtf((float)D), out);
  count_launch();
  return check_launch("pair_attention_tc5_kernel");
}

int attention_tc5_profile(unsigned long long* out16) {  // 32 counters
  ABX_CUDA(cudaDeviceSynchronize());
  ABX_CUDA(cudaMemcpyFromSymbol(out16, g_attn_prof, sizeof(g_attn_prof)));
  return ABX_OK;
}

// shared memory the kernel needs for key length L (the dispatcher falls back to the mma.sync kernel beyond 227 KB: D = 64)
template <int D> size_t attention_tc5_smem(int L) { return Tc5Smem<D>::total(L); }
template size_t attention_tc5_smem<16>(int);
template size_t attention_tc5_smem<32>(int);
template size_t attention_tc5_smem<48>(int);

template int launch_attention_tc5<16>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, float*);
template int launch_attention_tc5<32>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, float*);
template int launch_attention_tc5<48>(cudaStream_t, int, int, int, int, const float*, const float*, const float*, int, const float*,
                                      const float*, float*);

}  // namespace abx
