// Fused Invariant Point Attention core (abx/model/folding.py:79-128) — ONE warp-specialised kernel for
//   logits (scalar q.k + point distances + pair bias, mask)  :79-109
//   softmax over the keys                                     :110
//   attention over scalar / point values, inverse rigid transform, point norms   :114-123
//   attention over the pair activations  o_pair[i,h,:] = sum_j a[h,i,j] z[i,j,:]  :126-127
// with the pair tensor z [B,N,N,128] read exactly once and no [B,H,N,N] logits / probability tensor.
//
// The O(N^2 Cz) part — 1536 of the ~2350 multiply-adds per (query, key) pair — does not fit the fp32 SIMT pipe at
// HBM speed (6 flop per z byte), so it runs on the 5th-generation tensor cores as 3xTF32 (fp32-level accuracy):
//     D_i[c, h] += sum_k z[i, j0+k, c] * p[h, i, j0+k]        M = 128 channels, N = 16 (12 heads), K = 8 keys
// with A = z in TENSOR MEMORY (lane = channel, column = key; hi = the raw fp32 word — the tensor core ignores the 13 low
// mantissa bits — and lo = z - hi, written with tcgen05.st), B = the chunk's probabilities (hi / lo tiles in shared
// memory, K-major, no swizzle) and the accumulator D_i in tensor memory (16 columns per query row).
//
// One CTA (28 warps, register budgets re-balanced with setmaxnreg) = a tile of R <= 20 query rows of one batch
// element; it walks the keys in chunks of 8 with every row in lock step, so the chunk's packed key / value operands
// (26 KB, one bulk copy, served by L2) are shared by the R rows.  Roles:
//   warp 16       z producer: lane r issues one 4 KB cp.async.bulk per chunk for query row r — a (row, chunk) "item" — into a
//                 shared-memory ring (mbarrier tx counts); lane 31: key/value + pair-bias producer (double-buffered chunk)
//   warps 20-27   two converter warpgroups (alternating pairs of items): thread = channel reads the item's 8 keys, splits
//                 hi / lo, tcgen05.st into a 12-slot A ring in tensor memory, releases the z slot
//   warps 8-11    logits + online softmax: thread = (head, 2 query rows) with its queries in registers, all 8 keys of the
//                 chunk; exact fp32 FFMA2 arithmetic.  The softmax reference point m is fixed by the first chunk and only
//                 moves when a later logit exceeds it by 2^64 (flagged, see below), so accumulators are never rescaled in
//                 the common case.  Writes p (hi / lo MMA operand tiles + a plain copy for the value warps).
//   warps 0-7     attention over the 40-wide value rows: thread = (10 query rows x 4 value dims) register tile
//   warps 17-19   MMA issuers (query rows mod 3): per item 3 x tcgen05.mma (hi*lo, lo*hi, hi*hi), commit -> frees the A slot
//   warps 12-15   accumulator service: rescales D_i in tensor memory when a reference point moved (rare), and at the end
//                 reads D (tcgen05.ld), normalises by the row sums and writes o_pair
// The pair bias sqrt(1/3)(z W^T + b) is read from the chunked key-major tensor written by ipa_pair_bias_kernel
// ([B, ceil(N/8), N, 100]: row i of chunk c holds bias[j0+k][h] at 12 k + h): one 8 KB bulk copy per chunk and tile.
#include <float.h>
#include <stdlib.h>

#include <type_traits>

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
constexpr int kChunk = 8;                     // keys per chunk = K of one tcgen05.mma.kind::tf32
constexpr int kMaxRows = 20;                  // query rows per CTA
constexpr int kHalfRows = kMaxRows / 2;       // a logit thread owns rows rp and rp + 10; a value thread rows 10 rg .. 10 rg + 9
constexpr int kBiasRow = ABX_IPA_BIAS_ROW;    // floats per (chunk, row) of the chunked pair bias (96 used)
constexpr int kPD = 2;                        // chunks of probabilities in flight
constexpr int kPTileBytes = 1040;             // per item: hi tile 512 B | lo tile 512 B | 16 B stagger (bank spread)
constexpr int kPfRow = 24;                    // plain probabilities pf[key][head][24]: row r at r + 2 (r / 10)
constexpr int kASlots = 12, kACol0 = kMaxRows * 16;   // tensor memory: D_i at columns 16 i, A rings (3 x 4 slots) at 320 + 16 slot
constexpr uint32_t kTmemCols = 512;
constexpr int kZSlotBytes = kChunk * kCz * 4; // 4096
constexpr int kKVChunkBytes = kChunk * kKVRow * 4;    // 26112
constexpr int kCols = 3;                      // "columns": query row r is converted by warpgroup r % 3 and issued by MMA warp r % 3
constexpr int kThreads = 896;
constexpr int kValThreads = 128;
constexpr int kIssuers = kCols;               // MMA issuer warps
constexpr int kRegsCtl = 48, kRegsConv = 48, kRegsSvc = 40, kRegsLogit = 128, kRegsVal = 128;   // 62464 of 65536
static_assert(128 * kRegsCtl + 384 * kRegsConv + 128 * kRegsSvc + 128 * kRegsLogit + 128 * kRegsVal <= 65536, "register budget");
constexpr int kFeatPt = kH * kSv, kFeatNorm = kFeatPt + 3 * kH * kPv, kFeatPair = kFeatNorm + kH * kPv;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleGap = 64.f;           // log2 units: the reference point moves when a logit exceeds it by this much
                                              // (ABX_IPA_RESCALE_GAP overrides it: a small gap makes the rare path the common one in tests)

// Warp ids are priorities: among the eligible warps of a scheduler the highest id issues first.  The latency-critical roles
// (converters, MMA issuers, producers) therefore sit above the FFMA2-dense logit / value warps, which have slack.
enum Warps { kWarpVal = 0, kWarpLogit = 4, kWarpSvc = 8, kWarpZ = 12, kWarpMma = 13, kWarpConv0 = 16 };
enum WarpGroups { kWgLogit = kWarpLogit / 4, kWgSvc = kWarpSvc / 4, kWgCtl = kWarpZ / 4, kWgConv0 = kWarpConv0 / 4 };

// ---- shared-memory carve-up (byte offsets from a 128-byte aligned base) ----
struct Smem {
  uint32_t kv, bias, ptile, pf, al, linv, mask, ov, zring, bars, total;
};
__host__ __device__ inline Smem smem_layout(int N, int zslots) {
  Smem s;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
  s.kv = take(2 * kKVChunkBytes);
  s.bias = take(2 * kMaxRows * kBiasRow * 4);
  s.ptile = take(kPD * kMaxRows * kPTileBytes);
  s.pf = take(kPD * kChunk * kH * kPfRow * 4);
  s.al = take(kPD * kH * kPfRow * 4);
  s.linv = take(kH * kPfRow * 4);
  s.mask = take((uint32_t)((N + kChunk - 1) / kChunk) * kChunk * 4);
  s.ov = s.kv;                                   // epilogue staging of the value outputs reuses the key/value chunks
  s.zring = take((uint32_t)zslots * kZSlotBytes);
  s.bars = take(1024);
  s.total = o + 128;                             // alignment slack
  return s;
}
static_assert(kMaxRows * kH * kVD * 4 <= 2 * kKVChunkBytes, "value staging must fit the key/value buffers");

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
// Watchdog: a wait that spins for more than ~0.25 s records (tag, block, thread, parity) in g_ipa_watchdog and raises the
// abort flag; every wait of every CTA then returns at once, so a synchronisation bug ends the kernel (with garbage results and
// a readable record, abx_ipa_watchdog_read) instead of hanging the device.
__device__ unsigned long long g_ipa_watchdog[8];
__device__ unsigned long long g_ipa_prof[64];
template <bool kSleep = false, bool kTime = false, bool kSpin = false>
__device__ __forceinline__ unsigned mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {   // kTime: returns the cycles spent waiting
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  const long long t0 = kTime ? clock64() : 0ll;    // try_wait itself blocks for a while: time the first attempt too
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  if (ok) return kTime ? (unsigned)(clock64() - t0) : 0u;
  const long long t1 = kTime ? t0 : clock64();
  // Slow path: the thread is suspended inside try_wait (up to the hinted time) instead of spinning on the issue slots the
  // working warps of the same scheduler need; the watchdog is only consulted when a suspension timed out.
  const uint32_t hint_ns = kSleep ? 200000u : 20000u;
  do {
    if constexpr (kSpin) {                           // latency-critical warps (producers, issuers, converters): plain polling
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(addr), "r"(parity), "r"(hint_ns) : "memory");
    }
    if (!ok) {
      if (*reinterpret_cast<volatile unsigned long long*>(&g_ipa_watchdog[0]) != 0ull) break;
      if (clock64() - t1 > 500000000ll) {
        if (atomicCAS(&g_ipa_watchdog[0], 0ull, 1ull) == 0ull) {
          g_ipa_watchdog[1] = (unsigned long long)tag;
          g_ipa_watchdog[2] = blockIdx.x;
          g_ipa_watchdog[3] = threadIdx.x;
          g_ipa_watchdog[4] = parity;
          __threadfence();
        }
        break;
      }
    }
  } while (!ok);
  return kTime ? (unsigned)(clock64() - t0) : 0u;
}

// optional per-role stall profile (tools/bench_ipa.py --prof): cycles of every role's main loop and of its two waits, summed
// over the CTAs, at prof[8 role + {0: loop, 1: first wait, 2: second wait, 3: contributors}]
struct RoleProf {
  unsigned t0, w[4];                               // 32-bit cycle counts: a kernel runs far less than 2^32 cycles
  __device__ __forceinline__ void start() { t0 = (unsigned)clock64(); w[0] = w[1] = w[2] = w[3] = 0u; }
  __device__ __forceinline__ void flush(unsigned long long* prof, int role) {
    if (prof == nullptr) return;
    atomicAdd(prof + 8 * role + 0, (unsigned long long)((unsigned)clock64() - t0));
    atomicAdd(prof + 8 * role + 1, (unsigned long long)w[0]);
    atomicAdd(prof + 8 * role + 2, (unsigned long long)w[1]);
    atomicAdd(prof + 8 * role + 3, 1ull);
    atomicAdd(prof + 8 * role + 4, (unsigned long long)w[2]);
    atomicAdd(prof + 8 * role + 5, (unsigned long long)w[3]);
  }
};
// global -> shared bulk copy (bytes and both addresses multiples of 16) completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// B operand tile [16 heads x 8 keys] tf32, K-major without swizzle: core matrices of 8 rows x 16 bytes (128 B, rows 16 B
// apart); the two k-halves 128 B apart (LBO), the two 8-row groups 256 B apart (SBO).  Descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_ptile(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t ptile_offset(int h, int k) {   // byte offset of p[h][k] inside a tile
  return (uint32_t)((h >> 3) * 256 + (k >> 2) * 128 + (h & 7) * 16 + (k & 3) * 4);
}
// instruction descriptor, kind::tf32: D f32, A/B tf32, both K-major, N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Ring ownership.  An mbarrier wait only carries a parity, so a waiter must never reach the wait for use n of a slot before
// use n-1 has completed: every ring therefore has ONE producer and ONE consumer that walk it in the same order.  Query row r
// belongs to "column" r % 3 in every chunk: it is streamed into z ring r % 3 by producer lane r, converted by converter
// warpgroup r % 3 into A ring r % 3 (4 tensor-memory slots) and issued by MMA warp r % 3.
__device__ __forceinline__ int col_rows(int nvalid, int j) { return nvalid > j ? (nvalid - j + kCols - 1) / kCols : 0; }
__device__ __forceinline__ int zring_base(int zslots, int j) { return j * (zslots / kCols) + min(j, zslots % kCols); }
__device__ __forceinline__ int zring_size(int zslots, int j) { return zslots / kCols + (j < zslots % kCols ? 1 : 0); }

__device__ __forceinline__ int pf_row(int r) { return r + 2 * (r / kHalfRows); }

}  // namespace

// ---------------------------------------------------------------------------------------------------
// pack: row of the fused node projection [q_scalar 192 | kv_scalar 384 | q_point_local 144 | kv_point_local 432]
// (folding.py:69-86) -> packed queries Qp [B,N,12,28] = (q_s * sqrt(1/48), 4 query points in the global frame) and
// packed keys / values KVp [B,N,816] = 12 x (k_s 16, 4 key points) then 12 x (v_s 16, 8 value points), points
// moved to the global frame (r3.rigids_apply, r3.py:9-16).  64 threads per residue: the projection row is staged in shared
// memory with coalesced reads, every thread then writes consecutive output floats.
// ---------------------------------------------------------------------------------------------------
constexpr int kProj = 1152, kOffKV = kH * kSqk, kOffQP = kOffKV + kH * (kSqk + kSv), kOffKVP = kOffQP + 3 * kH * kPqk;
constexpr int kPackRes = 4, kPackThreads = 64 * kPackRes;   // residues per CTA, 64 threads each

__global__ void __launch_bounds__(kPackThreads) ipa_pack_nodes_kernel(int BN, const float* __restrict__ proj,
                                                                      const float* __restrict__ rots, const float* __restrict__ trans,
                                                                      float* __restrict__ Qp, float* __restrict__ KVp) {
  __shared__ __align__(16) float row_s[kPackRes][kProj];
  __shared__ float rt_s[kPackRes][12];
  griddep_wait();                                    // proj comes from the node GEMM launched just before
  griddep_launch_dependents();
  const int sub = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int bn = blockIdx.x * kPackRes + sub;
  const bool live = bn < BN;
  if (live) {                                        // coalesced 16-byte reads of the projection row; frame of the residue
    const float4* src = reinterpret_cast<const float4*>(proj + (size_t)bn * kProj);
    for (int k = t; k < kProj / 4; k += 64) reinterpret_cast<float4*>(row_s[sub])[k] = src[k];
    if (t < 9) rt_s[sub][t] = __ldg(rots + (size_t)bn * 9 + t);
    else if (t < 12) rt_s[sub][t] = __ldg(trans + (size_t)bn * 3 + (t - 9));
  }
  __syncthreads();
  if (!live) return;
  const float* row = row_s[sub];
  const float* Rm = rt_s[sub];
  const float* tr = rt_s[sub] + 9;
  // one global point coordinate: component k of R l + t, l = the point's three local coordinates ('(r n)' layout, stride)
  auto point = [&](const float* l, int stride, int k) {
    return tr[k] + (Rm[3 * k] * l[0] + Rm[3 * k + 1] * l[stride] + Rm[3 * k + 2] * l[2 * stride]);
  };
  const float w_scalar = sqrtf(1.0f / (3.0f * kSqk));                           // folding.py:59,79
  float* q = Qp + (size_t)bn * kQRow;
  for (int e = t; e < kQRow; e += 64) {              // packed queries: per head 16 scalars * w_scalar, 4 global points
    const int h = e / kQK, o = e % kQK;
    float v;
    if (o < kSqk) v = row[h * kSqk + o] * w_scalar;
    else { const int p = (o - kSqk) / 3, k = (o - kSqk) % 3; v = point(row + kOffQP + h * kPqk + p, kH * kPqk, k); }
    q[e] = v;
  }
  float* kv = KVp + (size_t)bn * kKVRow;
  for (int e = t; e < kKVRow; e += 64) {             // packed keys (12 x 28) then values (12 x 40)
    float v;
    if (e < kVOff) {
      const int h = e / kQK, o = e % kQK;
      if (o < kSqk) v = row[kOffKV + h * (kSqk + kSv) + o];
      else { const int p = (o - kSqk) / 3, k = (o - kSqk) % 3; v = point(row + kOffKVP + h * (kPqk + kPv) + p, kH * (kPqk + kPv), k); }
    } else {
      const int h = (e - kVOff) / kVD, o = (e - kVOff) % kVD;
      if (o < kSv) v = row[kOffKV + h * (kSqk + kSv) + kSqk + o];
      else { const int p = (o - kSv) / 3, k = (o - kSv) % 3; v = point(row + kOffKVP + h * (kPqk + kPv) + kPqk + p, kH * (kPqk + kPv), k); }
    }
    kv[e] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------------
// timed wait of role profile slot i (compiled out of the product instantiation)
#define ABX_WAIT(i, ...)                                                   \
  do {                                                                     \
    if constexpr (kProf) rp_.w[i] += mbar_wait<false, true>(__VA_ARGS__);  \
    else mbar_wait<false, false>(__VA_ARGS__);                             \
  } while (0)
#define ABX_WAIT_SPIN(i, ...)                                                    \
  do {                                                                           \
    if constexpr (kProf) rp_.w[i] += mbar_wait<false, true, true>(__VA_ARGS__);  \
    else mbar_wait<false, false, true>(__VA_ARGS__);                             \
  } while (0)

template <bool kProf>
__global__ void __launch_bounds__(kThreads, 1)
ipa_fused_kernel(int N, int R, int tiles_per_b, int zslots, const float* __restrict__ Qp, const float* __restrict__ KVp,
                 const float* __restrict__ bias, const float* __restrict__ mask, const float* __restrict__ rots,
                 const float* __restrict__ trans, const float* __restrict__ point_weights, const float* __restrict__ z,
                 float* __restrict__ feats, unsigned long long* __restrict__ prof, float rescale_gap) {
  // no swizzled operand tiles here: 16-byte alignment (bulk copies, descriptors) is all the carve-up needs, so the dynamic
  // shared array is used as is and every access below stays a shared-state-space instruction (LDS / STS)
  extern __shared__ __align__(128) uint8_t sm[];
  const Smem L = smem_layout(N, zslots);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x / tiles_per_b, i0 = (blockIdx.x % tiles_per_b) * R;
  const int nvalid = min(R, N - i0);                // query rows of this tile
  const int nchunks = (N + kChunk - 1) / kChunk;

  float* KVs = reinterpret_cast<float*>(sm + L.kv);           // [2][8][816]
  float* BSs = reinterpret_cast<float*>(sm + L.bias);         // [2][R][100]
  uint8_t* PT = sm + L.ptile;                                  // [kPD][R] tiles of kPTileBytes
  float* PF = reinterpret_cast<float*>(sm + L.pf);            // [kPD][8][12][24]
  float* AL = reinterpret_cast<float*>(sm + L.al);            // [kPD][12][24] rescale factors of the chunk
  float* LINV = reinterpret_cast<float*>(sm + L.linv);        // [12][24] 1 / row sum
  float* MS = reinterpret_cast<float*>(sm + L.mask);          // [nchunks * 8] key mask (0 beyond N)
  uint8_t* ZR = sm + L.zring;                                  // [zslots][4096]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint64_t* kv_full = bars;                         // [2]
  uint64_t* kv_empty = bars + 2;                    // [2]  4 logit warps + 4 value warps
  uint64_t* p_full = bars + 4;                      // [kPD] 4 logit warps
  uint64_t* p_empty = bars + 6;                     // [kPD] MMA issuers' commits + 4 value warps
  uint64_t* a_full = bars + 8;                      // [12] 4 converter warps
  uint64_t* a_empty = bars + 20;                    // [12] MMA commit
  uint64_t* req = bars + 32;                        // MMA -> accumulator service
  uint64_t* resp = bars + 33;                       // 4 service warps -> MMA
  uint64_t* drain = bars + 34;                      // [kIssuers] issuer's commit: everything it issued so far has completed
  uint64_t* lsum_ready = bars + 37;                 // 4 logit warps: LINV is final
  uint64_t* z_full = bars + 40;                     // [zslots]
  uint64_t* z_empty = z_full + zslots;              // [zslots] 4 converter warps
  volatile int* req_info = reinterpret_cast<volatile int*>(z_empty + zslots);    // {row or -1, probability buffer}
  unsigned* resc = reinterpret_cast<unsigned*>(const_cast<int*>(req_info) + 2);  // [4] rows whose reference point moved in chunk c & 3
  unsigned* svc_lock = resc + 4;                    // issuers take turns talking to the accumulator service
  unsigned* svc_phase = resc + 5;                   // number of rescale requests served so far (parity of `resp`)
  unsigned* issuers_done = resc + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resc + 8);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 4 + kValThreads / 32); }
    for (int s = 0; s < kPD; ++s) { mbar_init(p_full + s, 4); mbar_init(p_empty + s, kValThreads / 32 + kIssuers); }
    for (int s = 0; s < kASlots; ++s) { mbar_init(a_full + s, 4); mbar_init(a_empty + s, 1); }
    mbar_init(req, 1); mbar_init(resp, 4); for (int i = 0; i < kIssuers; ++i) mbar_init(drain + i, 1);
    mbar_init(lsum_ready, 4);
    for (int s = 0; s < zslots; ++s) { mbar_init(z_full + s, 1); mbar_init(z_empty + s, 4); }
    for (int s = 0; s < 8; ++s) resc[s] = 0u;       // flags, service lock / phase, issuers_done
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // probability tiles (rows 12-15 of every B tile stay zero), plain probabilities (rows that do not exist stay zero),
  // rescale factors (1), key mask
  for (int k = threadIdx.x; k < kPD * kMaxRows * kPTileBytes / 16; k += kThreads) reinterpret_cast<float4*>(PT)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = threadIdx.x; k < kPD * kChunk * kH * kPfRow; k += kThreads) PF[k] = 0.f;
  for (int k = threadIdx.x; k < kPD * kH * kPfRow; k += kThreads) AL[k] = 1.f;
  for (int k = threadIdx.x; k < kH * kPfRow; k += kThreads) LINV[k] = 0.f;
  for (int k = threadIdx.x; k < nchunks * kChunk; k += kThreads) MS[k] = (k < N) ? __ldg(mask + (size_t)b * N + k) : 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // z is an input of the whole layer, not a product of the kernels launched just before: the first chunk of every row starts
  // streaming before the grid-dependency wait (programmatic dependent launch), everything else after it.  The wait is executed
  // by every thread at a converged point.
  if (warp == kWarpZ && lane < nvalid) {
    const int slot = zring_base(zslots, lane % kCols) + lane / kCols;
    const uint32_t bytes = (uint32_t)min(kChunk, N) * kCz * 4;
    mbar_expect_tx(z_full + slot, bytes);
    bulk_g2s(ZR + (size_t)slot * kZSlotBytes, z + ((size_t)b * N + i0 + lane) * (size_t)N * kCz, bytes, z_full + slot);
  }
  __syncwarp();
  griddep_wait();

  const int wg = warp >> 2;
  if (wg == kWgCtl) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
    if (warp == kWarpZ) {
      // ---------------- producers: lane r streams query row r of z (z is an input of the whole layer: no wait for the
      // preceding kernels), lane 31 the key/value + pair-bias chunks.  The warp stays converged: every lane polls its own slot
      // without blocking, the lanes whose slot is free issue their copy in the same pass, the others retry.  (Independent
      // per-lane loops would be executed one lane after the other: 20 lanes x ~50 instructions per chunk is the kernel's pace.)
      const bool zlane = lane < nvalid, kvlane = lane == 31;
      const uint8_t* zb = reinterpret_cast<const uint8_t*>(z + ((size_t)b * N + i0 + (zlane ? lane : 0)) * (size_t)N * kCz);
      const int col = lane % kCols, zbase = zring_base(zslots, col), zn = zring_size(zslots, col);   // this row's ring
      const int nw = col_rows(nvalid, col);          // rows of the tile in that ring = ring positions per chunk (<= zn)
      int pos = lane / kCols + nw;                   // ring position of (chunk c, this row), modulo zn; chunk 0 went out before
      uint32_t ph = 1u;                              // the grid-dependency wait; parity of the "empty" phase before the slot's use
      if (pos >= zn) { pos -= zn; ph ^= 1u; }
      int c = zlane ? 1 : 0;                         // next chunk this lane issues
      const bool has_work = zlane || kvlane;
      const long long t_start = clock64();
      for (;;) {
        const bool work = has_work && c < nchunks;
        if (!__any_sync(0xffffffffu, work)) break;
        bool ready = false;
        if (work) {
          uint64_t* bar = zlane ? (z_empty + zbase + pos) : (kv_empty + (c & 1));
          const uint32_t par = zlane ? ph : (uint32_t)(((c >> 1) & 1) ^ 1);
          uint32_t ok;
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
              "selp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(ok) : "r"(smem_u32(bar)), "r"(par) : "memory");
          ready = ok != 0u;
        }
        if (ready) {
          const int nk = min(kChunk, N - c * kChunk);
          if (zlane) {
            const int slot = zbase + pos;
            const uint32_t bytes = (uint32_t)nk * kCz * 4;
            mbar_expect_tx(z_full + slot, bytes);
            bulk_g2s(ZR + (size_t)slot * kZSlotBytes, zb + (size_t)c * kZSlotBytes, bytes, z_full + slot);
            pos += nw;
            if (pos >= zn) { pos -= zn; ph ^= 1u; }
          } else {
            const int buf = c & 1;
            const uint32_t kvb = (uint32_t)nk * kKVRow * 4, bb = (uint32_t)nvalid * kBiasRow * 4;
            mbar_expect_tx(kv_full + buf, kvb + bb);
            bulk_g2s(KVs + (size_t)buf * kChunk * kKVRow, KVp + ((size_t)b * N + c * kChunk) * kKVRow, kvb, kv_full + buf);
            bulk_g2s(BSs + (size_t)buf * kMaxRows * kBiasRow, bias + (((size_t)b * nchunks + c) * N + i0) * kBiasRow, bb, kv_full + buf);
          }
          ++c;
        } else if (work && clock64() - t_start > 1000000000ll) {   // watchdog (same record as mbar_wait's)
          if (atomicCAS(&g_ipa_watchdog[0], 0ull, 1ull) == 0ull) {
            g_ipa_watchdog[1] = zlane ? 101ull : 201ull; g_ipa_watchdog[2] = blockIdx.x; g_ipa_watchdog[3] = threadIdx.x;
            __threadfence();
          }
          c = nchunks;
        } else if (*reinterpret_cast<volatile unsigned long long*>(&g_ipa_watchdog[0]) != 0ull) {
          c = nchunks;
        }
      }
    } else if (warp >= kWarpMma && warp < kWarpMma + kIssuers) {
      // ---------------- MMA issuers: warps 1, 2, 3 take the query rows r = 0, 1, 2 (mod 3) ----------------
      // The whole warp walks the loop (uniform control flow keeps the operand addresses in uniform registers); one elected
      // lane issues.  Row r is always served by the same warp, so the accumulations into D_r stay ordered.
      uint64_t* my_drain = drain + (warp - kWarpMma);
      uint32_t dr = 0;
      const uint32_t ptile0 = smem_u32(PT);
      RoleProf rp_; if constexpr (kProf) rp_.start();
      const int mi = warp - kWarpMma;
      const unsigned my_rows = 0x249249u << mi;      // bits r = mi (mod 3)
      int seq = 0;                                   // items issued so far = position in A ring mi
      for (int c = 0; c < nchunks; ++c) {
        const int pb = c % kPD;
        ABX_WAIT_SPIN(0, p_full + pb, (c / kPD) & 1, 301);
        tc_fence_after();
        const unsigned rm = *reinterpret_cast<volatile unsigned*>(resc + (c & 3)) & my_rows;
        const uint32_t acc = c > 0 ? 1u : 0u;
        if (rm != 0u) {
          // Rare: the reference point of some head of some of this issuer's rows moved in this chunk.  All MMAs issued so
          // far (chunk c - 1 of those rows included) are drained, then the accumulator service rescales D_r row by row.
          if (elect_one()) {
            umma_commit(my_drain);
            mbar_wait(my_drain, dr, 303);
            for (int r = mi; r < nvalid; r += kIssuers) {
              if (!((rm >> r) & 1u)) continue;
              while (atomicCAS(svc_lock, 0u, 1u) != 0u) { }
              const uint32_t ph = *reinterpret_cast<volatile uint32_t*>(svc_phase);
              req_info[0] = r; req_info[1] = pb;
              mbar_arrive(req);
              mbar_wait(resp, ph & 1u, 304);
              *reinterpret_cast<volatile uint32_t*>(svc_phase) = ph + 1u;
              __threadfence_block();
              atomicExch(svc_lock, 0u);
            }
          }
          dr ^= 1;
          __syncwarp();
          tc_fence_after();
        }
        const uint32_t pt = ptile0 + (uint32_t)(pb * kMaxRows) * kPTileBytes;
        for (int r = mi; r < nvalid; r += kIssuers, ++seq) {
          const int aslot = 4 * mi + (seq & 3);
          ABX_WAIT_SPIN(1, a_full + aslot, (uint32_t)(seq >> 2) & 1u, 302);
          tc_fence_after();
          const uint32_t d = tmem_base + 16u * r;
          const uint32_t a_hi = tmem_base + kACol0 + 16u * aslot, a_lo = a_hi + 8u;
          const uint64_t b_hi = umma_desc_ptile(pt + (uint32_t)r * kPTileBytes);
          const uint64_t b_lo = b_hi + (512 >> 4);
          if (elect_one()) {
            umma_tf32_ts(d, a_hi, b_lo, acc);
            umma_tf32_ts(d, a_lo, b_hi, 1u);
            umma_tf32_ts(d, a_hi, b_hi, 1u);
            umma_commit(a_empty + aslot);            // A slot free once these MMAs have read it
          }
        }
        if (elect_one()) umma_commit(p_empty + pb);  // this warp's share of the chunk's probability tiles is consumed
      }
      __syncwarp();
      if constexpr (kProf) { if (lane == 0) rp_.flush(prof, 2); }
      if (elect_one()) {
        umma_commit(my_drain);
        mbar_wait(my_drain, dr, 305);
        if (atomicAdd(issuers_done, 1u) == kIssuers - 1) {       // the last issuer tells the service to write the results
          req_info[0] = -1; req_info[1] = 0;
          mbar_arrive(req);
        }
      }
      __syncwarp();
    }
  } else if (wg >= kWgConv0) {
    // ---------------- converters: z item (8 keys x 128 channels in shared memory) -> A hi / lo in tensor memory ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsConv));
    const int cw = wg - kWgConv0;                    // this warpgroup converts the query rows r = cw (mod 3), in chunk order
    const int q = warp & 3, ch = 32 * q + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16) + kACol0 + 64u * cw;   // A ring cw: 4 slots of 16 columns
    const int zbase = zring_base(zslots, cw), zn = zring_size(zslots, cw);
    RoleProf rp_; if constexpr (kProf) rp_.start();
    // Two items per round, every stage for both at once: one shared-memory round trip and one hand-over to the MMA issuer per
    // pair; the hand-over of a round (tcgen05.wait::st + arrive) is done while the next round's loads are in flight.
    {
      int seq = 0;                                   // items converted so far = position in A ring cw
      int zpos = 0;
      uint32_t zph = 0u;
      const uint32_t zr0 = smem_u32(ZR) + (uint32_t)zbase * kZSlotBytes + (uint32_t)ch * 4u;
      uint64_t* const zf = z_full + zbase;
      uint64_t* const ze = z_empty + zbase;
      uint64_t* const af = a_full + 4 * cw;
      uint64_t* const ae = a_empty + 4 * cw;
      auto load_item = [&](int zs, int nk, uint32_t (&hi)[16]) {
        const uint32_t addr = zr0 + (uint32_t)zs * kZSlotBytes;
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
          uint32_t v = 0u;
          if (kk < nk) asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr + kk * kCz * 4));
          hi[kk] = v;
        }
      };
      auto low_part = [&](uint32_t (&v)[16]) {
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) v[8 + kk] = __float_as_uint(__uint_as_float(v[kk]) - __uint_as_float(v[kk] & 0xffffe000u));
      };
      int pendA = -1, pendB = -1;                    // A slots whose stores are in flight (signalled one round later)
      auto hand_over = [&]() {                       // previous round: stores done -> hand the slots to the MMA issuer
        if (pendA >= 0) {
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(af + pendA);
            if (pendB >= 0) mbar_arrive(af + pendB);
          }
        }
      };
      for (int c = 0; c < nchunks; ++c) {
        const int nk = min(kChunk, N - c * kChunk);
        for (int rA = cw; rA < nvalid; rA += 2 * kCols) {
          const bool two = rA + kCols < nvalid;
          const int asA = seq & 3, asB = two ? ((seq + 1) & 3) : -1;
          const uint32_t apA = ((uint32_t)(seq >> 2) & 1u) ^ 1u, apB = ((uint32_t)((seq + 1) >> 2) & 1u) ^ 1u;
          seq += two ? 2 : 1;
          const int zsA = zpos;
          const uint32_t zpA = zph;
          if (++zpos == zn) { zpos = 0; zph ^= 1u; }
          const int zsB = zpos;
          const uint32_t zpB = zph;
          if (two) { if (++zpos == zn) { zpos = 0; zph ^= 1u; } }
          uint32_t vA[16], vB[16];                  // [0, 8): hi = the raw words, [8, 16): lo — one 16-column tcgen05.st per item
          ABX_WAIT_SPIN(0, zf + zsA, zpA, 401);
          if (two) ABX_WAIT_SPIN(0, zf + zsB, zpB, 401);
          load_item(zsA, nk, vA);                    // 16 independent loads in flight
          if (two) load_item(zsB, nk, vB);
          hand_over();                               // overlaps the shared-memory round trip of this round's loads
          low_part(vA);
          if (two) low_part(vB);
          __syncwarp();                              // every lane has read (and used) the z slots
          if (lane == 0) {
            mbar_arrive(ze + zsA);
            if (two) mbar_arrive(ze + zsB);
          }
          ABX_WAIT_SPIN(1, ae + asA, apA, 402);
          if (two) ABX_WAIT_SPIN(1, ae + asB, apB, 402);
          tc_fence_after();
          tmem_st16(lane_base + 16u * asA, vA);
          if (two) tmem_st16(lane_base + 16u * asB, vB);
          pendA = asA; pendB = asB;
        }
      }
      hand_over();
    }
    if constexpr (kProf) { if (lane == 0) rp_.flush(prof, 3); }
  } else if (wg == kWgSvc) {
    // ---------------- accumulator service: rescale D_r on request; final normalisation + o_pair store ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsSvc));
    const int q = warp & 3, ch = 32 * q + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t ph = 0;
    for (;;) {
      mbar_wait<true>(req, ph, 501); ph ^= 1;
      const int r = req_info[0], pb = req_info[1];
      if (r < 0) break;
      if (threadIdx.x == kWarpSvc * 32) atomicAdd(&g_ipa_watchdog[6], 1ull);     // statistics: rescale requests served
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(lane_base + 16u * r, v);
      const float* al = AL + (size_t)pb * kH * kPfRow + pf_row(r);
#pragma unroll
      for (int h = 0; h < kH; ++h) v[h] = __float_as_uint(__uint_as_float(v[h]) * al[h * kPfRow]);
      {
        uint32_t lo8[8], hi8[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) { lo8[h] = v[h]; hi8[h] = v[8 + h]; }
        tmem_st8(lane_base + 16u * r, lo8);
        tmem_st8(lane_base + 16u * r + 8u, hi8);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(resp);
    }
    tc_fence_after();
    mbar_wait(lsum_ready, 0, 502);
    for (int r = 0; r < nvalid; ++r) {
      uint32_t v[16];
      tmem_ld16(lane_base + 16u * r, v);
      float* frow = feats + ((size_t)b * N + i0 + r) * kFeat + kFeatPair + ch;       // 'b i h c -> b i (h c)'  folding.py:126-127
      const float* li = LINV + pf_row(r);
#pragma unroll
      for (int h = 0; h < kH; ++h) frow[h * kCz] = __uint_as_float(v[h]) * li[h * kPfRow];
    }
  } else if (wg == kWgLogit) {
    // ---------------- logits + softmax: thread = (head h, query rows rp and rp + 10), all keys of the chunk ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsLogit));
    const int t = threadIdx.x - kWarpLogit * 32;
    const bool worker = t < kH * kHalfRows;
    const int h = worker ? t / kHalfRows : 0, rp = worker ? t % kHalfRows : 0;
    const int r0 = rp, r1 = rp + kHalfRows;
    const bool act0 = worker && r0 < nvalid, act1 = worker && r1 < nvalid;
    float q0[kQK], q1[kQK];
    {
      const float4* s0 = reinterpret_cast<const float4*>(Qp + ((size_t)b * N + i0 + (act0 ? r0 : 0)) * kQRow + h * kQK);
      const float4* s1 = reinterpret_cast<const float4*>(Qp + ((size_t)b * N + i0 + (act1 ? r1 : 0)) * kQRow + h * kQK);
#pragma unroll
      for (int u = 0; u < kQK / 4; ++u) {
        const float4 a = __ldg(s0 + u), c4 = __ldg(s1 + u);
        q0[4 * u] = a.x; q0[4 * u + 1] = a.y; q0[4 * u + 2] = a.z; q0[4 * u + 3] = a.w;
        q1[4 * u] = c4.x; q1[4 * u + 1] = c4.y; q1[4 * u + 2] = c4.z; q1[4 * u + 3] = c4.w;
      }
    }
    const float pw = __ldg(point_weights + h);
    const float gamma = (pw > 20.f) ? pw : log1pf(expf(pw));                       // F.softplus  folding.py:96
    const float coef = -0.5f * sqrtf(1.0f / (3.0f * kPqk * 9.0f / 2.0f)) * gamma;  // -1/2 w_point gamma  :97-99
    const float mi0 = act0 ? __ldg(mask + (size_t)b * N + i0 + r0) : 0.f;
    const float mi1 = act1 ? __ldg(mask + (size_t)b * N + i0 + r1) : 0.f;
    float m0 = -FLT_MAX, m1 = -FLT_MAX, l0 = 0.f, l1 = 0.f;
    const float2 neg1 = make_float2(-1.f, -1.f);

    RoleProf rp_; if constexpr (kProf) rp_.start();
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1, pb = c % kPD, j0 = c * kChunk;
      ABX_WAIT_SPIN(0, kv_full + buf, (c >> 1) & 1, 601);
      ABX_WAIT_SPIN(1, p_empty + pb, ((c / kPD) & 1) ^ 1, 602);
      if (t == 0) resc[(c + 2) & 3] = 0u;
      const float* kvp = KVs + (size_t)buf * kChunk * kKVRow + h * kQK;
      const float* bs0 = BSs + ((size_t)buf * kMaxRows + r0) * kBiasRow + h;
      const float* bs1 = BSs + ((size_t)buf * kMaxRows + r1) * kBiasRow + h;
      float s0[kChunk], s1[kChunk];
#pragma unroll
      for (int kk = 0; kk < kChunk; ++kk) {
        const float4* kp = reinterpret_cast<const float4*>(kvp + kk * kKVRow);
        float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f), e0 = make_float2(0.f, 0.f), e1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kSqk / 4; ++u) {
          const float4 kv = kp[u];
          const float2 ka = make_float2(kv.x, kv.y), kb = make_float2(kv.z, kv.w);
          d0 = ffma2(make_float2(q0[4 * u], q0[4 * u + 1]), ka, d0);
          d0 = ffma2(make_float2(q0[4 * u + 2], q0[4 * u + 3]), kb, d0);
          d1 = ffma2(make_float2(q1[4 * u], q1[4 * u + 1]), ka, d1);
          d1 = ffma2(make_float2(q1[4 * u + 2], q1[4 * u + 3]), kb, d1);
        }
#pragma unroll
        for (int u = kSqk / 4; u < kQK / 4; ++u) {
          const float4 kv = kp[u];
          const float2 ka = make_float2(kv.x, kv.y), kb = make_float2(kv.z, kv.w);
          const float2 a0 = ffma2(ka, neg1, make_float2(q0[4 * u], q0[4 * u + 1])), b0 = ffma2(kb, neg1, make_float2(q0[4 * u + 2], q0[4 * u + 3]));
          const float2 a1 = ffma2(ka, neg1, make_float2(q1[4 * u], q1[4 * u + 1])), b1 = ffma2(kb, neg1, make_float2(q1[4 * u + 2], q1[4 * u + 3]));
          e0 = ffma2(a0, a0, e0); e0 = ffma2(b0, b0, e0);
          e1 = ffma2(a1, a1, e1); e1 = ffma2(b1, b1, e1);
        }
        const float mj = MS[j0 + kk];
        const float lg0 = (((d0.x + d0.y) + coef * (e0.x + e0.y)) + bs0[kk * kH]) * kLog2e;   // log-2 units
        const float lg1 = (((d1.x + d1.y) + coef * (e1.x + e1.y)) + bs1[kk * kH]) * kLog2e;
        s0[kk] = (mi0 * mj != 0.f) ? lg0 : -FLT_MAX;                                           // mask_2d  folding.py:106-109
        s1[kk] = (mi1 * mj != 0.f) ? lg1 : -FLT_MAX;
      }
      const int nk = min(kChunk, N - j0);
      auto softmax_row = [&](float (&s)[kChunk], float& m, float& l, int r, bool act) {
        if (!act) return;
        float cm = -FLT_MAX;
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) cm = (kk < nk) ? fmaxf(cm, s[kk]) : cm;
        float alpha = 1.f;
        if (c == 0) {
          m = cm;
        } else if (cm > m + rescale_gap) {           // move the reference point: accumulators are scaled by alpha
          alpha = exp2f(m - cm);
          m = cm;
          l *= alpha;
          atomicOr(resc + (c & 3), 1u << r);
        }
        AL[((size_t)pb * kH + h) * kPfRow + pf_row(r)] = alpha;
        float p[kChunk];
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
          p[kk] = (kk < nk) ? exp2f(s[kk] - m) : 0.f;
          l += p[kk];
        }
        uint8_t* tile = PT + (size_t)(pb * kMaxRows + r) * kPTileBytes;
        float ph[kChunk], pl[kChunk];
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
          ph[kk] = __uint_as_float(__float_as_uint(p[kk]) & 0xffffe000u);
          pl[kk] = p[kk] - ph[kk];
        }
        *reinterpret_cast<float4*>(tile + ptile_offset(h, 0)) = make_float4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<float4*>(tile + ptile_offset(h, 4)) = make_float4(ph[4], ph[5], ph[6], ph[7]);
        *reinterpret_cast<float4*>(tile + 512 + ptile_offset(h, 0)) = make_float4(pl[0], pl[1], pl[2], pl[3]);
        *reinterpret_cast<float4*>(tile + 512 + ptile_offset(h, 4)) = make_float4(pl[4], pl[5], pl[6], pl[7]);
        float* pf = PF + ((size_t)pb * kChunk * kH + h) * kPfRow + pf_row(r);
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) pf[kk * kH * kPfRow] = p[kk];
      };
      softmax_row(s0, m0, l0, r0, act0);
      softmax_row(s1, m1, l1, r1, act1);
      fence_proxy_async();                           // the probability tiles are read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) { mbar_arrive(p_full + pb); mbar_arrive(kv_empty + buf); }
    }
    if constexpr (kProf) { if (lane == 0) rp_.flush(prof, 4); }
    if (act0) LINV[h * kPfRow + pf_row(r0)] = 1.f / l0;
    if (act1) LINV[h * kPfRow + pf_row(r1)] = 1.f / l1;
    __syncwarp();
    if (lane == 0) mbar_arrive(lsum_ready);
  } else {
    // ---------------- attention over the value rows: thread = (head h, dims 4 d4 ..) x all 20 query rows ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsVal));
    const int t = threadIdx.x - kWarpVal * 32;
    const bool worker = t < kH * (kVD / 4);
    const int dg = worker ? t : 0;
    const int h = dg / (kVD / 4), d4 = dg % (kVD / 4);
    float2 acc[kMaxRows / 2][4];                     // [row pair][dim]: (row 2j, row 2j + 1)
#pragma unroll
    for (int j = 0; j < kMaxRows / 2; ++j)
#pragma unroll
      for (int d = 0; d < 4; ++d) acc[j][d] = make_float2(0.f, 0.f);

    RoleProf rp_; if constexpr (kProf) rp_.start();
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1, pb = c % kPD, nk = min(kChunk, N - c * kChunk);
      ABX_WAIT_SPIN(0, kv_full + buf, (c >> 1) & 1, 701);
      ABX_WAIT_SPIN(1, p_full + pb, (c / kPD) & 1, 702);
#pragma unroll
      for (int g = 0; g < 2; ++g) {                  // rows 10 g .. 10 g + 9 live at [12 g, 12 g + 10) of a 24-float row
        const float4* ap = reinterpret_cast<const float4*>(AL + ((size_t)pb * kH + h) * kPfRow + g * 12);
        const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2];
        const bool moved = (a0.x != 1.f) | (a0.y != 1.f) | (a0.z != 1.f) | (a0.w != 1.f) | (a1.x != 1.f) | (a1.y != 1.f) |
                           (a1.z != 1.f) | (a1.w != 1.f) | (a2.x != 1.f) | (a2.y != 1.f);
        if (moved) {
          const float2 f[5] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w),
                               make_float2(a2.x, a2.y)};
#pragma unroll
          for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int d = 0; d < 4; ++d) { acc[5 * g + j][d].x *= f[j].x; acc[5 * g + j][d].y *= f[j].y; }
        }
      }
      const float* vbase = KVs + (size_t)buf * kChunk * kKVRow + kVOff + h * kVD + 4 * d4;
      const float* pbase = PF + ((size_t)pb * kChunk * kH + h) * kPfRow;
      auto one_key = [&](int kk) {
        const float4 v = *reinterpret_cast<const float4*>(vbase + kk * kKVRow);
        const float2 vx = make_float2(v.x, v.x), vy = make_float2(v.y, v.y), vz = make_float2(v.z, v.z), vw = make_float2(v.w, v.w);
        const float4* pp = reinterpret_cast<const float4*>(pbase + kk * kH * kPfRow);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const float4 p0 = pp[3 * g], p1 = pp[3 * g + 1], p2 = pp[3 * g + 2];
          const float2 pr[5] = {make_float2(p0.x, p0.y), make_float2(p0.z, p0.w), make_float2(p1.x, p1.y), make_float2(p1.z, p1.w),
                                make_float2(p2.x, p2.y)};
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            acc[5 * g + j][0] = ffma2(pr[j], vx, acc[5 * g + j][0]);
            acc[5 * g + j][1] = ffma2(pr[j], vy, acc[5 * g + j][1]);
            acc[5 * g + j][2] = ffma2(pr[j], vz, acc[5 * g + j][2]);
            acc[5 * g + j][3] = ffma2(pr[j], vw, acc[5 * g + j][3]);
          }
        }
      };
      if (nk == kChunk) {
#pragma unroll 2
        for (int kk = 0; kk < kChunk; ++kk) one_key(kk);
      } else {
        for (int kk = 0; kk < nk; ++kk) one_key(kk);
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(p_empty + pb); mbar_arrive(kv_empty + buf); }
    }

    if constexpr (kProf) { if (lane == 0) rp_.flush(prof, 5); }
    // ---- normalise, stage the 480 value outputs of every row, then write o_scalar / o_point / o_point_norm ----
    mbar_wait(lsum_ready, 0, 703);
    asm volatile("bar.sync 1, 128;" ::: "memory");   // every value warp is done with the key/value buffers
    float* OV = reinterpret_cast<float*>(sm + L.ov); // [R][480]
    if (worker) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4* lp = reinterpret_cast<const float4*>(LINV + h * kPfRow + g * 12);
        const float4 i0v = lp[0], i1v = lp[1], i2v = lp[2];
        const float inv[10] = {i0v.x, i0v.y, i0v.z, i0v.w, i1v.x, i1v.y, i1v.z, i1v.w, i2v.x, i2v.y};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int ra = g * kHalfRows + 2 * j;
          const float2* a = acc[5 * g + j];
          *reinterpret_cast<float4*>(OV + (size_t)ra * (kH * kVD) + h * kVD + 4 * d4) =
              make_float4(a[0].x * inv[2 * j], a[1].x * inv[2 * j], a[2].x * inv[2 * j], a[3].x * inv[2 * j]);
          *reinterpret_cast<float4*>(OV + (size_t)(ra + 1) * (kH * kVD) + h * kVD + 4 * d4) =
              make_float4(a[0].y * inv[2 * j + 1], a[1].y * inv[2 * j + 1], a[2].y * inv[2 * j + 1], a[3].y * inv[2 * j + 1]);
        }
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int e = t; e < nvalid * kH * kSv; e += kValThreads) {                                          // 'b i h c -> b i (h c)'  :115
      const int r = e / (kH * kSv), o = e % (kH * kSv);
      feats[((size_t)b * N + i0 + r) * kFeat + o] = OV[(size_t)r * (kH * kVD) + (o / kSv) * kVD + (o % kSv)];
    }
    for (int e = t; e < nvalid * kH * kPv; e += kValThreads) {
      const int r = e / (kH * kPv), pi = e % (kH * kPv), hh = pi / kPv, p = pi % kPv;
      const size_t bn = (size_t)b * N + i0 + r;
      float Rm[9], tr[3], it[3], l[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) Rm[k] = __ldg(rots + bn * 9 + k);
#pragma unroll
      for (int k = 0; k < 3; ++k) tr[k] = __ldg(trans + bn * 3 + k);
      // invert_rigids (r3.py:54-59): R^T, -(R^T t); then rigids_apply  folding.py:121
#pragma unroll
      for (int k = 0; k < 3; ++k) it[k] = -(Rm[k] * tr[0] + Rm[3 + k] * tr[1] + Rm[6 + k] * tr[2]);
      const float* g = OV + (size_t)r * (kH * kVD) + hh * kVD + kSv + 3 * p;
#pragma unroll
      for (int k = 0; k < 3; ++k) l[k] = it[k] + (Rm[k] * g[0] + Rm[3 + k] * g[1] + Rm[6 + k] * g[2]);
      float* frow = feats + bn * kFeat;
#pragma unroll
      for (int k = 0; k < 3; ++k) frow[kFeatPt + k * (kH * kPv) + pi] = l[k];                  // '(r n)'  folding.py:122
      frow[kFeatNorm + pi] = sqrtf(l[0] * l[0] + l[1] * l[1] + l[2] * l[2] + 1e-8f);           // :123
    }
  }

  griddep_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// chunked key-major pair bias: out[b][c][i][12 k + h] = sqrt(1/3) (z[b,i,8c+k,:] . w[h,:] + b[h])   folding.py:101-104
// Persistent CTAs walk the (b, i, 128-key tile) list; the z tile is staged in shared memory by cp.async (row stride 132
// floats: conflict-free float4 reads with lanes on consecutive keys), the weights once per CTA; three CTAs per SM overlap
// each other's copy and compute phases.  thread = (2 keys, 6 heads): a 16-byte shared-memory read costs four passes whether it
// is a broadcast or not, so the weight reads (6 per 4 channels) dominate — with two keys per thread they are shared by twice
// the multiply-adds (the first version, one key per thread, sat at 3.1 TB/s on exactly that).
// ---------------------------------------------------------------------------------------------------
constexpr int kBiasJ = 128, kZld = kCz + 4, kBiasThreads = 128;
constexpr size_t kBiasSmem = (size_t)(kBiasJ * kZld + kH * kCz) * sizeof(float);

__global__ void __launch_bounds__(kBiasThreads) ipa_pair_bias_chunked_kernel(int N, int tiles_per_row, int total_tiles,
                                                                             const float* __restrict__ z,
                                                                             const float* __restrict__ w_pair,
                                                                             const float* __restrict__ b_pair,
                                                                             float* __restrict__ bias) {
  extern __shared__ __align__(16) float bias_smem[];
  float* zs = bias_smem;                              // [kBiasJ * kZld]
  float* ws = bias_smem + kBiasJ * kZld;              // [kH * kCz]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = (N + kChunk - 1) / kChunk;
  for (int k = tid; k < kH * kCz / 4; k += kBiasThreads)
    reinterpret_cast<float4*>(ws)[k] = __ldg(reinterpret_cast<const float4*>(w_pair) + k);
  const int hg = warp >> 1;                           // heads 6 hg .. 6 hg + 5
  const int ka = (warp & 1) * 64 + lane, kb = ka + 32; // this thread's two keys of the tile
  float bsc[6];
  const float w_pair_scale = sqrtf(1.0f / 3.0f);
#pragma unroll
  for (int u = 0; u < 6; ++u) bsc[u] = __ldg(b_pair + 6 * hg + u);
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    // tile t -> (row = b N + i, 128-key tile jt); all in 32 bits
    const unsigned row = (unsigned)t / (unsigned)tiles_per_row, jt = (unsigned)t - row * (unsigned)tiles_per_row;
    const unsigned b = row / (unsigned)N, i = row - b * (unsigned)N;
    const int j0 = (int)jt * kBiasJ, nj = min(kBiasJ, N - j0);
    const float4* zrow = reinterpret_cast<const float4*>(z + ((size_t)row * N + j0) * kCz);
    for (int k = tid; k < nj * (kCz / 4); k += kBiasThreads) {
      const int jj = k / (kCz / 4), c4 = k % (kCz / 4);
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(zs + jj * kZld + 4 * c4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(zrow + k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                   // the tile (and, the first time, the weights) are in shared memory
    if (ka < nj) {
      const bool two = kb < nj;
      const float* za_p = zs + ka * kZld;
      const float* zb_p = zs + (two ? kb : ka) * kZld;
      float2 acc[2][6];
#pragma unroll
      for (int u = 0; u < 6; ++u) acc[0][u] = acc[1][u] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int c = 0; c < kCz; c += 4) {
        const float4 za = *reinterpret_cast<const float4*>(za_p + c), zb = *reinterpret_cast<const float4*>(zb_p + c);
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[(6 * hg + u) * kCz + c]);
          const float2 w0 = make_float2(wv.x, wv.y), w1 = make_float2(wv.z, wv.w);
          acc[0][u] = ffma2(make_float2(za.x, za.y), w0, acc[0][u]);
          acc[0][u] = ffma2(make_float2(za.z, za.w), w1, acc[0][u]);
          acc[1][u] = ffma2(make_float2(zb.x, zb.y), w0, acc[1][u]);
          acc[1][u] = ffma2(make_float2(zb.z, zb.w), w1, acc[1][u]);
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q == 1 && !two) break;
        const int j = j0 + (q ? kb : ka);
        float* dst = bias + (((size_t)b * nchunks + j / kChunk) * N + i) * kBiasRow + (j % kChunk) * kH + 6 * hg;
#pragma unroll
        for (int u = 0; u < 6; u += 2)
          *reinterpret_cast<float2*>(dst + u) = make_float2(w_pair_scale * ((acc[q][u].x + acc[q][u].y) + bsc[u]),
                                                            w_pair_scale * ((acc[q][u + 1].x + acc[q][u + 1].y) + bsc[u + 1]));
      }
    }
    __syncthreads();                                   // the tile is overwritten by the next iteration's copy
  }
}

// Tile height: time ~ rounds * (R + 6.4) — R rows of z per tile plus the tile's key / value chunks (3264 B per key
// against 512 B of z per key and row), rounds = ceil(tiles / SMs).
static int choose_rows(int B, int N, int sms) {
  static int forced = [] { const char* e = getenv("ABX_IPA_ROWS"); return e ? atoi(e) : 0; }();   // test knob: fixed tile height
  if (forced >= 1 && forced <= kMaxRows) return forced < N ? forced : N;
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

// copies and clears the watchdog record: out[0] != 0 -> a wait timed out; out[1..4] = tag, block, thread, parity
int ipa_watchdog_read(unsigned long long* out) {
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("ipa_watchdog_read: %s", cudaGetErrorString(cudaGetLastError())); return ABX_ERR_CUDA; }
  ABX_CUDA(cudaMemcpyFromSymbol(out, g_ipa_watchdog, sizeof(zero)));
  ABX_CUDA(cudaMemcpyToSymbol(g_ipa_watchdog, zero, sizeof(zero)));
  return ABX_OK;
}

static float rescale_gap() {
  static float v = [] { const char* e = getenv("ABX_IPA_RESCALE_GAP"); return e ? (float)atof(e) : kRescaleGap; }();
  return v;
}

static bool g_prof_on = false;
// enable != 0: later launches accumulate the per-role stall profile; out64 (optional): copy and clear what was gathered
int ipa_prof(int enable, unsigned long long* out64) {
  unsigned long long zero[64] = {0};
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("ipa_prof: %s", cudaGetErrorString(cudaGetLastError())); return ABX_ERR_CUDA; }
  if (out64) ABX_CUDA(cudaMemcpyFromSymbol(out64, g_ipa_prof, sizeof(zero)));
  ABX_CUDA(cudaMemcpyToSymbol(g_ipa_prof, zero, sizeof(zero)));
  g_prof_on = enable != 0;
  return ABX_OK;
}

size_t ipa_fused_qp_floats(int B, int N) { return (size_t)B * N * kQRow; }
size_t ipa_fused_kvp_floats(int B, int N) { return (size_t)B * N * kKVRow; }
size_t ipa_pair_bias_floats(int B, int N) { return (size_t)B * ceil_div(N, kChunk) * N * kBiasRow; }

int launch_ipa_pair_bias(cudaStream_t s, int B, int N, const float* z, const float* w_pair, const float* b_pair, float* bias) {
  const int tiles_per_row = ceil_div(N, kBiasJ);
  const long long total = (long long)B * N * tiles_per_row;
  ABX_REQUIRE(total < 2147483647LL, "abx_ipa_pair_bias: too many tiles");
  ABX_CUDA(cudaFuncSetAttribute(ipa_pair_bias_chunked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBiasSmem));
  const long long ctas = 3LL * fused_sm_count();      // 73.7 KB of shared memory each: three resident CTAs per SM
  ipa_pair_bias_chunked_kernel<<<(unsigned)(total < ctas ? total : ctas), kBiasThreads, kBiasSmem, s>>>(N, tiles_per_row, (int)total, z,
                                                                                                       w_pair, b_pair, bias);
  count_launch();
  return check_launch("ipa_pair_bias_chunked_kernel");
}

int launch_ipa_pack_nodes(cudaStream_t s, int B, int N, const float* proj, const float* rots, const float* trans,
                          float* Qp, float* KVp) {
  const cudaError_t le = launch_kernel(ipa_pack_nodes_kernel, dim3(ceil_div(B * N, kPackRes)), dim3(kPackThreads), 0, s, B * N,
                                       proj, rots, trans, Qp, KVp);
  count_launch();
  if (le != cudaSuccess) { set_error("launch of ipa_pack_nodes_kernel failed: %s", cudaGetErrorString(le)); return ABX_ERR_CUDA; }
  return check_launch("ipa_pack_nodes_kernel");
}

int launch_ipa_fused(cudaStream_t s, int B, int N, const float* Qp, const float* KVp, const float* bias, const float* mask,
                     const float* rots, const float* trans, const float* point_weights, const float* z, float* feats) {
  const int R = choose_rows(B, N, fused_sm_count());
  const int tiles_per_b = ceil_div(N, R);
  // z ring: as many 4 KB slots as the 227 KB of shared memory leave (at most 32)
  int zslots = 32;
  while (zslots > 4 && smem_layout(N, zslots).total > 227 * 1024) --zslots;
  const Smem L = smem_layout(N, zslots);
  ABX_REQUIRE(L.total <= 227 * 1024, "ipa_fused: N=%d needs %u bytes of shared memory", N, L.total);
  ABX_REQUIRE(zslots / kCols >= (R + kCols - 1) / kCols, "ipa_fused: N=%d leaves %d z slots, too few for %d-row tiles", N, zslots, R);
  unsigned long long* prof = nullptr;
  if (g_prof_on) ABX_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&prof), g_ipa_prof));
  auto* kernel = g_prof_on ? ipa_fused_kernel<true> : ipa_fused_kernel<false>;
  ABX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  const cudaError_t le = launch_kernel(kernel, dim3(B * tiles_per_b), dim3(kThreads), (size_t)L.total, s, N, R, tiles_per_b,
                                       zslots, Qp, KVp, bias, mask, rots, trans, point_weights, z, feats, prof, rescale_gap());
  count_launch();
  if (le != cudaSuccess) { set_error("launch of ipa_fused_kernel failed: %s", cudaGetErrorString(le)); return ABX_ERR_CUDA; }
  return check_launch("ipa_fused_kernel");
}

}  // namespace abx
