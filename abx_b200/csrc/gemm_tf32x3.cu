// fp32-accurate dense GEMM on the 5th-generation tensor cores:  y = act(x W^T + b) (+ residual)
//
// torch.nn.Linear semantics of the reference's `Linear` (abx/model/common_modules.py:11-59) for the node
// and pair GEMMs of the score network (IPA projections folding.py:69-86,130-132; IpaScore / trunk layers).
// The 1e-4 A parity budget of the sampler rules out single-pass TF32/bf16, so every fp32 operand is split
// exactly into hi = x & ~0x1fff (a TF32 value) and lo = x - hi, and
//     D = A_hi B_lo + A_lo B_hi + A_hi B_hi        (3xTF32; dropped A_lo B_lo term <= 2^-22 relative)
// accumulates in fp32 in tensor memory.
//
// Persistent CTAs (one per SM) walk the 128 x BN output tiles; 16 warps in four role-aligned warpgroups
// (register budgets re-balanced with setmaxnreg):
//   WG0  warp 0: TMA producer — cp.async.bulk.tensor (128-byte swizzle) of the raw fp32 A / W k-slabs
//                (128 x 32 and BN x 32 elements) into a shared-memory ring
//        warp 1: allocates TMEM; one lane issues 3 x 4 tcgen05.mma.kind::tf32 per slab — A FROM TENSOR MEMORY
//                (TS mode), W from shared memory — into a ring of TMEM partial-sum buffers and commits to the
//                slab's `empty` and the buffer's `acc_ready` barriers
//   WG1  converters: thread = row of the landed A slab: reads its 32 floats from the swizzled tile, splits
//        hi / lo and writes both with tcgen05.st into the slab's stage of the TMEM A ring (the A operand never
//        goes back to shared memory: the SS-mode loop sat on the shared-memory port, 192 KB per slab against
//        128 KB now); then writes the lo tile of W (the tensor core ignores the 13 low mantissa bits of a tf32
//        operand, so the raw W tile already is its hi part), fence generic->async proxy, arrive `conv`
//   WG2, WG3  accumulator warpgroups, alternating tiles (ping-pong): per slab read the partial sum with
//        tcgen05.ld (warp w owns TMEM lanes 32 (w%4)..) and add it to register accumulators with
//        round-to-nearest — the tensor core itself accumulates with truncation, which over K/8 chained steps
//        would bias the result by ~K/8 ulp — then run the epilogue (bias, activation, gate, row scale,
//        residual) straight from registers while the other warpgroup accumulates the next tile.
#include <cuda.h>
#include <stdlib.h>

#include <map>
#include <string>
#include <type_traits>

#include "common.cuh"

namespace abx {

namespace {

#ifndef ABX_GEMM_BK
#define ABX_GEMM_BK 32
#endif
// k-slab width: 32 fp32 = one 128-byte swizzle row, or 16 fp32 = one 64-byte swizzle row (twice the pipeline depth
// in the same shared memory: the TMA -> convert -> MMA -> commit round trip, not bandwidth, paces the main loop)
constexpr int kBM = 128, kBK = ABX_GEMM_BK;
static_assert(kBK == 32, "k-slab = one 128-byte swizzle row (the converter reads the A tile row by row)");
constexpr int kDrainDefault = 2;                // k-slabs per TMEM partial sum (register accumulation once per 64 columns of K)
constexpr int kThreads = 512, kConvThreads = 128, kAccThreads = 128;
constexpr int kRegsCtl = 48, kRegsConv = 72, kRegsAcc = 192;   // 128*(48+72) + 256*192 = 65536 - 1024
constexpr uint32_t kTileABytes = kBM * kBK * 4;
static_assert(kBM == kConvThreads, "converter thread = A row = tensor-memory lane");

template <int BN> struct GemmCfg {
  static constexpr uint32_t kTileBBytes = BN * kBK * 4;
  static constexpr uint32_t kStageBytes = kTileABytes + 2 * kTileBBytes;       // raw A (staging only), raw W (= hi), W lo
  static constexpr int kStages = 4 * (32 / kBK);                               // = stages of the TMEM A ring
  static constexpr int kAccBufs = BN >= 128 ? 2 : 4;                           // ring of TMEM partial-sum buffers
  static constexpr uint32_t kAColBase = kAccBufs * BN;                         // A ring: stage s at columns kAColBase + 2 kBK s (hi | lo)
  static constexpr uint32_t kTmemCols = 512;
  static_assert(kAColBase + kStages * 2 * kBK <= kTmemCols, "tensor memory budget");
  static constexpr uint32_t kStagingBytes = 8 * 4096;                          // epilogue: 32 rows x 32 columns per accumulator warp
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /* alignment slack */ + 512 /* barriers */ + kStagingBytes;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

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

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// K-major operand tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * kBK * 4) >> 4) << 32;     // SBO: 8 rows of kBK floats (1024 B / 512 B)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(kBK == 32 ? 2 : 4) << 61;      // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// instruction descriptor, kind::tf32: D f32 (bits 4-5 = 1), A/B tf32 (bits 7-9 / 10-12 = 2), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (128 lanes x 8 columns per k-step), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Per-role wait / work cycle counters of CTA 0 (ABX_GEMM_PROF=1; read with abx_gemm_profile): what the roles wait for.
// Compiled in only with -DABX_GEMM_PROFILE=1 (ABX_GEMM_PROFILE=1 python -m abx_b200.build): the counters cost registers.
#ifndef ABX_GEMM_PROFILE
#define ABX_GEMM_PROFILE 0
#endif
__device__ unsigned long long g_gemm_prof[32];
#define ABX_GEMM_PWAIT(slot, bar, par)                                                  \
  do {                                                                                  \
    if (prof) { const long long t0__ = clock64(); mbar_wait(bar, par); pw[slot] += clock64() - t0__; } \
    else mbar_wait(bar, par);                                                           \
  } while (0)

struct Epilogue {
  const float* bias;       // [Nout] or null
  const float* residual;   // [M, ldy] or null (added last)
  const float* gate;       // [M, ldy]: act 2: v * sigmoid(gate);  act 4: sigmoid(v) * gate
  const float* row_scale;  // [M] or null: multiplies the activated value (pair / sequence masks)
  int act;                 // 0 none, 1 relu, 2 v * sigmoid(gate), 3 sigmoid(v), 4 sigmoid(v) * gate
  int transpose_n;         // n > 0: rows are (b,i,j) of a [B,n,n,*] tensor and row (b,i,j) is stored to (b,j,i);
                           // y and residual use the transposed row, gate / row_scale the GEMM's own row
  int cm_n, cm_np;         // GLU only, cm_n = n > 0: rows are (b,i,k) of a [B,n,n,*] tensor and output channel c of row
                           // (b,i,k) goes to y[b][c][i][k] with rows of cm_np >= n floats (channel-major operands of
                           // the triangle-multiplication product)
};

// Batched "NT" product mode of the kernel (triangle multiplication, seqformer.py:470-500): problem bc of `batches`
// multiplies rows [base(bc), base(bc)+rows) of the A and B operand maps; base(bc) = (bc / inner) * outer_rows +
// (bc % inner) * rows.  Output tile rows/cols are local to the problem: y[bc][i][j], leading dimension ldy.
struct Batched {
  int batches, rows, inner, outer_rows;
};

template <int ACT>
__device__ __forceinline__ float epilogue_op(float a, float bias, float gate, float scale, float res) {
  a += bias;
  if (ACT == 1) a = fmaxf(a, 0.f);
  else if (ACT == 2) a = a * sigmoid_fast(gate);
  else if (ACT == 3) a = sigmoid_fast(a);
  else if (ACT == 4) a = gate * sigmoid_fast(a);
  return a * scale + res;
}

// Epilogue of one output row held by this lane: acc[0..BN) are columns n0.. of row `row`.  Each lane streams
// its own row in 16-column pieces (4 x 16-byte stores; gate / residual pieces are loaded first).
template <int ACT, int BN>
__device__ __forceinline__ void store_row(const float (&acc)[BN], const Epilogue& ep, float* __restrict__ y, int ldy,
                                          int row, int n0, int Nout, bool vec_ok) {
  const size_t go = (size_t)row * ldy;                // gate offset (GEMM row)
  size_t ro = go;                                     // output / residual offset
  if (ep.transpose_n > 0) {
    const long long n = ep.transpose_n, nn = n * n, b = row / nn, r = row % nn;
    ro = (size_t)(b * nn + (r % n) * n + r / n) * ldy;
  }
  const float sc = ep.row_scale ? __ldg(ep.row_scale + row) : 1.f;
#pragma unroll
  for (int c0 = 0; c0 < BN; c0 += 16) {
    const int col = n0 + c0;
    if (col >= Nout) break;
    if (vec_ok && col + 15 < Nout) {
      float4 g[4], r[4];
      if (ACT == 2 || ACT == 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = *reinterpret_cast<const float4*>(ep.gate + go + col + 4 * u);
      }
      if (ep.residual) {
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = *reinterpret_cast<const float4*>(ep.residual + ro + col + 4 * u);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 bv = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = epilogue_op<ACT>(acc[c0 + 4 * u + 0], bv.x, g[u].x, sc, r[u].x);
        o.y = epilogue_op<ACT>(acc[c0 + 4 * u + 1], bv.y, g[u].y, sc, r[u].y);
        o.z = epilogue_op<ACT>(acc[c0 + 4 * u + 2], bv.z, g[u].z, sc, r[u].z);
        o.w = epilogue_op<ACT>(acc[c0 + 4 * u + 3], bv.w, g[u].w, sc, r[u].w);
        *reinterpret_cast<float4*>(y + ro + col + 4 * u) = o;
      }
    } else {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (col + u < Nout) {
          const float bv = ep.bias ? __ldg(ep.bias + col + u) : 0.f;
          const float gg = (ACT == 2 || ACT == 4) ? ep.gate[go + col + u] : 0.f;
          const float rr = ep.residual ? ep.residual[ro + col + u] : 0.f;
          y[ro + col + u] = epilogue_op<ACT>(acc[c0 + u], bv, gg, sc, rr);
        }
      }
    }
  }
}

// Coalesced epilogue of one warp's 32 rows (lane = row holds acc[0..BN) = columns n0..): 32 columns at a time go through a
// 4 KB per-warp staging tile (16-byte pieces XOR-swizzled by the row: both directions at the 4-wavefront minimum), then lane
// (row 4 i + lane / 8, piece lane % 8) loads gate / residual and stores y as whole 128-byte row segments — a warp-level store
// touches 4 lines instead of 32 half-used sectors (the row-per-lane stores of store_row kept the LSU busy ~10k clk per tile and
// slowed the other warpgroup's tensor-memory reads down with them).  Requires Nout % 32 == 0, 16-byte aligned operands, no transpose.
// With a gate operand the chunk / row-step loops stay rolled (measured: 1.88 -> 1.28 ms for the 980000 x 192 x 128 gate + residual
// layer), without one they are unrolled (the plain store is ~7 % faster that way).
template <int ACT, int BN>
__device__ __forceinline__ void store_tile_coalesced(const float (&acc)[BN], const Epilogue& ep, float* __restrict__ y, int ldy,
                                                     int row_base, int n0, int Nout, int M, uint8_t* stg, int lane) {
  const int my_row = row_base + lane;
  const float sc_own = (ep.row_scale && my_row < M) ? __ldg(ep.row_scale + my_row) : 1.f;
  const int rsub = lane >> 3, c4 = lane & 7;
  float4* srow = reinterpret_cast<float4*>(stg + lane * 128);
  const int sw = lane & 7;
  constexpr int kChunkUnroll = (ACT == 2 || ACT == 4) ? 1 : BN / 32, kStepUnroll = (ACT == 2 || ACT == 4) ? 1 : 2;
  auto stage = [&](auto c0tag) {
    constexpr int c0 = decltype(c0tag)::value;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      srow[u ^ sw] = make_float4(acc[c0 + 4 * u], acc[c0 + 4 * u + 1], acc[c0 + 4 * u + 2], acc[c0 + 4 * u + 3]);
  };
#pragma unroll kChunkUnroll
  for (int ch = 0; ch < BN / 32; ++ch) {
    const int c0 = ch * 32;
    if (n0 + c0 >= Nout) break;
    switch (ch) {
      case 0: stage(std::integral_constant<int, 0>{}); break;
      case 1: if constexpr (BN > 32) stage(std::integral_constant<int, 32>{}); break;
      case 2: if constexpr (BN > 64) stage(std::integral_constant<int, 64>{}); break;
      default: if constexpr (BN > 96) stage(std::integral_constant<int, 96>{}); break;
    }
    __syncwarp();
    const int col = n0 + c0 + 4 * c4;
    const float4 bv = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll kStepUnroll
    for (int ib = 0; ib < 8; ib += 4) {               // 4 row steps at a time: their gate / residual loads are in flight together
      float4 g[4], r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = row_base + 4 * (ib + i) + rsub;
        g[i] = r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < M) {
          const size_t o = (size_t)row * ldy + col;
          if (ACT == 2 || ACT == 4) g[i] = *reinterpret_cast<const float4*>(ep.gate + o);
          if (ep.residual) r[i] = *reinterpret_cast<const float4*>(ep.residual + o);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = 4 * (ib + i) + rsub, row = row_base + rl;
        const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((c4 ^ (rl & 7)) << 4));
        const float sc = __shfl_sync(0xffffffffu, sc_own, rl);
        if (row < M) {
          float4 v;
          v.x = epilogue_op<ACT>(a.x, bv.x, g[i].x, sc, r[i].x);
          v.y = epilogue_op<ACT>(a.y, bv.y, g[i].y, sc, r[i].y);
          v.z = epilogue_op<ACT>(a.z, bv.z, g[i].z, sc, r[i].z);
          v.w = epilogue_op<ACT>(a.w, bv.w, g[i].w, sc, r[i].w);
          *reinterpret_cast<float4*>(y + (size_t)row * ldy + col) = v;
        }
      }
    }
    __syncwarp();
  }
}

// GLU epilogue (act 5): the tile's first BN/2 accumulator columns are projections, the last BN/2 their gates:
//   y[row, n0/2 + c] = (acc[c] + b[n0+c]) * sigmoid(acc[BN/2+c] + b[n0+BN/2+c]) * row_scale[row]
// (left/right projections and gates of TriangleMultiplication, seqformer.py:452-460, as one GEMM with the weight
// rows interleaved per tile); ldy is the leading dimension of the half-width output.
template <int BN>
__device__ __forceinline__ void store_row_glu(const float (&acc)[BN], const Epilogue& ep, float* __restrict__ y, int ldy,
                                              int row, int n0, int Nout) {
  constexpr int HB = BN / 2;
  if (n0 >= Nout) return;
  const float sc = ep.row_scale ? __ldg(ep.row_scale + row) : 1.f;
  if (ep.cm_n > 0) {                                // channel-major store: lanes = consecutive k -> coalesced per channel
    const long long n = ep.cm_n, nn = n * n, b = row / nn, r = row % nn, i = r / n, k = r % n;
    const long long C2 = Nout / 2;
    float* yc = y + (((b * C2 + n0 / 2) * n + i) * ep.cm_np + k);
    const long long cstride = n * ep.cm_np;
#pragma unroll
    for (int c = 0; c < HB; ++c) {
      const float pb = ep.bias ? __ldg(ep.bias + n0 + c) : 0.f, gb = ep.bias ? __ldg(ep.bias + n0 + HB + c) : 0.f;
      yc[c * cstride] = (acc[c] + pb) * sigmoid_fast(acc[HB + c] + gb) * sc;
    }
    return;
  }
  float* yr = y + (size_t)row * ldy + n0 / 2;
#pragma unroll
  for (int c = 0; c < HB; c += 4) {
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float pb = ep.bias ? __ldg(ep.bias + n0 + c + u) : 0.f, gb = ep.bias ? __ldg(ep.bias + n0 + HB + c + u) : 0.f;
      o[u] = (acc[c + u] + pb) * sigmoid_fast(acc[HB + c + u] + gb) * sc;
    }
    *reinterpret_cast<float4*>(yr + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_blo, int has_blo, int M, int Nout, int K, Epilogue ep, float* __restrict__ y, int ldy, int trust_trunc, int kb_per_drain,
                   int splits, long long split_stride, Batched bt) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto stage_a = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes; };
  auto stage_b = [&](int s) { return stage_a(s) + kTileABytes; };
  auto stage_blo = [&](int s) { return stage_b(s) + Cfg::kTileBBytes; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                        // TMA -> converters
  uint64_t* conv = bars + kStages;              // converters -> MMA
  uint64_t* empty = bars + 2 * kStages;         // MMA -> TMA
  // The TMEM partial-sum buffers form one ring shared by both accumulator warpgroups, but every (warpgroup, buffer) pair
  // has its own barriers: an mbarrier wait only carries a parity, so each barrier must be waited on use after use by a
  // single party (a warpgroup that skipped the other one's uses would pass on a stale phase).
  uint64_t* acc_ready = bars + 3 * kStages;     // [2 WGs][4] MMA -> accumulator WG (partial sum of one k-slab in TMEM)
  uint64_t* acc_free = bars + 3 * kStages + 8;  // [2 WGs][4] accumulator WG -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split-K: tile = (split, m block, n block); split s covers k-slabs [s kbs, min((s+1) kbs, total)) and writes its
  // raw partial product to y + s split_stride (the caller reduces); splits == 1 is the plain GEMM
  const int nkb_total = (K + kBK - 1) / kBK;
  const int kbs = (nkb_total + splits - 1) / splits;
  // batched mode: M and Nout are the per-problem sizes (bt.rows), the tile index also runs over the problems
  const int nt = (Nout + BN - 1) / BN, mt = (M + kBM - 1) / kBM;
  const int mnt = nt * mt * (bt.batches > 0 ? bt.batches : 1), tiles = mnt * splits;
  auto batch_base = [&](int bc) { return (bc / bt.inner) * bt.outer_rows + (bc % bt.inner) * bt.rows; };
  auto tile_nkb = [&](int tile) { const int k0 = (tile / mnt) * kbs; return min(kbs, nkb_total - k0); };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(conv + s, kConvThreads / 32);      // one arrival per converter warp
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 8; ++b) {
      mbar_init(acc_ready + b, 1);
      mbar_init(acc_free + b, kAccThreads / 32);   // one arrival per accumulator warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const bool prof = ABX_GEMM_PROFILE && (trust_trunc & 32) && blockIdx.x == 0;
  // programmatic dependent launch: the set-up above overlaps the tail of the previous kernel in the stream; nothing
  // it wrote is touched before this point (no-ops for an ordinary launch)
  griddep_wait();
  griddep_launch_dependents();

  // Every role walks the same static schedule: this CTA's j-th tile is blockIdx.x + j gridDim.x (n fastest, so
  // CTAs running side by side share the A rows in L2); `it` = j nkb + kb counts k-slabs and drives the smem
  // stages, the two TMEM partial-sum buffers and all barrier phases.  Tile j belongs to accumulator WG j & 1.
  const int wg = warp >> 2;
  if (wg == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
    if (warp == 0 && lane == 0) {
      // ---------------- TMA producer ----------------
      uint32_t it = 0;
      unsigned long long pw[2] = {0, 0};
      const long long tstart = clock64();
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int rem = tile % mnt, kb0 = (tile / mnt) * kbs, nkb = tile_nkb(tile);
        int m0 = ((rem / nt) % mt) * kBM, n0 = (rem % nt) * BN;
        if (bt.batches > 0) { const int base = batch_base(rem / (nt * mt)); m0 += base; n0 += base; }
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          ABX_GEMM_PWAIT(0, empty + s, ph ^ 1);
          if (trust_trunc & 16) { mbar_arrive(full + s); continue; }      // timing probe: no TMA at all
          mbar_expect_tx(full + s, kTileABytes + Cfg::kTileBBytes * (has_blo ? 2u : 1u));
          tma_load_2d(stage_a(s), &map_a, full + s, (kb0 + kb) * kBK, m0);
          tma_load_2d(stage_b(s), &map_b, full + s, (kb0 + kb) * kBK, n0);
          if (has_blo) tma_load_2d(stage_blo(s), &map_blo, full + s, (kb0 + kb) * kBK, n0);   // precomputed lo part of a static weight
        }
      }
      if (prof) { g_gemm_prof[0] = pw[0]; g_gemm_prof[1] = clock64() - tstart; g_gemm_prof[2] = it; }
    } else if ((warp == 1 || warp == 2) && lane == 0) {
      // ---------------- MMA issuers (two threads, alternating partial sums) ----------------
      // One thread cannot keep the tensor pipe fed: a 12-MMA slab (~1.1k clk of pipe time at the measured tf32 rate) blocks
      // the issuing thread for about as long (the MMA queue is shallow), and only then can it run the barrier waits of the next
      // slab (~200 clk each even when already complete) while the pipe drains.  Two threads issue alternate partial sums —
      // disjoint TMEM buffers and smem / TMEM-A stages, and tcgen05.commit tracks the issuing thread's own MMAs — so one
      // thread's waits overlap the other's MMAs.
      constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
      const uint32_t X = warp - 1;                    // this thread issues the partial sums with (index & 1) == X
      // Partial sums: `kb_per_drain` consecutive k-slabs are chained in one TMEM buffer (the first MMA of a group
      // overwrites), then handed to the accumulator WG, which adds them up in registers with round-to-nearest.
      uint32_t it = 0, pit = 0;                       // k-slabs / partial sums so far (ring of kAccBufs TMEM buffers)
      uint32_t used = 0, owner = 0, fpar = 0;         // per buffer: holds a partial / warpgroup of its last partial;
                                                      // fpar bit 4 g + b: parity of the next acc_free[g][b] completion
      int j = 0;
      unsigned long long pw[3] = {0, 0, 0};
      const long long tstart = clock64();
      auto a_addr = [&](int s) { return tmem_base + Cfg::kAColBase + (uint32_t)s * 2 * kBK; };
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++j) {
        const uint32_t g = j & 1;
        const int nkb = tile_nkb(tile);
        for (int kb0 = 0; kb0 < nkb; kb0 += kb_per_drain, ++pit) {
          const int cnt = min(kb_per_drain, nkb - kb0);
          const uint32_t it0 = it;
          it += cnt;
          if ((pit & 1) != X) continue;
          const uint32_t buf = pit % Cfg::kAccBufs;
          if ((used >> buf) & 1) {                    // the buffer's previous partial must have been drained by its owner
            const uint32_t idx = 4 * ((owner >> buf) & 1) + buf;
            ABX_GEMM_PWAIT(0, acc_free + idx, (fpar >> idx) & 1);
            fpar ^= 1u << idx;
          }
          used |= 1u << buf;
          owner = (owner & ~(1u << buf)) | (g << buf);
          const uint32_t tmem_acc = tmem_base + buf * BN;
          if (cnt == 2) {
            // two slabs per partial: every small cross term (hi*lo, lo*hi) of both slabs first, then the hi*hi terms — the
            // tensor core truncates when it adds into the accumulator, and the number of additions made at full magnitude
            // (4 per slab) stays what it is with one slab per partial
            const int s0 = it0 % kStages, s1 = (it0 + 1) % kStages;
            ABX_GEMM_PWAIT(1, conv + s0, (it0 / kStages) & 1);
            ABX_GEMM_PWAIT(1, conv + s1, ((it0 + 1) / kStages) & 1);
            const long long tissue = prof ? clock64() : 0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!(trust_trunc & 4)) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int s = u ? s1 : s0;
                const uint32_t a_hi = a_addr(s), a_lo = a_hi + kBK;
                const uint64_t b_hi = umma_desc_sw128(smem_u32(stage_b(s))), b_lo = umma_desc_sw128(smem_u32(stage_blo(s)));
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {
                  umma_tf32_ts(tmem_acc, a_hi + 8 * k, b_lo + 2 * k, idesc, (u != 0 || k != 0) ? 1u : 0u);
                  umma_tf32_ts(tmem_acc, a_lo + 8 * k, b_hi + 2 * k, idesc, 1);
                }
              }
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int s = u ? s1 : s0;
                const uint32_t a_hi = a_addr(s);
                const uint64_t b_hi = umma_desc_sw128(smem_u32(stage_b(s)));
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) umma_tf32_ts(tmem_acc, a_hi + 8 * k, b_hi + 2 * k, idesc, 1);
                umma_commit(empty + s);               // smem slab and TMEM A stage free once the MMAs so far have read them
              }
            } else {
              umma_commit(empty + s0);
              umma_commit(empty + s1);
            }
            umma_commit(acc_ready + 4 * g + buf);     // partial sum complete
            if (prof) pw[2] += clock64() - tissue;
          } else {
            for (int u = 0; u < cnt; ++u) {           // any other group size: slab by slab (small cross terms first, then hi*hi)
              const uint32_t iu = it0 + u;
              const int s = iu % kStages;
              ABX_GEMM_PWAIT(1, conv + s, (iu / kStages) & 1);
              const long long tissue = prof ? clock64() : 0;
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t a_hi = a_addr(s), a_lo = a_hi + kBK;
              const uint64_t b_hi = umma_desc_sw128(smem_u32(stage_b(s))), b_lo = umma_desc_sw128(smem_u32(stage_blo(s)));
              if (!(trust_trunc & 4)) {
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {
                  umma_tf32_ts(tmem_acc, a_hi + 8 * k, b_lo + 2 * k, idesc, (u != 0 || k != 0) ? 1u : 0u);
                  umma_tf32_ts(tmem_acc, a_lo + 8 * k, b_hi + 2 * k, idesc, 1);
                }
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) umma_tf32_ts(tmem_acc, a_hi + 8 * k, b_hi + 2 * k, idesc, 1);
              }
              umma_commit(empty + s);
              if (u + 1 == cnt) umma_commit(acc_ready + 4 * g + buf);
              if (prof) pw[2] += clock64() - tissue;
            }
          }
        }
      }
      if (prof) {
        unsigned long long* o = g_gemm_prof + (X ? 20 : 4);
        o[0] = pw[0]; o[1] = pw[1]; o[2] = clock64() - tstart; o[3] = pw[2];
      }
    }
  } else if (wg == 1) {
    // ---------------- converters: raw fp32 slab -> lo tile (hi = the raw tile, low bits ignored) ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsConv));
    const int ct = threadIdx.x - 128;                // 0..127
    uint32_t it = 0;
    unsigned long long pw[1] = {0};
    const long long tstart = clock64();
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int nkb = tile_nkb(tile);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        ABX_GEMM_PWAIT(0, full + s, ph);
        if (trust_trunc & 2) { __syncwarp(); if (lane == 0) mbar_arrive(conv + s); continue; }   // timing probe: no conversion
        // A: thread = row ct of the 128-byte-swizzled tile (16-byte chunk c of row r sits at chunk c ^ (r & 7)); the stage's
        // TMEM columns are free: `full` was armed only after the MMAs of the stage's previous use had committed `empty`
        {
          const uint8_t* arow = stage_a(s) + ct * 128;
          const uint32_t ta = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + Cfg::kAColBase + (uint32_t)s * 2 * kBK;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 v = *reinterpret_cast<const float4*>(arow + (((4 * half + c) ^ (ct & 7)) << 4));
              const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                hi[4 * c + u] = __float_as_uint(x[u]) & 0xffffe000u;
                lo[4 * c + u] = __float_as_uint(x[u] - __uint_as_float(hi[4 * c + u]));
              }
            }
            tmem_st16(ta + 16 * half, hi);
            tmem_st16(ta + kBK + 16 * half, lo);
          }
        }
        // W: lo tile next to the raw tile (element-wise, layout preserved) — unless the caller supplied it (static weights)
        if (!has_blo) {
          const float4* b = reinterpret_cast<const float4*>(stage_b(s));
          float4* blo = reinterpret_cast<float4*>(stage_blo(s));
#pragma unroll 8
          for (int i = ct; i < (int)(Cfg::kTileBBytes / 16); i += kConvThreads) {
            const float4 v = b[i];
            float4 l;
            l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            blo[i] = l;
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();                               // every lane's lo-tile writes are fenced before the warp's single arrival
        if (lane == 0) mbar_arrive(conv + s);
      }
    }
    if (prof && ct == 0) { g_gemm_prof[8] = pw[0]; g_gemm_prof[9] = clock64() - tstart; }
  } else {
    // ---------------- accumulator / epilogue warpgroups (ping-pong on tiles): lane = one output row ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsAcc));
    const int g = wg - 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const bool vec_ok = (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                        (!ep.bias || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) &&
                        (!ep.residual || (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0) &&
                        (!ep.gate || (reinterpret_cast<uintptr_t>(ep.gate) & 15) == 0);
    // position of a tile's first partial sum in the shared ring = partial sums of all earlier tiles of this CTA
    auto ring_pos = [&](int tile) -> uint32_t {
      const int j = (tile - (int)blockIdx.x) / (int)gridDim.x;
      if (splits == 1) return (uint32_t)(j * ((nkb_total + kb_per_drain - 1) / kb_per_drain));
      uint32_t p = 0;
      for (int t2 = blockIdx.x; t2 < tile; t2 += gridDim.x) p += (tile_nkb(t2) + kb_per_drain - 1) / kb_per_drain;
      return p;
    };
    uint32_t rpar = 0;                               // bit b: parity of the next acc_ready[g][b] completion
    unsigned long long pw[3] = {0, 0, 0};
    const long long tstart = clock64();
    for (int tile = blockIdx.x + g * gridDim.x; tile < tiles; tile += 2 * gridDim.x) {
      uint32_t pit = ring_pos(tile);
      const int rem = tile % mnt, nkb = tile_nkb(tile);
      const int m0 = ((rem / nt) % mt) * kBM, n0 = (rem % nt) * BN;
      float* __restrict__ yt = y + (long long)(tile / mnt) * split_stride +
                               (bt.batches > 0 ? (long long)(rem / (nt * mt)) * bt.rows * ldy : 0);
      const int row = m0 + 32 * q + lane;
      if (row < M) {                                 // pull this row's gate / residual pieces into L2 ahead of the epilogue
        const int ncol = min(BN, Nout - n0);
        if (ep.gate)
          for (int c = 0; c < ncol; c += 32) prefetch_l2(ep.gate + (size_t)row * ldy + n0 + c);
        if (ep.residual && ep.transpose_n == 0)
          for (int c = 0; c < ncol; c += 32) prefetch_l2(ep.residual + (size_t)row * ldy + n0 + c);
      }
      float acc[BN];
#pragma unroll
      for (int c = 0; c < BN; ++c) acc[c] = 0.f;
      const int ngroups = (nkb + kb_per_drain - 1) / kb_per_drain;
      for (int gi = 0; gi < ngroups; ++gi, ++pit) {
        const uint32_t buf = pit % Cfg::kAccBufs;
        ABX_GEMM_PWAIT(0, acc_ready + 4 * g + buf, (rpar >> buf) & 1);
        rpar ^= 1u << buf;
        const long long tdrain = prof ? clock64() : 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (trust_trunc & 8) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + buf * BN + (uint32_t)c0, v);
#pragma unroll
          for (int c = 0; c < 32; ++c) acc[c0 + c] += __uint_as_float(v[c]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free + 4 * g + buf);
        if (prof) pw[1] += clock64() - tdrain;
      }
      const long long tepi = prof ? clock64() : 0;
      if (vec_ok && ep.transpose_n == 0 && ep.act != 5 && Nout % 32 == 0) {
        uint8_t* stg = smem + (size_t)kStages * Cfg::kStageBytes + 512 + (size_t)(4 * g + q) * 4096;
        const int rb = m0 + 32 * q;
        switch (ep.act) {
          case 0: store_tile_coalesced<0, BN>(acc, ep, yt, ldy, rb, n0, Nout, M, stg, lane); break;
          case 1: store_tile_coalesced<1, BN>(acc, ep, yt, ldy, rb, n0, Nout, M, stg, lane); break;
          case 2: store_tile_coalesced<2, BN>(acc, ep, yt, ldy, rb, n0, Nout, M, stg, lane); break;
          case 3: store_tile_coalesced<3, BN>(acc, ep, yt, ldy, rb, n0, Nout, M, stg, lane); break;
          default: store_tile_coalesced<4, BN>(acc, ep, yt, ldy, rb, n0, Nout, M, stg, lane); break;
        }
      } else if (row < M) {
        switch (ep.act) {
          case 0: store_row<0, BN>(acc, ep, yt, ldy, row, n0, Nout, vec_ok); break;
          case 1: store_row<1, BN>(acc, ep, yt, ldy, row, n0, Nout, vec_ok); break;
          case 2: store_row<2, BN>(acc, ep, yt, ldy, row, n0, Nout, vec_ok); break;
          case 3: store_row<3, BN>(acc, ep, yt, ldy, row, n0, Nout, vec_ok); break;
          case 4: store_row<4, BN>(acc, ep, yt, ldy, row, n0, Nout, vec_ok); break;
          default: if constexpr (BN == 128) store_row_glu<BN>(acc, ep, yt, ldy, row, n0, Nout); break;
        }
      }
      if (prof) pw[2] += clock64() - tepi;
    }
    if (prof && (threadIdx.x & 127) == 0) {
      unsigned long long* o = g_gemm_prof + 12 + 4 * g;
      o[0] = pw[0]; o[1] = pw[1]; o[2] = pw[2]; o[3] = clock64() - tstart;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// [rows, K] fp32 row-major (row stride ld elements) -> tiles of box_rows x 32 elements, 128-byte swizzle,
// out-of-bounds elements read as zero (handles the M / Nout / K tails)
int make_map(CUtensorMap* map, const float* base, int rows, int K, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  ABX_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, kBK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ABX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a [%d,%d] ld=%d operand", (int)r, rows, K, ld);
  return ABX_OK;
}

int sm_count() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;
    return v;
  }();
  return n;
}

// Run-time switches of the kernel (the `trust_trunc` argument, named after its first use).  The tensor core ignores the 13 low
// mantissa bits of a tf32 operand read from shared memory (verified bit for bit on B200 in round 1 with an explicit-hi-write-back build of this kernel), so the raw
// W tile already acts as its hi part and only the lo tile is written; the A operand in tensor memory is masked explicitly.
int trust_trunc() {
  static int v = [] {
    const char* d = getenv("ABX_GEMM_DEBUG_SKIP");      // timing experiments only: 2 skip split, 4 skip MMA, 8 skip drain, 16 skip TMA
    const char* pf = getenv("ABX_GEMM_PROF");           // per-role cycle counters of CTA 0 (abx_gemm_profile)
    return 1 | (d ? (atoi(d) & 30) : 0) | ((pf && pf[0] == '1') ? 32 : 0);
  }();
  return v;
}

// k-slabs (of 32) chained in tensor memory between two register accumulations: 1 = every slab (the tensor
// core's truncating accumulator never chains more than 12 MMAs), larger values trade a little rounding bias for
// fewer TMEM reads.  ABX_GEMM_KB_PER_DRAIN overrides the default.
int kb_per_drain() {
  static int v = [] { const char* e = getenv("ABX_GEMM_KB_PER_DRAIN"); int n = e ? atoi(e) : kDrainDefault; return n < 1 ? 1 : (n > 64 ? 64 : n); }();
  return v;
}

// ABX_GEMM_TIMING=1 (diagnostics, never inside a CUDA-graph capture): every launch is bracketed by events and synchronised;
// a per-shape table (calls, total ms) is printed to stderr at exit — where the dense-layer time of a sampler step goes.
struct GemmTimingTable {
  std::map<std::string, std::pair<double, long>> rows;
  ~GemmTimingTable() {
    if (rows.empty()) return;
    double tot = 0;
    for (auto& r : rows) tot += r.second.first;
    fprintf(stderr, "abx_gemm_tf32x3 timing (ms total %.3f)\n", tot);
    for (auto& r : rows)
      fprintf(stderr, "  %-58s calls %6ld  ms %10.3f  share %5.1f%%  ms/call %8.4f\n", r.first.c_str(), r.second.second, r.second.first,
              100.0 * r.second.first / tot, r.second.first / r.second.second);
  }
};
bool gemm_timing_on() {
  static bool on = [] { const char* e = getenv("ABX_GEMM_TIMING"); return e && e[0] == '1'; }();
  return on;
}
GemmTimingTable& gemm_timing_table() {
  static GemmTimingTable t;
  return t;
}

template <int BN>
int launch_bn(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw, const Epilogue& ep,
              float* y, int ldy, int splits = 1, long long split_stride = 0, Batched bt = Batched{0, 0, 1, 0},
              int map_rows_a = 0, int map_rows_b = 0, const float* w_lo = nullptr) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ma, mb, mblo;
  int rc;
  if ((rc = make_map(&ma, x, map_rows_a ? map_rows_a : M, K, ldx, kBM))) return rc;
  if ((rc = make_map(&mb, w, map_rows_b ? map_rows_b : Nout, K, ldw, BN))) return rc;
  if ((rc = make_map(&mblo, w_lo ? w_lo : w, map_rows_b ? map_rows_b : Nout, K, ldw, BN))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ABX_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = ceil_div(Nout, BN) * ceil_div(M, kBM) * splits * (bt.batches > 0 ? bt.batches : 1);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (gemm_timing_on()) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
  }
  const cudaError_t le = launch_kernel(gemm_tf32x3_kernel<BN>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, s, ma, mb, mblo,
                                       w_lo ? 1 : 0, M, Nout, K, ep, y, ldy, trust_trunc(), kb_per_drain(), splits, split_stride, bt);
  if (e0) {
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    char key[160];
    snprintf(key, sizeof key, "M=%d N=%d K=%d bn=%d act=%d%s%s%s%s%s%s batches=%d splits=%d", M, Nout, K, BN, ep.act, ep.bias ? " bias" : "",
             ep.residual ? " res" : "", ep.gate ? " gate" : "", ep.row_scale ? " scale" : "", ep.transpose_n ? " transp" : "",
             ep.cm_n ? " cm" : "", bt.batches, splits);
    auto& r = gemm_timing_table().rows[key];
    r.first += ms;
    r.second += 1;
  }
  count_launch();
  if (le != cudaSuccess) {
    set_error("launch of gemm_tf32x3_kernel failed: %s", cudaGetErrorString(le));
    return ABX_ERR_CUDA;
  }
  return check_launch("gemm_tf32x3_kernel");
}

}  // namespace

bool gemm_tf32x3_supported(int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw) {
  return M > 0 && Nout > 0 && K >= 4 && K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && ldx >= K && ldw >= K &&
         (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0;
}

// act: see Epilogue.  tile_n: 0 = choose, else 32/64/128.
int launch_gemm_tf32x3(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                       const float* bias, const float* residual, const float* gate, const float* row_scale, int act,
                       int transpose_n, float* y, int ldy, int tile_n, int cm_n = 0, int cm_np = 0, const float* w_lo = nullptr) {
  Epilogue ep{bias, residual, gate, row_scale, act, transpose_n, cm_n, cm_np};
  if (tile_n == 0) {
    // enough CTAs to cover the 148 SMs beats wide tiles for the small-M node GEMMs
    // the main loop costs about the same per k-slab for every tile width (barrier round trips, not bytes, pace
    // it), so the widest tile that is not mostly padding wins even when it leaves SMs idle
    if (Nout <= 32) tile_n = 32;
    else if (Nout <= 64) tile_n = 64;
    else tile_n = 128;
  }
  switch (tile_n) {
    case 32: return launch_bn<32>(s, M, Nout, K, x, ldx, w, ldw, ep, y, ldy, 1, 0, Batched{0, 0, 1, 0}, 0, 0, w_lo);
    case 64: return launch_bn<64>(s, M, Nout, K, x, ldx, w, ldw, ep, y, ldy, 1, 0, Batched{0, 0, 1, 0}, 0, 0, w_lo);
    case 128: return launch_bn<128>(s, M, Nout, K, x, ldx, w, ldw, ep, y, ldy, 1, 0, Batched{0, 0, 1, 0}, 0, 0, w_lo);
  }
  set_error("abx_gemm_tf32x3: tile_n must be 0, 32, 64 or 128 (got %d)", tile_n);
  return ABX_ERR_INVALID;
}

// Split-K form for skinny problems (few output tiles, long K — IPA's final projection): `splits` raw partial
// products x W^T restricted to consecutive K ranges are written to partials[s][M][Nout]; the caller sums them.
int launch_gemm_tf32x3_splitk(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                              int* splits_io, float* partials, int tile_n) {
  int splits = *splits_io;
  const int nkb = ceil_div(K, kBK);
  if (splits > nkb) splits = nkb;
  while (splits > 1 && (splits - 1) * ceil_div(nkb, splits) >= nkb) --splits;   // no empty split
  *splits_io = splits;
  Epilogue ep{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
  const long long stride = (long long)M * Nout;
  switch (tile_n) {
    case 32: return launch_bn<32>(s, M, Nout, K, x, ldx, w, ldw, ep, partials, Nout, splits, stride);
    case 64: return launch_bn<64>(s, M, Nout, K, x, ldx, w, ldw, ep, partials, Nout, splits, stride);
    default: return launch_bn<128>(s, M, Nout, K, x, ldx, w, ldw, ep, partials, Nout, splits, stride);
  }
}

// out[bc][i][j] = sum_k a[bc][i][k] b[bc][j][k] for bc < batches; a / b rows of problem bc start at row
// (bc / inner) * outer_rows + (bc % inner) * n of their [*, kpad] matrices (kpad >= n, zero padded), out [batches][n][ldo].
int launch_gemm_tf32x3_batched_nt(cudaStream_t s, int batches, int n, int kpad, int inner, int outer_rows, int total_rows,
                                  const float* a, const float* b, float* out, int ldo) {
  Epilogue ep{nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0};
  Batched bt{batches, n, inner, outer_rows};
  return launch_bn<128>(s, n, n, kpad, a, kpad, b, kpad, ep, out, ldo, 1, 0, bt, total_rows, total_rows);
}

}  // namespace abx

// Diagnostics: copies the 32 per-role cycle counters the last ABX_GEMM_PROF=1 launch recorded for CTA 0
// ([0] TMA wait-empty, [1] TMA total, [2] k-slabs; [4] issuer wait-acc_free, [5] wait-conv, [6] total, [7] issue;
//  [8] converter wait-full, [9] total; [12 + 4 g ..] accumulator WG g: wait-ready, drain, epilogue, total).
extern "C" int abx_gemm_profile(unsigned long long* out32) {
  ABX_REQUIRE(out32 != nullptr, "abx_gemm_profile: null output");
  ABX_CUDA(cudaDeviceSynchronize());
  ABX_CUDA(cudaMemcpyFromSymbol(out32, abx::g_gemm_prof, 32 * sizeof(unsigned long long)));
  return ABX_OK;
}

extern "C" int abx_gemm_tf32x3_batched_nt(void* stream, int batches, int n, int kpad, int inner, int outer_rows, int total_rows,
                                          const float* a, const float* b, float* out, int ldo) {
  ABX_REQUIRE(batches > 0 && n > 0 && a && b && out, "abx_gemm_tf32x3_batched_nt: bad shape or null argument");
  ABX_REQUIRE(kpad % 4 == 0 && kpad >= n && ldo % 4 == 0 && ldo >= n && inner > 0,
              "abx_gemm_tf32x3_batched_nt: kpad and ldo must be multiples of 4 and >= n");
  ABX_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)out % 16 == 0),
              "abx_gemm_tf32x3_batched_nt: operands must be 16-byte aligned");
  return abx::launch_gemm_tf32x3_batched_nt((cudaStream_t)stream, batches, n, kpad, inner, outer_rows, total_rows, a, b, out, ldo);
}

extern "C" int abx_gemm_tf32x3_wlo(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w, const float* w_lo,
                                   int ldw, const float* bias, const float* residual, const float* gate, const float* row_scale,
                                   int act, int transpose_n, float* y, int ldy, int tile_n) {
  ABX_REQUIRE(M > 0 && Nout > 0 && K > 0 && x && w && y, "abx_gemm_tf32x3: bad shape or null argument");
  ABX_REQUIRE(abx::gemm_tf32x3_supported(M, Nout, K, x, ldx, w, ldw),
              "abx_gemm_tf32x3: K, ldx, ldw must be multiples of 4 with ldx, ldw >= K, and x, w 16-byte aligned "
              "(K=%d ldx=%d ldw=%d)", K, ldx, ldw);
  ABX_REQUIRE(act >= 0 && act <= 5 && ((act != 2 && act != 4) || gate), "abx_gemm_tf32x3: bad activation / missing gate");
  if (act == 5) {
    ABX_REQUIRE(Nout % 128 == 0 && ldy >= Nout / 2 && ldy % 4 == 0 && (uintptr_t)y % 16 == 0 && !residual && !transpose_n,
                "abx_gemm_tf32x3: the GLU epilogue needs Nout %% 128 == 0, ldy >= Nout/2 (multiple of 4), aligned y, no residual");
    tile_n = 128;
  } else {
    ABX_REQUIRE(ldy >= Nout, "abx_gemm_tf32x3: ldy < Nout");
  }
  ABX_REQUIRE(transpose_n >= 0 && (transpose_n == 0 || M % (transpose_n * transpose_n) == 0),
              "abx_gemm_tf32x3: M must be a multiple of transpose_n^2");
  ABX_REQUIRE(w_lo == nullptr || ((uintptr_t)w_lo & 15) == 0, "abx_gemm_tf32x3: w_lo must be 16-byte aligned");
  return abx::launch_gemm_tf32x3((cudaStream_t)stream, M, Nout, K, x, ldx, w, ldw, bias, residual, gate, row_scale, act,
                                 transpose_n, y, ldy, tile_n, 0, 0, w_lo);
}

extern "C" int abx_gemm_tf32x3(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                               const float* bias, const float* residual, const float* gate, const float* row_scale,
                               int act, int transpose_n, float* y, int ldy, int tile_n) {
  return abx_gemm_tf32x3_wlo(stream, M, Nout, K, x, ldx, w, nullptr, ldw, bias, residual, gate, row_scale, act, transpose_n, y, ldy,
                             tile_n);
}

extern "C" int abx_gemm_tf32x3_glu_cm_wlo(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w,
                                          const float* w_lo, int ldw, const float* bias, const float* row_scale, int n, int np,
                                          float* y) {
  ABX_REQUIRE(M > 0 && Nout > 0 && K > 0 && x && w && y, "abx_gemm_tf32x3_glu_cm: bad shape or null argument");
  ABX_REQUIRE(abx::gemm_tf32x3_supported(M, Nout, K, x, ldx, w, ldw), "abx_gemm_tf32x3_glu_cm: unsupported operand layout");
  ABX_REQUIRE(Nout % 128 == 0 && n > 0 && np >= n && M % (n * n) == 0, "abx_gemm_tf32x3_glu_cm: Nout %% 128, M %% n^2 and np >= n required");
  ABX_REQUIRE(w_lo == nullptr || ((uintptr_t)w_lo & 15) == 0, "abx_gemm_tf32x3_glu_cm: w_lo must be 16-byte aligned");
  return abx::launch_gemm_tf32x3((cudaStream_t)stream, M, Nout, K, x, ldx, w, ldw, bias, nullptr, nullptr, row_scale, 5, 0, y,
                                 Nout / 2, 128, n, np, w_lo);
}

extern "C" int abx_gemm_tf32x3_glu_cm(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                                      const float* bias, const float* row_scale, int n, int np, float* y) {
  return abx_gemm_tf32x3_glu_cm_wlo(stream, M, Nout, K, x, ldx, w, nullptr, ldw, bias, row_scale, n, np, y);
}
