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
  float *proj, *Qp, *KVp, *feats, *bias, *partials;
  size_t total;
};

// The GEMM main loop costs about the same per 32-wide k-slab whatever the tile width, so the final projection
// uses the widest tiles (128 columns: 2 per row block) and splits K so that every SM gets at most one tile.
static int final_proj_splits(int M) {
  const int tiles = ceil_div(M, 128) * (kC / 128);
  int s = 148 / tiles;
  return s < 1 ? 1 : (s > kMaxSplits ? kMaxSplits : s);
}

static IpaWorkspace carve(void* base, int B, int N, bool with_bias, bool with_feats) {
  IpaWorkspace w;
  size_t off = 0;
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(p + off); off += align_up(floats * sizeof(float)); return r; };
  const size_t bn = (size_t)B * N;
  w.proj = take(bn * kProj);
  w.Qp = take(ipa_fused_qp_floats(B, N));                // packed queries [B,N,12,28]
  w.KVp = take(ipa_fused_kvp_floats(B, N));              // packed keys + values [B,N,816]
  w.feats = with_feats ? take(bn * kFeat) : nullptr;
  w.partials = (with_feats && final_proj_splits((int)bn) > 1) ? take((size_t)final_proj_splits((int)bn) * bn * kC) : nullptr;
  w.bias = with_bias ? take(ipa_pair_bias_floats(B, N)) : nullptr;
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
  if (pair_bias == nullptr) {
    ABX_REQUIRE(ws.bias != nullptr, "ipa: no pair bias and no workspace room for it");
    if ((rc = launch_ipa_pair_bias(s, B, N, z, w->w_pair, w->b_pair, ws.bias))) return rc;
    pair_bias = ws.bias;
  }
  if ((rc = launch_ipa_pack_nodes(s, B, N, ws.proj, rots, trans, ws.Qp, ws.KVp))) return rc;
  return launch_ipa_fused(s, B, N, ws.Qp, ws.KVp, pair_bias, mask, rots, trans, w->point_weights, z, feats);
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
  IpaWorkspace ws = carve(workspace, B, N, pair_bias == nullptr, false);
  ABX_REQUIRE(workspace_bytes >= ws.total, "abx_ipa_attention_features: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
  return ipa_features((cudaStream_t)stream, B, N, x, z, mask, rots, trans, w, pair_bias, feats, ws);
}

extern "C" int abx_ipa_forward(void* stream, int B, int N, const float* x, const float* z, const float* mask,
                               const float* rots, const float* trans, const abx_ipa_weights* w, const float* pair_bias,
                               const float* residual, float* out, void* workspace, size_t workspace_bytes) {
  int rc = ipa_check("abx_ipa_forward", B, N, x, z, mask, rots, trans, w);
  if (rc) return rc;
  ABX_REQUIRE(out && workspace, "abx_ipa_forward: null output or workspace");
  IpaWorkspace ws = carve(workspace, B, N, pair_bias == nullptr, true);
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
