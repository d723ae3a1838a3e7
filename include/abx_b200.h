/* abx_b200 — C ABI of the B200 (sm_100a) kernels behind AbX's reverse-diffusion sampling path.
 *
 * The reference (CarbonMatrixLab/AbX) is pure Python/PyTorch and has no FFI of its own, so the
 * drop-in boundary is its Python API; this library sits directly underneath the
 * Python classes that mirror that API (abx_b200/diffuser/*.py, abx_b200/model/*.py).  Every entry
 * point names the reference code (file:line under the AbX checkout) whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is DEVICE memory unless stated otherwise
 *   - the caller owns every buffer, including outputs and workspaces (kernels never allocate)
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no internal syncs
 *   - tensors are contiguous row-major with the shapes given per function
 *   - return value: 0 = ok, non-zero = error (message via abx_last_error()); never throws
 *   - no global mutable state besides the launch counter and the last-error string (thread-local)
 */
#ifndef ABX_B200_H_
#define ABX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABX_OK 0
#define ABX_ERR_INVALID 1   /* bad argument (shape, null pointer, workspace too small) */
#define ABX_ERR_CUDA 2      /* a CUDA runtime call failed */

/* ---- library -------------------------------------------------------------------------------- */
const char* abx_last_error(void);          /* message of the last failing call on this thread   */
int abx_version(void);                     /* ABI version, bumps on any signature change        */
uint64_t abx_launch_count(void);           /* kernels launched by this library since load/reset */
void abx_reset_launch_count(void);
/* ABI checks for a device with compute capability 10.x; fills sm count.  Host call. */
int abx_device_check(int device, int* sm_count);

/* ---- SE(3) / categorical diffusers ---------------------------------------------------------- */
/* Scalars of the noise schedules, computed on the HOST by the Python layer exactly as torch does
 * (float32 0-d tensors promoted against t), so device arithmetic matches the reference bit for bit
 * in the t-dependent factors.  diffuser/so3_diffuser.py:198-216, r3_diffuser.py:29-46,150-164. */
typedef struct {
  double so3_exp_max;      /* (double)float32(exp(max_sigma))                                   */
  double so3_exp_min;      /* (double)float32(exp(min_sigma))                                   */
  double so3_g2_coef;      /* (double)float32(2*(exp(max_sigma)-exp(min_sigma)))                */
  double r3_min_b;         /* (double)float32(min_b)                                            */
  double r3_delta_b;       /* (double)float32(max_b-min_b)                                      */
  double r3_coord_scale;   /* (double)float32(coordinate_scaling)                               */
  double seq_rate;         /* CTMC off-diagonal rate (rate_const)                               */
  int num_sigma;           /* rows of the IGSO(3) tables                                        */
  int num_omega;           /* columns of the IGSO(3) tables                                     */
} abx_diffuser_consts;

/* rot_score/trans_score of a predicted x0 — replaces FullDiffuser.calc_quat_score
 * (diffuser/full_diffuser.py:135-142 -> quat_affine.py:76-131,234-238 -> so3_diffuser.py:264-297,
 * cached-table branch) and FullDiffuser.calc_trans_score (full_diffuser.py:131-133 ->
 * r3_diffuser.py:158-164).
 *   quat_t, quat_0 [B,N,4] f32; trans_t, trans_0 [B,N,3] f32 (Angstrom); t [B] f64
 *   t_is_f32: the reference was handed a float32 t (warm-up call, inference.py:210): sigma(t) and
 *             the translation score are then evaluated in float32
 *   score_norms [num_sigma,num_omega] f32; discrete_sigma [num_sigma] f32; discrete_omega [num_omega] f32
 *   rot_score [B,N,3] f32 (may be NULL);  trans_score [B,N,3] f64 (f32 when t_is_f32; may be NULL) */
int abx_se3_scores(void* stream, int B, int N, const abx_diffuser_consts* c,
                   const float* quat_t, const float* quat_0, const float* trans_t, const float* trans_0,
                   const double* t, int t_is_f32,
                   const float* score_norms, const float* discrete_sigma, const float* discrete_omega,
                   float* rot_score, void* trans_score);

/* SO3Diffuser.score on a rotation vector (so3_diffuser.py:264-297, cached branch), used by
 * FullDiffuser.score / forward_marginal.   rotvec [B,N,3] f32 -> rot_score [B,N,3] f32 */
int abx_so3_score_rotvec(void* stream, int B, int N, const abx_diffuser_consts* c, const float* rotvec,
                         const double* t, int t_is_f32, const float* score_norms, const float* discrete_sigma,
                         const float* discrete_omega, float* rot_score);

/* Live (uncached) IGSO(3) score: the truncated character series evaluated per residue, one warp per
 * residue with warp-shuffle reductions — so3_diffuser.py:72-112 + :290-295 (use_cached_score=False).
 *   rotvec [B,N,3] f32 -> rot_score [B,N,3] f32;  L = series length (reference: 1000) */
int abx_igso3_score_series(void* stream, int B, int N, const abx_diffuser_consts* c, const float* rotvec,
                           const double* t, int t_is_f32, const float* discrete_sigma, int L, float* rot_score);

/* Poisson rates of the categorical tau-leap, rate*dt — discrete_diffuser.py:130-179 (softmax,
 * transition :53-67, reverse rates).   seq_t [B,N] i64; logits [B,N,20] f32; t [B] f64; dt scalar
 *   rate_dt [B,N,20] f32 (the tensor the reference hands to torch.poisson) */
int abx_seq_reverse_rates(void* stream, int B, int N, const abx_diffuser_consts* c, const int64_t* seq_t,
                          const float* logits, const double* t, double dt, float* rate_dt);

/* One Euler-Maruyama reverse step on SE(3)^N x {0..19}^N — replaces FullDiffuser.reverse
 * (full_diffuser.py:174-227): so3_diffuser.py:328-361 (geodesic step), r3_diffuser.py:110-148
 * (VP-SDE step incl. the g*dt*z noise term and centre of mass over all N), discrete_diffuser.py:181-190
 * (apply jumps), masking in rotvec space (:219-225), _assemble_rigid (:20-26).
 *   rigid_t [B,N,7] (f32 or f64: rigid_is_f64) = [qw,qx,qy,qz,tx,ty,tz]; seq_t [B,N] i64
 *   rot_score [B,N,3] f32; trans_score [B,N,3] f64; diffuse_mask [B,N] i32 (NULL = all ones)
 *   z_rot, z_trans [B,N,3] f32 (the two randn draws, reference order); jumps [B,N,20] f32 (Poisson draw)
 *   flags: bit0 diffuse_rot, bit1 diffuse_trans, bit2 diffuse_seq, bit3 center
 *   rigids_out [B,N,7] f64; seq_out [B,N] i64.  One CTA per batch element. */
int abx_se3_reverse_step(void* stream, int B, int N, const abx_diffuser_consts* c,
                         const void* rigid_t, int rigid_is_f64, const int64_t* seq_t,
                         const float* rot_score, const double* trans_score, const int32_t* diffuse_mask,
                         const double* t, double dt_f32, double sqrt_dt_f32, double noise_scale,
                         const float* z_rot, const float* z_trans, const float* jumps, int flags,
                         double* rigids_out, int64_t* seq_out);

/* IGSO(3) tables (pdf, cdf, score norms over a sigma x omega grid) — replaces the cache build in
 * SO3Diffuser.__init__ (so3_diffuser.py:150-166; igso3_expansion :15-49, density :52-69, score :72-112).
 *   discrete_sigma [num_sigma] f32, discrete_omega [num_omega] f32 -> three [num_sigma,num_omega] f32 */
int abx_igso3_build_tables(void* stream, int num_sigma, int num_omega, int L, const float* discrete_sigma,
                           const float* discrete_omega, float* pdf, float* cdf, float* score_norms);

/* ---- dense node GEMM ------------------------------------------------------------------------- */
/* y[M,Nout] = act(x[M,K] @ w[Nout,K]^T + bias) (+ residual) — torch.nn.Linear semantics
 * (abx/model/common_modules.py:11-59).  fp32 in/out, fp32-accurate accumulation.
 *   bias, residual may be NULL; relu: 0/1; ldx/ldy: row strides in elements (>= K / >= Nout). */
int abx_linear_f32(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w,
                   const float* bias, const float* residual, int relu, float* y, int ldy);

/* Same layer on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM accumulator, TMA-staged
 * operands) with fp32-level accuracy: every operand is split exactly into TF32 hi + lo parts and
 * D = A_hi W_lo + A_lo W_hi + A_hi W_hi is accumulated in fp32 (3xTF32).  Used for the node GEMMs of IPA
 * (folding.py:69-86,130-132), IpaScore (score_network.py:117-137) and the trunk's dense layers
 * (seqformer.py).   w [Nout, ldw] row-major (nn.Linear layout when ldw == K).
 *   v = acc + bias;  act: 0 none, 1 relu(v), 2 v * sigmoid(gate), 3 sigmoid(v), 4 sigmoid(v) * gate,
 *   5 GLU: per 128-column tile the first 64 columns are projections and the last 64 their gates,
 *     y[:, n0/2 + c] = v[n0+c] * sigmoid(v[n0+64+c]) (y has Nout/2 columns; seqformer.py:452-460 as one GEMM)
 *   (gate [M,ldy]);  then y = v * row_scale[row] + residual   (row_scale [M], residual [M,ldy]; all optional)
 *   — the gated projections, masked projections and residual adds of seqformer.py fused into the GEMM
 *   transpose_n = n > 0: the M rows are (b,i,j) of a [B,n,n,*] tensor and row (b,i,j) of the result is stored
 *   at (b,j,i) (y and residual are indexed by the transposed row; gate / row_scale by the GEMM row) — the
 *   'b i j c -> b j i c' rearrange after the per-column triangle attention (seqformer.py:547-548)
 *   requirements: K, ldx, ldw multiples of 4; x, w 16-byte aligned
 *   tile_n: output tile width 32/64/128, 0 = chosen from the problem shape */
int abx_gemm_tf32x3(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                    const float* bias, const float* residual, const float* gate, const float* row_scale, int act,
                    int transpose_n, float* y, int ldy, int tile_n);
/* Same with the low part of a STATIC weight supplied by the caller: w_lo[i] = w[i] - tf32_trunc(w[i]) (tf32_trunc clears the
 * 13 low mantissa bits), same layout as w; NULL = split inside the kernel.  The kernel then fetches both parts by TMA and its
 * converter warps only handle the activations (shorter load -> convert -> MMA latency per k-slab). */
int abx_gemm_tf32x3_wlo(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w, const float* w_lo, int ldw,
                        const float* bias, const float* residual, const float* gate, const float* row_scale, int act,
                        int transpose_n, float* y, int ldy, int tile_n);

/* Triangle multiplication (seqformer.py:413-504) on the tensor cores, three calls:
 * (1) abx_gemm_tf32x3_glu_cm: the left/right projections and gates of LN(pair) as ONE GEMM with the GLU epilogue
 *     (act 5), mask row scale, and the result stored channel-major y [B, Nout/2, n, np] (rows padded to np floats,
 *     pad columns must be zero: the caller zero-fills the buffer once) — the K-major operands of the product;
 * (2) abx_gemm_tf32x3_batched_nt: out[bc][i][j] = sum_k a[bc][i][k] b[bc][j][k] for the B*C channel problems
 *     (row base of problem bc = (bc / inner) * outer_rows + (bc % inner) * n in the [total_rows, kpad] matrices);
 * (3) abx_layernorm_cm: 'b c i j -> b i j c' + final LayerNorm, then the gated output projection is a GEMM.
 * The incoming orientation uses the same calls on the transposed LN output (abx_layernorm transpose_n). */
int abx_gemm_tf32x3_glu_cm(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                           const float* bias, const float* row_scale, int n, int np, float* y);
int abx_gemm_tf32x3_glu_cm_wlo(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w, const float* w_lo,
                               int ldw, const float* bias, const float* row_scale, int n, int np, float* y);   /* w_lo: see above */
int abx_gemm_tf32x3_batched_nt(void* stream, int batches, int n, int kpad, int inner, int outer_rows, int total_rows,
                               const float* a, const float* b, float* out, int ldo);
int abx_layernorm_cm(void* stream, int B, int C, int n, int np, const float* x, const float* gamma, const float* beta,
                     float eps, float* y);

/* ---- LayerNorm ------------------------------------------------------------------------------- */
/* y = (x - mean) / sqrt(var + eps) * gamma + beta over the last dimension (torch.nn.LayerNorm; every
 * LayerNorm of abx/model/seqformer.py, score_network.py:117-135).  x, y [rows, C] contiguous, C % 4 == 0,
 * C <= 1024.  transpose_n = n > 0: row (b,i,j) of a [B,n,n,C] input is written to (b,j,i) (the rearrange in
 * front of the per-column triangle attention, seqformer.py:537-538); not in place. */
int abx_layernorm(void* stream, long long rows, int C, const float* x, const float* gamma, const float* beta,
                  float eps, int transpose_n, float* y);

/* ---- pair-activation helpers of the trunk ----------------------------------------------------- */
/* Pair input of EmbeddingAndSeqformer.forward (seqformer.py:193-222) in one pass:
 *   y[b,i,j,:] = concat(stat[i,j,0:Cs], te[b,0:Ct], te[b,0:Ct]) + LayerNorm(prev_pair[b,i,j,:]) + emb[prev_pos[b,i,j],:]
 * stat [N,N,Cs] (step-invariant pair embedding of the complex), te [B,Ct] timestep embedding, prev_pair [B,N,N,C]
 * (NULL: term skipped) with LayerNorm gamma/beta [C], prev_pos [B,N,N] i64 + emb [bins,C] (NULL: skipped). */
int abx_pair_input(void* stream, int B, int N, int C, int Cs, int Ct, const float* stat, const float* te,
                   const float* prev_pair, const float* gamma, const float* beta, float eps,
                   const int64_t* prev_pos, const float* emb, float* y);
/* OuterProductMean features (seqformer.py:392-407): out[b,i,j,:] = concat(left[b,j,:] * right[b,i,:],
 * left[b,j,:] - right[b,i,:]);  left, right [B,N,C], out [B,N,N,2C]. */
int abx_outer_product(void* stream, int B, int N, int C, const float* left, const float* right, float* out);

/* ---- attention with pair bias ------------------------------------------------------------------ */
/* Attention core of seqformer.py:283-301 for TriangleAttention (:506-550), logits never materialised:
 *   out[b,s,i,h,:] = sum_j softmax_j(q[b,s,i,h,:].k[b,s,j,h,:]/sqrt(D) + bias[b,h,i,j]) v[b,s,j,h,:]
 * with keys whose key_mask[b,j] == 0 set to finfo.min before the softmax.
 *   q, k, v: element (b,s,l,h,d) at ptr[((b*S+s)*L + l)*ld + h*D + d]  (e.g. three slices of one fused
 *   q|k|v projection buffer, ld = 3*H*D);  bias [B,H,L,L];  key_mask [B,L] f32 or NULL;  out [B,S,L,H*D]
 *   D in {16,32,48,64}; 2*L*D*4 bytes of K/V must fit in shared memory (L <= ~520 at D = 48). */
int abx_pair_attention(void* stream, int B, int S, int L, int H, int D, const float* q, const float* k,
                       const float* v, int ld, const float* bias, const float* key_mask, float* out);
/* Same operation with a choice of implementation: impl 0 / 2 = tensor-core kernel (mma.sync m16n8k8 TF32 with the
 * 3xTF32 operand split, FlashAttention-2 dataflow in registers), impl 1 = the SIMT kernel above.
 * gate (impl 0 / 2 only, may be NULL): pre-activation of the output gate laid out like q (row stride ld);
 * out = sigmoid(gate) * attention  (seqformer.py:296-299) — with q|k|v|gate produced by one GEMM. */
int abx_pair_attention_impl(void* stream, int impl, int B, int S, int L, int H, int D, const float* q, const float* k,
                            const float* v, int ld, const float* bias, const float* key_mask, const float* gate,
                            float* out);
/* Same operation on the 5th-generation tensor cores (tcgen05.mma 3xTF32, S / P / O' in tensor memory, two 128-row query
 * tiles per CTA).  The pair bias and the key mask arrive pre-tiled in ONE tensor
 *   bias_tiles[b][h][kt][it][j][i], kt < ceil(L/64), it < ceil(L/32), j < 64, i < 32
 *     = log2(e) * bias[b,h, 32 it + i, 64 kt + j]   for keys inside L with key_mask != 0
 *     = -FLT_MAX                                    for keys with key_mask == 0 (the reference's masked_fill(finfo.min))
 *     = -inf                                        for padding keys 64 kt + j >= L  (rows beyond L: any finite value)
 * (abx_b200.ops.pair_attention builds it with three torch ops).  D in {16,32,48}; q, k, v, gate, out 16-byte aligned.
 * abx_pair_attention_tc5_supported(L, D) tells whether the operand tiles fit (1) or not (0). */
int abx_pair_attention_tc5(void* stream, int B, int S, int L, int H, int D, const float* q, const float* k,
                           const float* v, int ld, const float* bias_tiles, const float* gate, float* out);
int abx_pair_attention_tc5_supported(int L, int D);

/* Which GEMM the IPA pipeline uses for its node layers: 0 auto (tcgen05 when operands qualify),
 * 1 SIMT (abx_linear_f32), 2 tcgen05 only.  Process-wide; meant for A/B measurements and tests. */
int abx_set_gemm_backend(int backend);

/* Diagnostics of abx_gemm_tf32x3 (no reference counterpart): with ABX_GEMM_PROF=1 in the environment CTA 0 of every
 * launch records per-role wait / work cycle counters; this copies the 32 counters of the last launch (host buffer). */
int abx_gemm_profile(unsigned long long* out32);

/* Same for the tcgen05 triangle-attention kernel (library built with -DABX_ATTN_PROFILE=1): 32 counters of CTA (0,0,0). */
int abx_attention_profile(unsigned long long* out32);

/* ---- Invariant Point Attention ---------------------------------------------------------------- */
/* Weights of abx.model.folding.InvariantPointAttention (folding.py:23-45), reference state_dict
 * layout (out_features x in_features, row-major). */
typedef struct {
  const float* w_q_scalar;  const float* b_q_scalar;    /* [H*16, C], [H*16]        proj_q_scalar       */
  const float* w_kv_scalar; const float* b_kv_scalar;   /* [H*32, C], [H*32]        proj_kv_scalar      */
  const float* w_q_point;   const float* b_q_point;     /* [3*H*4, C], [3*H*4]      proj_q_point_local  */
  const float* w_kv_point;  const float* b_kv_point;    /* [3*H*12, C], [3*H*12]    proj_kv_point_local */
  const float* w_pair;      const float* b_pair;        /* [H, Cz], [H]             proj_pair           */
  const float* point_weights;                           /* [H]                      trainable_point_weights */
  const float* w_final;     const float* b_final;       /* [C, H*(16+8*4+Cz)], [C]  final_proj          */
  /* optional: the four projection weights / biases concatenated by rows in the order above
   * ([H*16 + H*32 + 3*H*4 + 3*H*12 = 1152, C] and [1152]) so that they run as one GEMM; NULL = four GEMMs */
  const float* w_proj_cat;  const float* b_proj_cat;
} abx_ipa_weights;

/* Geometry of the supported configuration (config/config_model.json:107-124): H=12 heads, 16 scalar
 * qk/v channels, 4 query/key points, 8 value points, C=256 node channels, Cz=128 pair channels. */
#define ABX_IPA_H 12
#define ABX_IPA_C 256
#define ABX_IPA_CZ 128
#define ABX_IPA_FEAT 2112   /* H*(16 + 8*3 + 8 + Cz) */
#define ABX_IPA_BIAS_ROW 100 /* floats per (key chunk, query row) of the chunked pair bias: 8 keys x 12 heads + 4 pad */

/* bytes of workspace abx_ipa_forward / abx_ipa_attention_features need for a [B,N] problem (upper bound:
 * includes room for the pair-bias tensor used when pair_bias == NULL) */
size_t abx_ipa_workspace_bytes(int B, int N);

/* sqrt(1/3) * (z[b,i,j,:] . w_pair[h,:] + b_pair[h]) — folding.py:101-104 — in the chunked key-major layout the
 * fused attention kernel streams with one bulk copy per (tile of query rows, chunk of 8 keys):
 *   pair_bias[b][j / 8][i][12 (j % 8) + h],  shape [B, ceil(N/8), N, ABX_IPA_BIAS_ROW]  (abx_ipa_pair_bias_floats floats)
 * Depends only on z and the weights, so IpaScore (score_network.py:126-163, 8 weight-shared
 * iterations over the same pair activations) evaluates it once per call. */
size_t abx_ipa_pair_bias_floats(int B, int N);
int abx_ipa_pair_bias(void* stream, int B, int N, const float* z, const float* w_pair, const float* b_pair,
                      float* pair_bias);

/* InvariantPointAttention.forward (folding.py:47-132).
 *   x [B,N,C]; z [B,N,N,Cz]; mask [B,N] f32; rots [B,N,3,3]; trans [B,N,3] (already / position_scale)
 *   pair_bias: the tensor written by abx_ipa_pair_bias, or NULL (then computed into the workspace)
 *   residual: optional [B,N,C] added to the output (score_network.py:128: seq_act += attn)
 *   out [B,N,C]
 * Limits: 1 <= N <= 1536 (ABX_ERR_INVALID beyond); x, z and pair_bias 16-byte aligned.  The kernels of one call are
 * chained with programmatic dependent launch on `stream`. */
int abx_ipa_forward(void* stream, int B, int N, const float* x, const float* z, const float* mask,
                    const float* rots, const float* trans, const abx_ipa_weights* w, const float* pair_bias,
                    const float* residual, float* out, void* workspace, size_t workspace_bytes);

/* Frame update of one IpaScore iteration (score_network.py:137-149), in place, one launch:
 *   upd [B,N,6] = affine_update(seq_act) = (quaternion update, translation update)
 *   delta_quat, curr_quats [B,N,4] <- normalize(q + q (x) (0, upd[:3]))   (quat_affine.quat_precompose_vec)
 *   curr_trans [B,N,3] (Angstrom / position_scale) <- curr_rots upd[3:] + curr_trans   (r3.rigids_mul_vecs)
 *   residues with fixed_mask != 0 are reset to init_quats / init_trans (already / position_scale)
 *   curr_rots [B,N,3,3] <- quat_to_rot(curr_quats) */
int abx_ipa_frame_update(void* stream, int B, int N, const float* upd, const float* init_quats,
                         const float* init_trans, const int32_t* fixed_mask, float* delta_quat, float* curr_quats,
                         float* curr_trans, float* curr_rots);

/* Watchdog of the fused attention kernel: every mbarrier wait in it gives up after ~0.25 s of spinning, records where
 * (out8[1..4] = wait tag, block, thread, parity) and makes all other waits return, so a synchronisation fault ends the
 * kernel with a readable record instead of hanging the device.  Synchronises the device, copies and clears the record;
 * out8[0] != 0 means a wait timed out since the last call (the results of that launch are garbage). */
int abx_ipa_watchdog_read(unsigned long long* out8);

/* Per-role stall profile of the fused attention kernel (development aid; process-wide switch, synchronises the device):
 * enable != 0 makes later launches add, per role r = 0..5 (z producer, key/value producer, MMA issuers, converters, logits,
 * values), out64[8 r + 0] = cycles in the role's main loop, [8 r + 1], [8 r + 2] = cycles in its two waits, [8 r + 3] =
 * contributing warps; out64 (may be NULL) receives and clears what was gathered so far. */
int abx_ipa_profile(int enable, unsigned long long* out64);

/* Stages of abx_ipa_forward, exported so tests and the benchmark can time/verify them separately. */
/* feats [B,N,2112] = concat(o_scalar 192, o_point_local (r n) 288, o_point_norm 96, o_pair 1536) */
int abx_ipa_attention_features(void* stream, int B, int N, const float* x, const float* z, const float* mask,
                               const float* rots, const float* trans, const abx_ipa_weights* w,
                               const float* pair_bias, float* feats, void* workspace, size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ABX_B200_H_ */
