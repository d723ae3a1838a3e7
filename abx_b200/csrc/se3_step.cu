// SE(3) + categorical diffuser kernels: scores of a predicted x0, tau-leap rates, and the fused
// Euler-Maruyama reverse step.  Each kernel cites the reference lines it re-implements; the CPU
// restatement used as the test oracle lives in oracle/diffusers.py.
#include "common.cuh"

namespace abx {

constexpr int kS = 20;   // residue types (residue_constants.restype_num)

// ---------------------------------------------------------------------------------------------------
// rot_score / trans_score of a predicted x0.  One thread per residue; the table lookup reproduces
// torch.bucketize on discrete_omega[:-1] and the sigma index of so3_diffuser.py:189-196.
// ---------------------------------------------------------------------------------------------------
// so3_diffuser.py:282-297 (cached branch): |v| -> bucket of the omega grid -> score norm of row sigma_idx(t)
__device__ __forceinline__ void so3_score_lookup(const Vec3<float>& v, const abx_diffuser_consts& c, double tb, bool t32,
                                                 const float* __restrict__ score_norms, const float* __restrict__ dsigma,
                                                 const float* __restrict__ domega, float* __restrict__ out) {
  float omega = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z) + 1e-6f;
  int row = so3_sigma_idx(dsigma, c.num_sigma, tb, t32, c.so3_exp_max, c.so3_exp_min);
  row = max(0, min(row, c.num_sigma - 1));
  int lo = 0, hi = c.num_omega - 1;          // bucketize(omega, discrete_omega[:-1]), right=False
  while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(domega + mid) < omega) lo = mid + 1; else hi = mid; }
  float val = __ldg(score_norms + (size_t)row * c.num_omega + lo);
  float den = omega + 1e-6f;
  out[0] = val * v.x / den;
  out[1] = val * v.y / den;
  out[2] = val * v.z / den;
}

template <bool kT32>
__global__ void __launch_bounds__(128) so3_score_rotvec_kernel(
    int B, int N, abx_diffuser_consts c, const float* __restrict__ rotvec, const double* __restrict__ t,
    const float* __restrict__ score_norms, const float* __restrict__ dsigma, const float* __restrict__ domega,
    float* __restrict__ rot_score) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  Vec3<float> v = {rotvec[idx * 3], rotvec[idx * 3 + 1], rotvec[idx * 3 + 2]};
  so3_score_lookup(v, c, t[idx / N], kT32, score_norms, dsigma, domega, rot_score + (size_t)idx * 3);
}

template <bool kT32>
__global__ void __launch_bounds__(128) se3_scores_kernel(
    int B, int N, abx_diffuser_consts c, const float* __restrict__ quat_t, const float* __restrict__ quat_0,
    const float* __restrict__ trans_t, const float* __restrict__ trans_0, const double* __restrict__ t,
    const float* __restrict__ score_norms, const float* __restrict__ dsigma, const float* __restrict__ domega,
    float* __restrict__ rot_score, void* __restrict__ trans_score_v) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  int b = idx / N;
  double tb = t[b];

  if (rot_score != nullptr) {
    // full_diffuser.py:135-142: rotvec(q0^-1 (x) q_t), all float32
    const float4 qt4 = reinterpret_cast<const float4*>(quat_t)[idx];
    const float4 q04 = reinterpret_cast<const float4*>(quat_0)[idx];
    Quat<float> qt = {qt4.x, qt4.y, qt4.z, qt4.w};
    Quat<float> q0 = {q04.x, q04.y, q04.z, q04.w};
    Vec3<float> v = quat_to_rotvec(quat_mul(quat_inv(q0), qt));
    so3_score_lookup(v, c, tb, kT32, score_norms, dsigma, domega, rot_score + (size_t)idx * 3);
  }
  if (trans_score_v != nullptr) {
    // r3_diffuser.py:158-164 with scale=True; marginal_b_t :45-46
    const float cs = (float)c.r3_coord_scale;
    float xt[3], x0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { xt[k] = trans_t[idx * 3 + k] * cs; x0[k] = trans_0[idx * 3 + k] * cs; }
    if (kT32) {
      float tf = (float)tb;
      float beta = tf * (float)c.r3_min_b + 0.5f * (tf * tf) * (float)c.r3_delta_b;
      float e_half = expf(-0.5f * beta), den = 1.0f - expf(-beta);
      float* out = reinterpret_cast<float*>(trans_score_v);
#pragma unroll
      for (int k = 0; k < 3; ++k) out[idx * 3 + k] = -(xt[k] - e_half * x0[k]) / den;
    } else {
      double beta = tb * c.r3_min_b + 0.5 * (tb * tb) * c.r3_delta_b;
      double e_half = exp(-0.5 * beta), den = 1.0 - exp(-beta);
      double* out = reinterpret_cast<double*>(trans_score_v);
#pragma unroll
      for (int k = 0; k < 3; ++k) out[idx * 3 + k] = -((double)xt[k] - e_half * (double)x0[k]) / den;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Truncated IGSO(3) character series, shared by the live score and the table builder.
//   f(w,s)   = sum_l (2l+1) e^{-l(l+1)s^2/2} sin((l+1/2)w) / sin(w/2)             so3_diffuser.py:15-49
//   df/dw    = sum_l (2l+1) e^{-l(l+1)s^2/2} (lo*dhi - hi*dlo)/lo^2               so3_diffuser.py:72-112
// One warp per (omega, sigma) point: lanes stride over l, then a shuffle reduction.  Accumulated in
// float64; terms whose Gaussian factor underflows to exactly 0 are skipped.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void igso3_series_warp(double omega, double sigma, int L, double* f_out, double* df_out) {
  const int lane = threadIdx.x & 31;
  const double lo = sin(0.5 * omega), dlo = 0.5 * cos(0.5 * omega);
  const double hs2 = 0.5 * sigma * sigma;
  // e^{-l(l+1) hs2} == 0 in float64 once l(l+1) hs2 > 745.2
  int lmax = L;
  if (hs2 > 0.0) {
    double lim = sqrt(745.2 / hs2) + 2.0;
    if (lim < (double)L) lmax = (int)lim;
  }
  double f = 0.0, df = 0.0;
  for (int l = lane; l < lmax; l += 32) {
    double lh = (double)l + 0.5;
    double a = (2.0 * l + 1.0) * exp(-(double)l * (double)(l + 1) * hs2);
    double s, co;
    sincos(omega * lh, &s, &co);
    f += a * s / lo;
    df += a * (lo * (lh * co) - s * dlo) / (lo * lo);
  }
  *f_out = warp_sum(f);
  *df_out = warp_sum(df);
}

template <bool kT32>
__global__ void __launch_bounds__(256) igso3_score_series_kernel(
    int B, int N, abx_diffuser_consts c, const float* __restrict__ rotvec, const double* __restrict__ t,
    const float* __restrict__ dsigma, int L, float* __restrict__ rot_score) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= B * N) return;
  int b = warp / N;
  float vx = rotvec[warp * 3 + 0], vy = rotvec[warp * 3 + 1], vz = rotvec[warp * 3 + 2];
  float omega = sqrtf(vx * vx + vy * vy + vz * vz) + 1e-6f;
  int row = so3_sigma_idx(dsigma, c.num_sigma, t[b], kT32, c.so3_exp_max, c.so3_exp_min);
  row = max(0, min(row, c.num_sigma - 1));
  double f, df;
  igso3_series_warp((double)omega, (double)__ldg(dsigma + row), L, &f, &df);
  if ((threadIdx.x & 31) == 0) {
    float val = (float)(df / (f + 1e-4));
    float den = omega + 1e-6f;
    rot_score[warp * 3 + 0] = val * vx / den;
    rot_score[warp * 3 + 1] = val * vy / den;
    rot_score[warp * 3 + 2] = val * vz / den;
  }
}

// Table build, pass 1: pdf and score norm at every grid point (so3_diffuser.py:150-166).
__global__ void __launch_bounds__(256) igso3_tables_kernel(int num_sigma, int num_omega, int L,
                                                           const float* __restrict__ dsigma,
                                                           const float* __restrict__ domega,
                                                           float* __restrict__ pdf, float* __restrict__ score_norms) {
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)num_sigma * num_omega) return;
  int si = (int)(warp / num_omega), wi = (int)(warp % num_omega);
  double omega = (double)__ldg(domega + wi), sigma = (double)__ldg(dsigma + si);
  double f, df;
  igso3_series_warp(omega, sigma, L, &f, &df);
  if ((threadIdx.x & 31) == 0) {
    pdf[warp] = (float)(f * (1.0 - cos(omega)) / 3.14159265358979323846);   // density(marginal=True) :52-69
    score_norms[warp] = (float)(df / (f + 1e-4));
  }
}

// Table build, pass 2: cdf = cumsum(pdf) / num_omega * pi along omega; one CTA per sigma row.
__global__ void __launch_bounds__(256) igso3_cdf_kernel(int num_omega, const float* __restrict__ pdf,
                                                        float* __restrict__ cdf) {
  __shared__ double warp_tot[8];
  __shared__ double carry_s;
  const int row = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  for (int base = 0; base < num_omega; base += 256) {
    int i = base + threadIdx.x;
    double v = (i < num_omega) ? (double)pdf[(size_t)row * num_omega + i] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    double pre = carry_s;
    for (int k = 0; k < wid; ++k) pre += warp_tot[k];
    v += pre;
    if (i < num_omega) cdf[(size_t)row * num_omega + i] = (float)(v / num_omega * 3.14159265358979323846);
    __syncthreads();
    if (threadIdx.x == 255) carry_s = v;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// Categorical tau-leap rates (discrete_diffuser.py:130-179).  One thread per residue.
// The transition matrix of the uniform-rate CTMC is V diag(e^{lambda t}) V^T with lambda in {0, -S r};
// its closed form e^{-S r t} I + (1 - e^{-S r t})/S 11^T is used instead of the eigendecomposition.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) seq_reverse_rates_kernel(int B, int N, abx_diffuser_consts c,
                                                                const int64_t* __restrict__ seq_t,
                                                                const float* __restrict__ logits,
                                                                const double* __restrict__ t, float dt,
                                                                float* __restrict__ rate_dt) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  int b = idx / N;
  long long xs = seq_t[idx];
  int x = (int)(xs < 0 ? 0 : (xs > kS - 1 ? kS - 1 : xs));          // clamp :148
  const float r = (float)c.seq_rate;
  float tf = (float)t[b];
  float e = expf(-(float)kS * r * tf);
  float off = (1.0f - e) / (float)kS;
  float dg = off + e;
  if (off < 1e-8f) off = 0.0f;                                       // transitions[<1e-8] = 0  :65
  if (dg < 1e-8f) dg = 0.0f;

  float p[kS];
  const float4* lp = reinterpret_cast<const float4*>(logits + (size_t)idx * kS);
#pragma unroll
  for (int k = 0; k < kS / 4; ++k) { float4 v = lp[k]; p[4 * k] = v.x; p[4 * k + 1] = v.y; p[4 * k + 2] = v.z; p[4 * k + 3] = v.w; }
  float mx = p[0];
#pragma unroll
  for (int s = 1; s < kS; ++s) mx = fmaxf(mx, p[s]);
  float sum = 0.f;
#pragma unroll
  for (int s = 0; s < kS; ++s) { p[s] = expf(p[s] - mx); sum += p[s]; }
  float wsum = 0.f;
#pragma unroll
  for (int s = 0; s < kS; ++s) {
    float den = ((s == x) ? dg : off) + 1e-9f;                       // qt0[s, x] + eps_ratio  :166-170
    p[s] = (p[s] / sum) / den;
    wsum += p[s];
  }
  float4* out = reinterpret_cast<float4*>(rate_dt + (size_t)idx * kS);
  float o[kS];
#pragma unroll
  for (int s = 0; s < kS; ++s) {
    float inner = off * (wsum - p[s]) + dg * p[s];                   // (p0/den) @ qt0
    o[s] = (s == x) ? 0.0f : (r * inner) * dt;                       // forward rate, diagonal zeroed :172-179
  }
#pragma unroll
  for (int k = 0; k < kS / 4; ++k) out[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

// ---------------------------------------------------------------------------------------------------
// Fused reverse step (full_diffuser.py:174-227).  One CTA per batch element, threads stride over the
// residues; the only cross-residue dependency is the centre of mass of the translated coordinates
// (r3_diffuser.py:141-146), reduced with warp shuffles + one shared-memory pass.  float64 throughout,
// as the reference computes once t is float64 (inference.py:216).
// ---------------------------------------------------------------------------------------------------
constexpr int kStepThreads = 256;

__global__ void __launch_bounds__(kStepThreads) se3_reverse_step_kernel(
    int N, abx_diffuser_consts c, const void* __restrict__ rigid_t_v, int rigid_is_f64,
    const int64_t* __restrict__ seq_t, const float* __restrict__ rot_score, const double* __restrict__ trans_score,
    const int32_t* __restrict__ diffuse_mask, const double* __restrict__ t, double dt, double sqrt_dt,
    float noise_scale, const float* __restrict__ z_rot, const float* __restrict__ z_trans,
    const float* __restrict__ jumps, int flags, double* __restrict__ rigids_out, int64_t* __restrict__ seq_out) {
  const int b = blockIdx.x;
  const bool do_rot = flags & 1, do_trans = flags & 2, do_seq = flags & 4, center = flags & 8;
  const double tb = t[b];
  const double sigma = so3_sigma<double>(tb, c.so3_exp_max, c.so3_exp_min);
  const double g_rot = sqrt(c.so3_g2_coef * sigma / exp(sigma));           // so3_diffuser.py:207-216
  const double b_t = c.r3_min_b + tb * c.r3_delta_b;                       // r3_diffuser.py:29-32
  const double g_tr = sqrt(b_t);
  const double cs = c.r3_coord_scale;
  const float* rig32 = reinterpret_cast<const float*>(rigid_t_v);
  const double* rig64 = reinterpret_cast<const double*>(rigid_t_v);

  double com[3] = {0.0, 0.0, 0.0};
  for (int n = threadIdx.x; n < N; n += kStepThreads) {
    const size_t i = (size_t)b * N + n;
    double r7[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) r7[k] = rigid_is_f64 ? rig64[i * 7 + k] : (double)rig32[i * 7 + k];
    const bool m = diffuse_mask ? (diffuse_mask[i] != 0) : true;

    // ---- rotation: geodesic random walk (so3_diffuser.py:351-361) ----
    Quat<double> qt = {r7[0], r7[1], r7[2], r7[3]};
    Vec3<double> rot_t = quat_to_rotvec(qt);                               // _extract_trans_rots :12-18
    Vec3<double> rot_o = rot_t;
    if (do_rot && m) {
      const double g2 = g_rot * g_rot, gs = g_rot * sqrt_dt;
      Vec3<double> p;
      p.x = g2 * (double)rot_score[i * 3 + 0] * dt + gs * (double)(noise_scale * z_rot[i * 3 + 0]);
      p.y = g2 * (double)rot_score[i * 3 + 1] * dt + gs * (double)(noise_scale * z_rot[i * 3 + 1]);
      p.z = g2 * (double)rot_score[i * 3 + 2] * dt + gs * (double)(noise_scale * z_rot[i * 3 + 2]);
      rot_o = quat_to_rotvec(quat_mul(rotvec_to_quat(rot_t), rotvec_to_quat(p)));
    }
    Quat<double> qo = rotvec_to_quat(rot_o);                               // _assemble_rigid :20-26
    rigids_out[i * 7 + 0] = qo.w; rigids_out[i * 7 + 1] = qo.x;
    rigids_out[i * 7 + 2] = qo.y; rigids_out[i * 7 + 3] = qo.z;

    // ---- translation: VP-SDE step on scaled coordinates (r3_diffuser.py:133-140) ----
    if (do_trans) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double x = rigid_is_f64 ? r7[4 + k] * cs : (double)((float)r7[4 + k] * (float)cs);
        double f = (-0.5 * b_t) * x;
        double pert = (f - (g_tr * g_tr) * trans_score[i * 3 + k]) * dt
                    + (g_tr * dt) * (double)(noise_scale * z_trans[i * 3 + k]);
        double x1 = x - pert;
        com[k] += x1;
        rigids_out[i * 7 + 4 + k] = x1;          // un-centred, fixed up below
      }
    }

    // ---- residue type: apply the tau-leap jumps (discrete_diffuser.py:181-188) ----
    long long s_in = seq_t[i];
    long long s_out = s_in;
    if (do_seq && m) {
      int x = (int)(s_in < 0 ? 0 : (s_in > kS - 1 ? kS - 1 : s_in));
      float acc = 0.f;
#pragma unroll
      for (int s = 0; s < kS; ++s) acc += jumps[i * kS + s] * (float)(s - x);
      float xp = fminf(fmaxf((float)x + acc, 0.f), (float)(kS - 1));
      s_out = (long long)(int)xp;
    }
    seq_out[i] = s_out;
  }

  if (!do_trans) {
    for (int n = threadIdx.x; n < N; n += kStepThreads) {
      const size_t i = (size_t)b * N + n;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        rigids_out[i * 7 + 4 + k] = rigid_is_f64 ? rig64[i * 7 + 4 + k] : (double)rig32[i * 7 + 4 + k];
    }
    return;
  }

  // centre of mass over ALL residues (mask=None -> ones, r3_diffuser.py:141-146)
  __shared__ double red[3][kStepThreads / 32];
  __shared__ double com_s[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = warp_sum(com[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < kStepThreads / 32; ++w) v += red[threadIdx.x][w];
    com_s[threadIdx.x] = center ? v / (double)(float)N : 0.0;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += kStepThreads) {       // same thread wrote these entries above
    const size_t i = (size_t)b * N + n;
    const bool m = diffuse_mask ? (diffuse_mask[i] != 0) : true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double x1 = (rigids_out[i * 7 + 4 + k] - com_s[k]) / cs;
      double x_in = rigid_is_f64 ? rig64[i * 7 + 4 + k] : (double)rig32[i * 7 + 4 + k];
      rigids_out[i * 7 + 4 + k] = m ? x1 : x_in;              // _apply_mask, full_diffuser.py:219-221
    }
  }
}

}  // namespace abx

using namespace abx;

extern "C" int abx_se3_scores(void* stream, int B, int N, const abx_diffuser_consts* c, const float* quat_t,
                              const float* quat_0, const float* trans_t, const float* trans_0, const double* t,
                              int t_is_f32, const float* score_norms, const float* discrete_sigma,
                              const float* discrete_omega, float* rot_score, void* trans_score) {
  ABX_REQUIRE(B > 0 && N > 0 && c && t, "abx_se3_scores: bad shape or null argument");
  if (rot_score) ABX_REQUIRE(quat_t && quat_0 && score_norms && discrete_sigma && discrete_omega,
                             "abx_se3_scores: rot_score requested without quaternions/tables");
  if (trans_score) ABX_REQUIRE(trans_t && trans_0, "abx_se3_scores: trans_score requested without translations");
  dim3 grid(ceil_div(B * N, 128)), block(128);
  cudaStream_t s = (cudaStream_t)stream;
  if (t_is_f32)
    se3_scores_kernel<true><<<grid, block, 0, s>>>(B, N, *c, quat_t, quat_0, trans_t, trans_0, t, score_norms,
                                                   discrete_sigma, discrete_omega, rot_score, trans_score);
  else
    se3_scores_kernel<false><<<grid, block, 0, s>>>(B, N, *c, quat_t, quat_0, trans_t, trans_0, t, score_norms,
                                                    discrete_sigma, discrete_omega, rot_score, trans_score);
  count_launch();
  return check_launch("se3_scores_kernel");
}

extern "C" int abx_so3_score_rotvec(void* stream, int B, int N, const abx_diffuser_consts* c, const float* rotvec,
                                    const double* t, int t_is_f32, const float* score_norms,
                                    const float* discrete_sigma, const float* discrete_omega, float* rot_score) {
  ABX_REQUIRE(B > 0 && N > 0 && c && rotvec && t && score_norms && discrete_sigma && discrete_omega && rot_score,
              "abx_so3_score_rotvec: bad shape or null argument");
  dim3 grid(ceil_div(B * N, 128)), block(128);
  cudaStream_t s = (cudaStream_t)stream;
  if (t_is_f32)
    so3_score_rotvec_kernel<true><<<grid, block, 0, s>>>(B, N, *c, rotvec, t, score_norms, discrete_sigma, discrete_omega, rot_score);
  else
    so3_score_rotvec_kernel<false><<<grid, block, 0, s>>>(B, N, *c, rotvec, t, score_norms, discrete_sigma, discrete_omega, rot_score);
  count_launch();
  return check_launch("so3_score_rotvec_kernel");
}

extern "C" int abx_igso3_score_series(void* stream, int B, int N, const abx_diffuser_consts* c, const float* rotvec,
                                      const double* t, int t_is_f32, const float* discrete_sigma, int L,
                                      float* rot_score) {
  ABX_REQUIRE(B > 0 && N > 0 && c && rotvec && t && discrete_sigma && rot_score && L > 0,
              "abx_igso3_score_series: bad shape or null argument");
  dim3 grid(ceil_div(B * N * 32, 256)), block(256);
  cudaStream_t s = (cudaStream_t)stream;
  if (t_is_f32)
    igso3_score_series_kernel<true><<<grid, block, 0, s>>>(B, N, *c, rotvec, t, discrete_sigma, L, rot_score);
  else
    igso3_score_series_kernel<false><<<grid, block, 0, s>>>(B, N, *c, rotvec, t, discrete_sigma, L, rot_score);
  count_launch();
  return check_launch("igso3_score_series_kernel");
}

extern "C" int abx_seq_reverse_rates(void* stream, int B, int N, const abx_diffuser_consts* c, const int64_t* seq_t,
                                     const float* logits, const double* t, double dt, float* rate_dt) {
  ABX_REQUIRE(B > 0 && N > 0 && c && seq_t && logits && t && rate_dt, "abx_seq_reverse_rates: bad shape or null argument");
  seq_reverse_rates_kernel<<<ceil_div(B * N, 128), 128, 0, (cudaStream_t)stream>>>(B, N, *c, seq_t, logits, t,
                                                                                  (float)dt, rate_dt);
  count_launch();
  return check_launch("seq_reverse_rates_kernel");
}

extern "C" int abx_se3_reverse_step(void* stream, int B, int N, const abx_diffuser_consts* c, const void* rigid_t,
                                    int rigid_is_f64, const int64_t* seq_t, const float* rot_score,
                                    const double* trans_score, const int32_t* diffuse_mask, const double* t,
                                    double dt_f32, double sqrt_dt_f32, double noise_scale, const float* z_rot,
                                    const float* z_trans, const float* jumps, int flags, double* rigids_out,
                                    int64_t* seq_out) {
  ABX_REQUIRE(B > 0 && N > 0 && c && rigid_t && seq_t && t && rigids_out && seq_out,
              "abx_se3_reverse_step: bad shape or null argument");
  if (flags & 1) ABX_REQUIRE(rot_score && z_rot, "abx_se3_reverse_step: diffuse_rot needs rot_score and z_rot");
  if (flags & 2) ABX_REQUIRE(trans_score && z_trans, "abx_se3_reverse_step: diffuse_trans needs trans_score and z_trans");
  if (flags & 4) ABX_REQUIRE(jumps, "abx_se3_reverse_step: diffuse_seq needs the Poisson jumps");
  se3_reverse_step_kernel<<<B, kStepThreads, 0, (cudaStream_t)stream>>>(
      N, *c, rigid_t, rigid_is_f64, seq_t, rot_score, trans_score, diffuse_mask, t, dt_f32, sqrt_dt_f32,
      (float)noise_scale, z_rot, z_trans, jumps, flags, rigids_out, seq_out);
  count_launch();
  return check_launch("se3_reverse_step_kernel");
}

extern "C" int abx_igso3_build_tables(void* stream, int num_sigma, int num_omega, int L, const float* discrete_sigma,
                                      const float* discrete_omega, float* pdf, float* cdf, float* score_norms) {
  ABX_REQUIRE(num_sigma > 0 && num_omega > 0 && L > 0 && discrete_sigma && discrete_omega && pdf && cdf && score_norms,
              "abx_igso3_build_tables: bad shape or null argument");
  long long warps = (long long)num_sigma * num_omega;
  long long blocks = (warps * 32 + 255) / 256;
  ABX_REQUIRE(blocks < 2147483647LL, "abx_igso3_build_tables: grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  igso3_tables_kernel<<<(unsigned)blocks, 256, 0, s>>>(num_sigma, num_omega, L, discrete_sigma, discrete_omega, pdf,
                                                      score_norms);
  count_launch();
  int rc = check_launch("igso3_tables_kernel");
  if (rc) return rc;
  igso3_cdf_kernel<<<num_sigma, 256, 0, s>>>(num_omega, pdf, cdf);
  count_launch();
  return check_launch("igso3_cdf_kernel");
}
