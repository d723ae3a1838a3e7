// Per-iteration frame update of the IpaScore structure module (reference abx/model/score_network.py:137-149):
//   (dq, dt)    = affine_update(seq_act)                       [B,N,6]  (computed by the GEMM)
//   delta_quat  = normalize(delta_quat + delta_quat (x) (0,dq))          quat_affine.quat_precompose_vec :84-92
//   curr_quats  = normalize(curr_quats + curr_quats (x) (0,dq))
//   curr_trans  = curr_rots dt + curr_trans                               r3.rigids_mul_vecs (old rotation)
//   fixed residues keep their input frame                                 :142-147
//   curr_rots   = quat_to_rot(curr_quats)                                 :149
// The reference spends ~45 eager launches per iteration on this; here it is one thread per residue.
#include "common.cuh"

namespace abx {

__device__ __forceinline__ Quat<float> precompose(const Quat<float>& q, float vx, float vy, float vz) {
  // q + q (x) (0, v), then / sqrt(|.|^2 + 1e-12)   (products and sums rounded separately, as eager torch does)
  Quat<float> r;
  r.w = q.w + __fsub_rn(__fsub_rn(__fsub_rn(__fmul_rn(q.w, 0.f), __fmul_rn(q.x, vx)), __fmul_rn(q.y, vy)), __fmul_rn(q.z, vz));
  r.x = q.x + __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.w, vx), __fmul_rn(q.x, 0.f)), __fmul_rn(q.y, vz)), __fmul_rn(q.z, vy));
  r.y = q.y + __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(q.w, vy), __fmul_rn(q.x, vz)), __fmul_rn(q.y, 0.f)), __fmul_rn(q.z, vx));
  r.z = q.z + __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(q.w, vz), __fmul_rn(q.x, vy)), __fmul_rn(q.y, vx)), __fmul_rn(q.z, 0.f));
  const float n = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.w, r.w), __fmul_rn(r.x, r.x)), __fmul_rn(r.y, r.y)),
                                            __fmul_rn(r.z, r.z)), 1e-12f));
  r.w /= n; r.x /= n; r.y /= n; r.z /= n;
  return r;
}

__global__ void __launch_bounds__(128) ipa_frame_update_kernel(
    int BN, const float* __restrict__ upd, const float* __restrict__ init_quats, const float* __restrict__ init_trans,
    const int* __restrict__ fixed_mask, float* __restrict__ delta_quat, float* __restrict__ curr_quats,
    float* __restrict__ curr_trans, float* __restrict__ curr_rots) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BN) return;
  const float* u = upd + (size_t)i * 6;
  const float qx = u[0], qy = u[1], qz = u[2], tx = u[3], ty = u[4], tz = u[5];
  float4 d4 = reinterpret_cast<float4*>(delta_quat)[i], c4 = reinterpret_cast<float4*>(curr_quats)[i];
  Quat<float> dq = precompose({d4.x, d4.y, d4.z, d4.w}, qx, qy, qz);
  Quat<float> cq = precompose({c4.x, c4.y, c4.z, c4.w}, qx, qy, qz);
  const float* R = curr_rots + (size_t)i * 9;
  float* T = curr_trans + (size_t)i * 3;
  float nt[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    nt[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3 * r], tx), __fmul_rn(R[3 * r + 1], ty)), __fmul_rn(R[3 * r + 2], tz)), T[r]);
  if (fixed_mask[i] != 0) {                      // keep * new + (1 - keep) * init with keep in {0,1}
    const float4 q0 = reinterpret_cast<const float4*>(init_quats)[i];
    cq = {q0.x, q0.y, q0.z, q0.w};
#pragma unroll
    for (int r = 0; r < 3; ++r) nt[r] = init_trans[(size_t)i * 3 + r];
  }
  reinterpret_cast<float4*>(delta_quat)[i] = make_float4(dq.w, dq.x, dq.y, dq.z);
  reinterpret_cast<float4*>(curr_quats)[i] = make_float4(cq.w, cq.x, cq.y, cq.z);
#pragma unroll
  for (int r = 0; r < 3; ++r) T[r] = nt[r];
  // quat_to_rot, quat_affine.py:60-67: products first, then sums left to right
  const float w = cq.w, x = cq.x, y = cq.y, z = cq.z;
  const float ww = __fmul_rn(w, w), xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
  float* Ro = curr_rots + (size_t)i * 9;
  Ro[0] = __fsub_rn(__fsub_rn(__fadd_rn(ww, xx), yy), zz);
  Ro[1] = __fmul_rn(2.f, __fsub_rn(__fmul_rn(x, y), __fmul_rn(w, z)));
  Ro[2] = __fmul_rn(2.f, __fadd_rn(__fmul_rn(x, z), __fmul_rn(w, y)));
  Ro[3] = __fmul_rn(2.f, __fadd_rn(__fmul_rn(x, y), __fmul_rn(w, z)));
  Ro[4] = __fsub_rn(__fadd_rn(__fsub_rn(ww, xx), yy), zz);
  Ro[5] = __fmul_rn(2.f, __fsub_rn(__fmul_rn(y, z), __fmul_rn(w, x)));
  Ro[6] = __fmul_rn(2.f, __fsub_rn(__fmul_rn(x, z), __fmul_rn(w, y)));
  Ro[7] = __fmul_rn(2.f, __fadd_rn(__fmul_rn(y, z), __fmul_rn(w, x)));
  Ro[8] = __fadd_rn(__fsub_rn(__fsub_rn(ww, xx), yy), zz);
}

}  // namespace abx

extern "C" int abx_ipa_frame_update(void* stream, int B, int N, const float* upd, const float* init_quats,
                                    const float* init_trans, const int32_t* fixed_mask, float* delta_quat, float* curr_quats,
                                    float* curr_trans, float* curr_rots) {
  using namespace abx;
  ABX_REQUIRE(B > 0 && N > 0 && upd && init_quats && init_trans && fixed_mask && delta_quat && curr_quats && curr_trans && curr_rots,
              "abx_ipa_frame_update: bad shape or null argument");
  ABX_REQUIRE(((uintptr_t)init_quats % 16 == 0) && ((uintptr_t)delta_quat % 16 == 0) && ((uintptr_t)curr_quats % 16 == 0),
              "abx_ipa_frame_update: quaternion arrays must be 16-byte aligned");
  const int BN = B * N;
  ipa_frame_update_kernel<<<ceil_div(BN, 128), 128, 0, (cudaStream_t)stream>>>(BN, upd, init_quats, init_trans, fixed_mask,
                                                                               delta_quat, curr_quats, curr_trans, curr_rots);
  count_launch();
  return check_launch("ipa_frame_update_kernel");
}
