// Shared host/device helpers for the abx_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "abx_b200.h"

namespace abx {

// ---- error plumbing (C ABI never throws) ---------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);   // cudaGetLastError after a launch -> ABX_OK / ABX_ERR_CUDA

#define ABX_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      ::abx::set_error(__VA_ARGS__);            \
      return ABX_ERR_INVALID;                   \
    }                                           \
  } while (0)

#define ABX_CUDA(call)                                                            \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::abx::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
      return ABX_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Logistic function without slow paths: 1 / (1 + 2^(-x log2 e)) on the bare MUFU.EX2 / MUFU.RCP (4 instructions; expf + an IEEE
// division is ~25 with two special-case branches, which made the gated GEMM epilogues compute-bound).  |error| < 4e-7.
__device__ __forceinline__ float sigmoid_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// A kernel launched with the programmatic-stream-serialization attribute may start while its predecessor in
// the stream is still draining; it must execute griddep_wait() before it touches anything the predecessor
// wrote (the wait returns once the predecessor grid has completed and its writes are visible).  Both
// instructions are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Scope flag set by a host-side pipeline (abx_ipa_forward) around a chain of short dependent kernels: launches
// made through launch_kernel() inside the scope carry the PDL attribute.  Thread-local, so it never leaks into
// another host thread's launches.
bool pdl_scope_active();
void pdl_scope_set(bool on);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_scope_active() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- quaternion / rotation-vector algebra ----------------------------------------------------------
// Real-first unit quaternions [w,x,y,z].  Templated on float/double: the model side of the reference
// runs these in float32 (abx/model/quat_affine.py), the diffuser side in float64 once t is float64.
template <typename T>
struct Quat {
  T w, x, y, z;
};
template <typename T>
struct Vec3 {
  T x, y, z;
};

template <typename T> __device__ __forceinline__ T t_sqrt(T v);
template <> __device__ __forceinline__ float t_sqrt<float>(float v) { return sqrtf(v); }
template <> __device__ __forceinline__ double t_sqrt<double>(double v) { return sqrt(v); }
template <typename T> __device__ __forceinline__ T t_atan2(T a, T b);
template <> __device__ __forceinline__ float t_atan2<float>(float a, float b) { return atan2f(a, b); }
template <> __device__ __forceinline__ double t_atan2<double>(double a, double b) { return atan2(a, b); }
template <typename T> __device__ __forceinline__ T t_sin(T v);
template <> __device__ __forceinline__ float t_sin<float>(float v) { return sinf(v); }
template <> __device__ __forceinline__ double t_sin<double>(double v) { return sin(v); }
template <typename T> __device__ __forceinline__ T t_cos(T v);
template <> __device__ __forceinline__ float t_cos<float>(float v) { return cosf(v); }
template <> __device__ __forceinline__ double t_cos<double>(double v) { return cos(v); }
template <typename T> __device__ __forceinline__ T t_abs(T v) { return v < T(0) ? -v : v; }

// quat_affine.py:76-82 (table :27-46): Hamilton product a (x) b
template <typename T>
__device__ __forceinline__ Quat<T> quat_mul(const Quat<T>& a, const Quat<T>& b) {
  Quat<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}

// quat_affine.py:234-238: conjugate / |q|
template <typename T>
__device__ __forceinline__ Quat<T> quat_inv(const Quat<T>& q) {
  T n = t_sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  Quat<T> r = {q.w / n, -q.x / n, -q.y / n, -q.z / n};
  return r;
}

// sin(angle/2)/angle with the 1/2 - angle^2/48 series below 1e-6 (quat_affine.py:120-130,136-146)
template <typename T>
__device__ __forceinline__ T half_sinc(T half, T angle) {
  return (t_abs(angle) < T(1e-6)) ? (T(0.5) - angle * angle / T(48)) : (t_sin(half) / angle);
}

// quat_affine.py:113-131: flip to w >= 0, angle = 2 atan2(|xyz|, w), v = xyz / (sin(angle/2)/angle)
template <typename T>
__device__ __forceinline__ Vec3<T> quat_to_rotvec(Quat<T> q) {
  if (q.w < T(0)) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  T n = t_sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  T half = t_atan2(n, q.w);
  T angle = T(2) * half;
  T k = half_sinc(half, angle);
  Vec3<T> v = {q.x / k, q.y / k, q.z / k};
  return v;
}

// quat_affine.py:133-150
template <typename T>
__device__ __forceinline__ Quat<T> rotvec_to_quat(const Vec3<T>& v) {
  T angle = t_sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  T half = angle * T(0.5);
  T k = half_sinc(half, angle);
  Quat<T> q = {t_cos(half), v.x * k, v.y * k, v.z * k};
  return q;
}

// quat_affine.py:60-67 (table :9-25): rotation matrix rows r[0..8]
template <typename T>
__device__ __forceinline__ void quat_to_rot(const Quat<T>& q, T* r) {
  T w = q.w, x = q.x, y = q.y, z = q.z;
  r[0] = w * w + x * x - y * y - z * z; r[1] = T(2) * (x * y - w * z); r[2] = T(2) * (x * z + w * y);
  r[3] = T(2) * (x * y + w * z); r[4] = w * w - x * x + y * y - z * z; r[5] = T(2) * (y * z - w * x);
  r[6] = T(2) * (x * z - w * y); r[7] = T(2) * (y * z + w * x); r[8] = w * w - x * x - y * y + z * z;
}

// so3_diffuser.py:198-205 (logarithmic schedule): sigma(t) = log(t e^{max} + (1-t) e^{min})
template <typename T>
__device__ __forceinline__ T so3_sigma(T t, T e_max, T e_min);
template <>
__device__ __forceinline__ double so3_sigma<double>(double t, double e_max, double e_min) {
  return log(t * e_max + (1.0 - t) * e_min);
}
template <>
__device__ __forceinline__ float so3_sigma<float>(float t, float e_max, float e_min) {
  return logf(t * e_max + (1.0f - t) * e_min);
}

// so3_diffuser.py:189-196: #{k : grid[k] <= sigma + 1e-5} - 1 on the (monotone) sigma grid.
// The comparison runs in the dtype torch promotes to: float64 when t is float64, else float32.
__device__ __forceinline__ int so3_sigma_idx(const float* __restrict__ grid, int n, double t, bool t_is_f32,
                                             double e_max, double e_min) {
  int lo = 0, hi = n;   // first index with grid[k] > key
  if (t_is_f32) {
    float key = so3_sigma<float>((float)t, (float)e_max, (float)e_min) + 1e-5f;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(grid + mid) <= key) lo = mid + 1; else hi = mid; }
  } else {
    double key = so3_sigma<double>(t, e_max, e_min) + 1e-5;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((double)__ldg(grid + mid) <= key) lo = mid + 1; else hi = mid; }
  }
  return lo - 1;
}

// ---- legacy tensor-core MMA (m16n8k8, TF32 operands, fp32 accumulate) with the exact 3xTF32 operand split ----
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace abx
