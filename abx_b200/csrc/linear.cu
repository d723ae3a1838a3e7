// Dense node GEMM  y = act(x W^T + b) (+ residual): torch.nn.Linear semantics of the reference's
// `Linear` (abx/model/common_modules.py:11-59), fp32 in / fp32 accumulate.
//
// v1: register-tiled SIMT kernel (128x64x16 CTA tile, 8x4 per thread, register-prefetched k-slabs).
#include "common.cuh"

namespace abx {

constexpr int kBM = 128, kBN = 64, kBK = 16, kGemmThreads = 256;

int launch_linear_simt(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w,
                       const float* bias, const float* residual, int relu, float* y, int ldy);

__global__ void __launch_bounds__(kGemmThreads) linear_f32_kernel(
    int M, int Nout, int K, const float* __restrict__ x, int ldx, const float* __restrict__ w,
    const float* __restrict__ bias, const float* __restrict__ residual, int relu, float* __restrict__ y, int ldy) {
  __shared__ __align__(16) float As[2][kBK][kBM + 4];
  __shared__ __align__(16) float Bs[2][kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int tm = (tid / 16) * 8, tn = (tid % 16) * 4;     // thread's 8x4 corner inside the CTA tile

  // loader mapping: a k-slab of A is 128 rows x 4 float4, of B 64 rows x 4 float4
  const int a_row0 = tid / 4, a_k4 = (tid % 4) * 4;       // rows a_row0 and a_row0 + 64
  const int b_row = tid / 4, b_k4 = (tid % 4) * 4;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb;
  auto load_slab = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = m0 + a_row0 + 64 * h, k = k0 + a_k4;
      ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < M && k < K) ra[h] = *reinterpret_cast<const float4*>(x + (size_t)r * ldx + k);
    }
    int r = n0 + b_row, k = k0 + b_k4;
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < Nout && k < K) rb = *reinterpret_cast<const float4*>(w + (size_t)r * K + k);
  };
  auto store_slab = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = a_row0 + 64 * h;
      As[buf][a_k4 + 0][r] = ra[h].x; As[buf][a_k4 + 1][r] = ra[h].y;
      As[buf][a_k4 + 2][r] = ra[h].z; As[buf][a_k4 + 3][r] = ra[h].w;
    }
    Bs[buf][b_k4 + 0][b_row] = rb.x; Bs[buf][b_k4 + 1][b_row] = rb.y;
    Bs[buf][b_k4 + 2][b_row] = rb.z; Bs[buf][b_k4 + 3][b_row] = rb.w;
  };

  const int nslab = (K + kBK - 1) / kBK;
  load_slab(0);
  store_slab(0);
  __syncthreads();
  for (int s = 0; s < nslab; ++s) {
    const int buf = s & 1;
    if (s + 1 < nslab) load_slab((s + 1) * kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm + 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (s + 1 < nslab) store_slab(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = m0 + tm + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int cidx = n0 + tn + j;
      if (cidx >= Nout) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + cidx);
      if (relu) v = fmaxf(v, 0.f);
      if (residual) v += residual[(size_t)r * ldy + cidx];
      y[(size_t)r * ldy + cidx] = v;
    }
  }
}

bool gemm_tf32x3_supported(int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw);
int launch_gemm_tf32x3(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w, int ldw,
                       const float* bias, const float* residual, const float* gate, const float* row_scale, int act,
                       int transpose_n, float* y, int ldy, int tile_n, int cm_n, int cm_np, const float* w_lo);

static int g_backend = 0;   // 0 auto (tensor cores when the operands qualify), 1 SIMT only, 2 tensor cores only
int gemm_backend() { return g_backend; }

// Dense layer dispatcher used by the IPA pipeline: 3xTF32 tcgen05 GEMM (gemm_tf32x3.cu) when the operand
// layout allows TMA, else the SIMT kernel below.
int launch_linear_f32(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w,
                      const float* bias, const float* residual, int relu, float* y, int ldy) {
  if (g_backend != 1 && M >= 32 && gemm_tf32x3_supported(M, Nout, K, x, ldx, w, K))
    return launch_gemm_tf32x3(s, M, Nout, K, x, ldx, w, K, bias, residual, nullptr, nullptr, relu ? 1 : 0, 0, y, ldy, 0, 0, 0, nullptr);
  if (g_backend == 2) {
    set_error("linear: operands do not qualify for the tcgen05 path (M=%d Nout=%d K=%d ldx=%d)", M, Nout, K, ldx);
    return ABX_ERR_INVALID;
  }
  return launch_linear_simt(s, M, Nout, K, x, ldx, w, bias, residual, relu, y, ldy);
}

int launch_linear_simt(cudaStream_t s, int M, int Nout, int K, const float* x, int ldx, const float* w,
                       const float* bias, const float* residual, int relu, float* y, int ldy) {
  dim3 grid(ceil_div(Nout, kBN), ceil_div(M, kBM));
  linear_f32_kernel<<<grid, kGemmThreads, 0, s>>>(M, Nout, K, x, ldx, w, bias, residual, relu, y, ldy);
  count_launch();
  return check_launch("linear_f32_kernel");
}

}  // namespace abx

extern "C" int abx_linear_f32(void* stream, int M, int Nout, int K, const float* x, int ldx, const float* w,
                              const float* bias, const float* residual, int relu, float* y, int ldy) {
  ABX_REQUIRE(M > 0 && Nout > 0 && K > 0 && x && w && y, "abx_linear_f32: bad shape or null argument");
  ABX_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldx >= K && ldy >= Nout,
              "abx_linear_f32: K and ldx must be multiples of 4, ldx >= K, ldy >= Nout (K=%d ldx=%d ldy=%d)", K, ldx, ldy);
  ABX_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0), "abx_linear_f32: x and w must be 16-byte aligned");
  return abx::launch_linear_simt((cudaStream_t)stream, M, Nout, K, x, ldx, w, bias, residual, relu, y, ldy);
}

extern "C" int abx_set_gemm_backend(int backend) {
  ABX_REQUIRE(backend >= 0 && backend <= 2, "abx_set_gemm_backend: 0 auto, 1 SIMT, 2 tcgen05");
  abx::g_backend = backend;
  return ABX_OK;
}
