// Elementwise pieces of the FFN / projection backward (mamba_decoder.py:39-43,86-88), for sm_100a:
//   bias_gelu_fwd      out = gelu(x + bias)                                    (exact erf GELU)
//   bias_gelu_bwd      dx = dout * gelu'(x + bias);  colsum[n] += sum_m dx[m, n]   (= d bias)
//   colsum             colsum[n] += sum_m x[m, n]                               (= d bias of a Linear)
// HBM-bound streaming kernels: 16-byte vectors, a warp covers 32 consecutive vectors of a row and walks
// rows; column sums stay in registers, warps of a CTA are combined in shared memory, one RED per column
// and CTA.  They replace aten gelu / gelu_backward / sum(0) (the latter runs at ~0.7 TB/s on (32768, 512)).
#include "common.cuh"

namespace mtts {

namespace {

constexpr int kGlueWarps = 4;
constexpr float kInvSqrt2 = 0.70710678118654752f;
constexpr float kInvSqrt2Pi = 0.39894228040143268f;

// Phi(x) = 0.5 (1 + erf(x / sqrt 2)) through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on erf):
//   erf(u) = 1 - (a1 t + ... + a5 t^5) exp(-u^2),  t = 1 / (1 + 0.3275911 u),  u >= 0
// with u = |x| / sqrt 2, so exp(-u^2) = exp(-x^2 / 2) is also the Gaussian of the derivative: one MUFU.EX2
// and one MUFU.RCP per element instead of libdevice erff's ~30 instructions, which made the exact-GELU
// kernels compute-bound (3.0 TB/s) rather than HBM-bound.
__device__ __forceinline__ void phi_and_gauss(float x, float& phi, float& gauss) {
  const float u = fabsf(x) * kInvSqrt2;
  const float t = rcpf(fmaf(0.3275911f, u, 1.f));
  gauss = ex2f(-0.5f * kLog2e * x * x);  // exp(-x^2/2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float tail = 0.5f * poly * t * gauss;  // 0.5 (1 - erf(u))
  phi = x >= 0.f ? 1.f - tail : tail;
}
__device__ __forceinline__ float gelu_f(float x) {
  float phi, g;
  phi_and_gauss(x, phi, g);
  return x * phi;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float phi, g;
  phi_and_gauss(x, phi, g);
  return fmaf(x * kInvSqrt2Pi, g, phi);
}

template <typename T>
__global__ void __launch_bounds__(256)
bias_gelu_fwd_kernel(const mtts_bias_gelu_params p) {
  constexpr int VE = Io<T>::kVecElems;
  const int vec_per_row = p.cols / VE;
  const int64_t total = (int64_t)p.rows * vec_per_row;
  const T* x = reinterpret_cast<const T*>(p.x);
  T* out = reinterpret_cast<T*>(p.out);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row) * VE;
    float v[VE], bv[VE];
    Io<T>::unpack(ldg16_stream(x + r * p.ld + c), v);
#pragma unroll
    for (int j = 0; j < VE; j += 4) {
      if (p.bias) {
        load4<float>(p.bias + c + j, bv + j);
      } else {
        bv[j] = bv[j + 1] = bv[j + 2] = bv[j + 3] = 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < VE; ++j) v[j] = gelu_f(v[j] + bv[j]);
    stg16_stream(out + r * p.ld + c, Io<T>::pack(v));
  }
}

// kGelu: dx = dout * gelu'(x + bias) is written and summed; else x itself is summed (no output).
template <typename T, bool kGelu>
__global__ void __launch_bounds__(kGlueWarps * 32)
colsum_kernel(const mtts_bias_gelu_params p, const int rows_per_warp) {
  constexpr int VE = Io<T>::kVecElems;
  __shared__ float red[kGlueWarps][32 * VE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (blockIdx.x * 32 + lane) * VE;  // this lane's VE columns
  const bool cok = c < p.cols;
  const int r0 = (blockIdx.y * kGlueWarps + warp) * rows_per_warp;
  const int r1 = min(p.rows, r0 + rows_per_warp);
  float acc[VE], bias[VE];
#pragma unroll
  for (int j = 0; j < VE; ++j) {
    acc[j] = 0.f;
    bias[j] = (kGelu && cok && p.bias) ? p.bias[c + j] : 0.f;
  }
  if (cok) {
    const T* x = reinterpret_cast<const T*>(p.x);
    const T* go = reinterpret_cast<const T*>(p.dout);
    T* dx = reinterpret_cast<T*>(p.out);
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      float v[VE];
      Io<T>::unpack(ldg16_stream(x + (int64_t)r * p.ld + c), v);
      if constexpr (kGelu) {
        float g[VE];
        Io<T>::unpack(ldg16_stream(go + (int64_t)r * p.ld + c), g);
#pragma unroll
        for (int j = 0; j < VE; ++j) v[j] = g[j] * gelu_grad_f(v[j] + bias[j]);
        const uint4 packed = Io<T>::pack(v);
        stg16_stream(dx + (int64_t)r * p.ld + c, packed);
        Io<T>::unpack(packed, v);  // sum what was stored (the rounded values the weight-grad GEMM sees)
      }
#pragma unroll
      for (int j = 0; j < VE; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < VE; ++j) red[warp][lane * VE + j] = acc[j];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VE; i += kGlueWarps * 32) {
    const int col = blockIdx.x * 32 * VE + i;
    if (col < p.cols) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kGlueWarps; ++w) t += red[w][i];
      atomicAdd(p.colsum + col, t);
    }
  }
}

template <typename T>
int check_vec(const mtts_bias_gelu_params& p, const void* a, const void* b, const void* c) {
  constexpr int VE = Io<T>::kVecElems;
  if (p.cols % VE || p.ld % VE || !aligned16(a) || (b && !aligned16(b)) || (c && !aligned16(c))) return MTTS_ERR_ALIGN;
  return MTTS_OK;
}

template <typename T>
int launch_fwd(const mtts_bias_gelu_params& p, cudaStream_t s) {
  if (int e = check_vec<T>(p, p.x, p.out, nullptr)) return e;
  const int64_t total = (int64_t)p.rows * (p.cols / Io<T>::kVecElems);
  const int64_t want = (total + 255) / 256, cap = (int64_t)kNumSMs * 16;
  const int grid = (int)(want < cap ? want : cap);
  bias_gelu_fwd_kernel<T><<<grid, 256, 0, s>>>(p);
  return launch_status();
}

template <typename T, bool kGelu>
int launch_colsum(const mtts_bias_gelu_params& p, cudaStream_t s) {
  if (int e = check_vec<T>(p, p.x, kGelu ? p.dout : nullptr, kGelu ? p.out : nullptr)) return e;
  constexpr int VE = Io<T>::kVecElems;
  const int gx = (p.cols + 32 * VE - 1) / (32 * VE);
  // ~8 waves of CTAs; every warp walks rows_per_warp consecutive rows
  int rows_per_warp = (int)(((int64_t)p.rows * gx + (int64_t)8 * kNumSMs * 8 * kGlueWarps - 1) /
                            ((int64_t)8 * kNumSMs * 8 * kGlueWarps));
  rows_per_warp = max(4, min(rows_per_warp, 256));
  const int gy = (p.rows + rows_per_warp * kGlueWarps - 1) / (rows_per_warp * kGlueWarps);
  if (gy > 65535) return MTTS_ERR_SHAPE;
  colsum_kernel<T, kGelu><<<dim3(gx, gy), kGlueWarps * 32, 0, s>>>(p, rows_per_warp);
  return launch_status();
}

}  // namespace
}  // namespace mtts

extern "C" int mtts_bias_gelu_fwd(const mtts_bias_gelu_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->out) return MTTS_ERR_NULL;
  if (p->rows < 0 || p->cols < 1 || p->ld < p->cols) return MTTS_ERR_SHAPE;
  if (p->rows == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::launch_fwd<float>(*p, s);
    case MTTS_BF16: return mtts::launch_fwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_bias_gelu_bwd(const mtts_bias_gelu_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->dout || !p->out || !p->colsum) return MTTS_ERR_NULL;
  if (p->rows < 0 || p->cols < 1 || p->ld < p->cols) return MTTS_ERR_SHAPE;
  if (p->rows == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::launch_colsum<float, true>(*p, s);
    case MTTS_BF16: return mtts::launch_colsum<__nv_bfloat16, true>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_colsum(const mtts_bias_gelu_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->colsum) return MTTS_ERR_NULL;
  if (p->rows < 0 || p->cols < 1 || p->ld < p->cols) return MTTS_ERR_SHAPE;
  if (p->rows == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::launch_colsum<float, false>(*p, s);
    case MTTS_BF16: return mtts::launch_colsum<__nv_bfloat16, false>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}
