// Shared device helpers for the sm_100a MambaTTSDecoder kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mamba_tts_b200.h"

namespace mtts {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kNumSMs = 148;  // B200

// ---- MUFU wrappers (approx, flush-to-zero: ~2 ulp, inside the 1e-4 fp32 budget) -----------------
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// softplus with torch's threshold (x > 20 -> x).  log1p is evaluated as a 3-term series when
// exp(x) is small so that tiny deltas keep full relative precision (lg2(1+e) alone loses it).
__device__ __forceinline__ float softplus_f(float x) {
  const float e = ex2f(fminf(x, 20.f) * kLog2e);
  const float big = lg2f(1.f + e) * kLn2;
  const float small = e * fmaf(e, fmaf(e, 0.33333333f, -0.5f), 1.f);
  const float r = (e < 0.02f) ? small : big;
  return (x > 20.f) ? x : r;
}
// d softplus / dx = sigmoid(x) (1 beyond the threshold)
__device__ __forceinline__ float sigmoid_f(float x) { return rcpf(1.f + ex2f(-x * kLog2e)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }

// ---- io dtype traits -------------------------------------------------------------------------------
template <typename T>
struct Io;
template <>
struct Io<float> {
  static constexpr int kVecElems = 4;  // elements per 16-byte vector
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
  __device__ static __forceinline__ void unpack(const uint4& r, float* o) {
    o[0] = __uint_as_float(r.x);
    o[1] = __uint_as_float(r.y);
    o[2] = __uint_as_float(r.z);
    o[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                      __float_as_uint(v[3]));
  }
};
template <>
struct Io<__nv_bfloat16> {
  static constexpr int kVecElems = 8;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ void unpack(const uint4& r, float* o) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[2 * i] = __uint_as_float(w[i] << 16);
      o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ uint4 ldg16(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
// streaming 16-byte load/store: data touched exactly once, keep it out of L1
__device__ __forceinline__ uint4 ldg16_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg16_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// Load kItems consecutive elements row[t .. t+kItems) as floats, zeros at/after `len`.
// kVec: row + t is 16-byte aligned and len % kVecElems == 0 (whole vectors are in or out).
template <typename T, int kItems, bool kVec>
__device__ __forceinline__ void load_items(const T* __restrict__ row, int t, int len, float* out) {
  if constexpr (kVec) {
    constexpr int VE = Io<T>::kVecElems;
    static_assert(kItems % VE == 0, "items per lane must be whole vectors");
#pragma unroll
    for (int v = 0; v < kItems / VE; ++v) {
      if (t + v * VE < len) {
        const uint4 r = ldg16_stream(row + t + v * VE);
        Io<T>::unpack(r, out + v * VE);
      } else {
#pragma unroll
        for (int i = 0; i < VE; ++i) out[v * VE + i] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < kItems; ++i) out[i] = (t + i < len) ? Io<T>::to_f(row[t + i]) : 0.f;
  }
}

template <typename T, int kItems, bool kVec>
__device__ __forceinline__ void store_items(T* __restrict__ row, int t, int len, const float* in) {
  if constexpr (kVec) {
    constexpr int VE = Io<T>::kVecElems;
#pragma unroll
    for (int v = 0; v < kItems / VE; ++v) {
      if (t + v * VE < len) stg16_stream(row + t + v * VE, Io<T>::pack(in + v * VE));
    }
  } else {
#pragma unroll
    for (int i = 0; i < kItems; ++i)
      if (t + i < len) row[t + i] = Io<T>::from_f(in[i]);
  }
}

// 4-element (8 or 16 byte) loads/stores with conversion to/from fp32
template <typename T>
__device__ __forceinline__ void load4(const T* p, float* o);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* o) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* o) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  o[0] = __uint_as_float(r.x << 16);
  o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16);
  o[3] = __uint_as_float(r.y & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float* v);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  const __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 r;
  r.x = *reinterpret_cast<const uint32_t*>(&a);
  r.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int launch_status() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MTTS_OK : -static_cast<int>(e);
}

}  // namespace mtts
