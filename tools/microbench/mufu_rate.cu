// MUFU.EX2 / FFMA2 issue-rate microbenchmark (B200): how many ex2.approx per clock per SM are sustained,
// alone and interleaved with packed FFMA2 the way the scan kernels issue them.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&d))
               : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
                 "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return d;
}

template <int kFmaPerEx>
__global__ void kern(float* out, int iters, long long* cycles) {
  float v[8];
  float2 acc[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.5f, 0.25f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = ex2f(v[i]);  // 8 independent chains
#pragma unroll
      for (int f = 0; f < kFmaPerEx; ++f) acc[(i + f) & 3] = ffma2(acc[(i + f) & 3], make_float2(0.999f, 1.001f), make_float2(v[i], v[i]));
      v[i] = v[i] - 1.0f;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kFmaPerEx>
void run(int warps_per_sm) {
  const int sms = 148, iters = 4096;
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * warps_per_sm * 32);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  kern<kFmaPerEx><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  kern<kFmaPerEx><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
  const double ex = (double)iters * 8 * warps_per_sm * 32;
  printf("fma2_per_ex2=%d warps/SM=%2d : %.2f ex2/clk/SM, %.2f ffma2-lanes/clk/SM (cycles %.0f)\n", kFmaPerEx, warps_per_sm,
         ex / avg, ex * kFmaPerEx / avg, avg);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16, 32}) run<0>(w);
  for (int w : {4, 8, 16, 32}) run<1>(w);
  for (int w : {4, 8, 16, 32}) run<2>(w);
  for (int w : {8, 16, 32}) run<4>(w);
  return 0;
}
