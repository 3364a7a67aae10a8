// Depthwise causal conv1d (+bias, +SiLU) for sm_100a: forward, backward, single-token update.
// Replaces causal_conv1d_cuda.{causal_conv1d_fwd, causal_conv1d_bwd, causal_conv1d_update}
// (Dao-AILab/causal-conv1d), reached from Mamba.forward / Mamba.step <- mamba_decoder.py:61,63.
//
// HBM-bound streaming kernels: every thread owns 8 consecutive timesteps of one (batch, channel)
// row (one or two 16-byte vectors), the width-1 halo comes from the neighbouring lane by shuffle
// (lane 0 / lane 31 fetch it themselves), nothing is staged twice.
#include "common.cuh"

namespace mtts {

constexpr int kConvElems = 8;      // timesteps per thread
constexpr int kConvThreads = 128;  // -> 1024 timesteps per CTA
constexpr int kConvTile = kConvElems * kConvThreads;

template <typename T>
__device__ __forceinline__ float conv_hist(const T* __restrict__ xrow, const T* __restrict__ irow,
                                           int s, int width) {
  // x at absolute timestep s (< current thread's first), s may be negative
  if (s >= 0) return Io<T>::to_f(xrow[s]);
  const int j = (width - 1) + s;  // index into initial_states (width-1 entries)
  return (irow != nullptr && j >= 0) ? Io<T>::to_f(irow[j]) : 0.f;
}

template <typename T, int W, bool kVec>
__global__ void __launch_bounds__(kConvThreads)
conv1d_fwd_kernel(const mtts_conv1d_fwd_params p) {
  const int b = blockIdx.z, d = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kConvTile + threadIdx.x * kConvElems;
  const int L = p.seqlen;
  const T* xrow = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.x_batch_stride + (int64_t)d * p.x_dim_stride;
  const T* irow = p.initial_states ? reinterpret_cast<const T*>(p.initial_states) +
                                         (int64_t)b * p.init_batch_stride + (int64_t)d * p.init_dim_stride
                                   : nullptr;
  float w[W];
#pragma unroll
  for (int k = 0; k < W; ++k) w[k] = p.weight[d * W + k];
  const float bias = p.bias ? p.bias[d] : 0.f;

  float v[kConvElems];
  load_items<T, kConvElems, kVec>(xrow, t, L, v);
  // halo: x[t-3], x[t-2], x[t-1]
  float hist[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) hist[j] = __shfl_up_sync(0xffffffffu, v[kConvElems - 3 + j], 1);
  if (lane == 0 && t < L) {
#pragma unroll
    for (int j = 0; j < 3; ++j) hist[j] = (3 - j <= W - 1) ? conv_hist<T>(xrow, irow, t - 3 + j, W) : 0.f;
  }
  if (t >= L) return;
  float xx[kConvElems + 3];
#pragma unroll
  for (int j = 0; j < 3; ++j) xx[j] = hist[j];
#pragma unroll
  for (int i = 0; i < kConvElems; ++i) xx[3 + i] = v[i];
  float out[kConvElems];
#pragma unroll
  for (int i = 0; i < kConvElems; ++i) {
    float acc = bias;
#pragma unroll
    for (int k = 0; k < W; ++k) acc = fmaf(w[k], xx[3 + i - (W - 1) + k], acc);
    out[i] = p.silu ? silu_f(acc) : acc;
  }
  T* orow = reinterpret_cast<T*>(p.out) + (int64_t)b * p.out_batch_stride + (int64_t)d * p.out_dim_stride;
  store_items<T, kConvElems, kVec>(orow, t, L, out);
}

// Backward.  pre_t = bias + sum_k w_k x_{t-(W-1)+k};  dpre_t = dout_t * act'(pre_t)
//   dx_s = sum_k w_k dpre_{s+(W-1)-k}      dw_k = sum_t dpre_t x_{t-(W-1)+k}      db = sum_t dpre_t
template <typename T, int W, bool kVec>
__global__ void __launch_bounds__(kConvThreads)
conv1d_bwd_kernel(const mtts_conv1d_bwd_params p) {
  __shared__ float red[kConvThreads / 32][W + 1];
  const int b = blockIdx.z, d = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * kConvTile + threadIdx.x * kConvElems;
  const int L = p.seqlen;
  const T* xrow = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.x_batch_stride + (int64_t)d * p.x_dim_stride;
  const T* grow = reinterpret_cast<const T*>(p.dout) + (int64_t)b * p.dout_batch_stride +
                  (int64_t)d * p.dout_dim_stride;
  const T* irow = p.initial_states ? reinterpret_cast<const T*>(p.initial_states) +
                                         (int64_t)b * p.init_batch_stride + (int64_t)d * p.init_dim_stride
                                   : nullptr;
  float w[W];
#pragma unroll
  for (int k = 0; k < W; ++k) w[k] = p.weight[d * W + k];
  const float bias = p.bias ? p.bias[d] : 0.f;

  float v[kConvElems], g[kConvElems];
  load_items<T, kConvElems, kVec>(xrow, t, L, v);
  load_items<T, kConvElems, kVec>(grow, t, L, g);
  float hist[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) hist[j] = __shfl_up_sync(0xffffffffu, v[kConvElems - 3 + j], 1);
  if (lane == 0 && t < L) {
#pragma unroll
    for (int j = 0; j < 3; ++j) hist[j] = (3 - j <= W - 1) ? conv_hist<T>(xrow, irow, t - 3 + j, W) : 0.f;
  }
  // xx[j] = x[t - 3 + j], j in [0, 8 + 3 + 3): own 8, 3 behind, 3 ahead (ahead only for lane 31)
  float xx[kConvElems + 3];
#pragma unroll
  for (int j = 0; j < 3; ++j) xx[j] = hist[j];
#pragma unroll
  for (int i = 0; i < kConvElems; ++i) xx[3 + i] = v[i];

  // dpre for own timesteps, + 3 ahead (from the next lane, or recomputed by lane 31)
  float dp[kConvElems + 3];
  float dw[W], db = 0.f;
#pragma unroll
  for (int k = 0; k < W; ++k) dw[k] = 0.f;
#pragma unroll
  for (int i = 0; i < kConvElems; ++i) {
    float pre = bias;
#pragma unroll
    for (int k = 0; k < W; ++k) pre = fmaf(w[k], xx[3 + i - (W - 1) + k], pre);
    float gi = g[i];
    if (p.silu) {
      const float sig = sigmoid_f(pre);
      gi *= sig * fmaf(pre, 1.f - sig, 1.f);
    }
    if (t + i >= L) gi = 0.f;
    dp[i] = gi;
    db += gi;
#pragma unroll
    for (int k = 0; k < W; ++k) dw[k] = fmaf(gi, xx[3 + i - (W - 1) + k], dw[k]);
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) dp[kConvElems + j] = __shfl_down_sync(0xffffffffu, dp[j], 1);
  if (lane == 31) {
    // the next warp / CTA owns t+8..t+10: recompute their dpre from global (rare: 1 lane in 32)
    float xa[6];  // x[t+5 .. t+10]
#pragma unroll
    for (int j = 0; j < 3; ++j) xa[j] = v[5 + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) xa[3 + j] = (t + 8 + j < L) ? Io<T>::to_f(xrow[t + 8 + j]) : 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float gi = 0.f;
      if (t + 8 + j < L && j < W - 1) {
        float pre = bias;
#pragma unroll
        for (int k = 0; k < W; ++k) pre = fmaf(w[k], xa[3 + j - (W - 1) + k], pre);
        gi = Io<T>::to_f(grow[t + 8 + j]);
        if (p.silu) {
          const float sig = sigmoid_f(pre);
          gi *= sig * fmaf(pre, 1.f - sig, 1.f);
        }
      }
      dp[kConvElems + j] = gi;
    }
  }
  if (t < L) {
    float dx[kConvElems];
#pragma unroll
    for (int i = 0; i < kConvElems; ++i) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < W; ++k) acc = fmaf(w[k], dp[i + (W - 1) - k], acc);
      dx[i] = acc;
    }
    T* dxrow = reinterpret_cast<T*>(p.dx) + (int64_t)b * p.dx_batch_stride + (int64_t)d * p.dx_dim_stride;
    store_items<T, kConvElems, kVec>(dxrow, t, L, dx);
  }
  // parameter gradients: warp -> CTA -> one RED per (channel, tap) per CTA
#pragma unroll
  for (int k = 0; k < W; ++k) dw[k] = warp_sum(dw[k]);
  db = warp_sum(db);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < W; ++k) red[warp][k] = dw[k];
    red[warp][W] = db;
  }
  __syncthreads();
  if (threadIdx.x <= W) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kConvThreads / 32; ++wv) s += red[wv][threadIdx.x];
    if (threadIdx.x < W) atomicAdd(p.dweight + d * W + threadIdx.x, s);
    else if (p.dbias) atomicAdd(p.dbias + d, s);
  }
}

template <typename T>
__global__ void conv1d_update_kernel(const mtts_conv1d_update_params p) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (d >= p.dim) return;
  const int W = p.width;
  T* st = reinterpret_cast<T*>(p.conv_state) + ((int64_t)b * p.dim + d) * W;
  const T xin = reinterpret_cast<const T*>(p.x)[(int64_t)b * p.x_batch_stride + d];
  float win[MTTS_MAX_CONV_WIDTH];
#pragma unroll
  for (int k = 0; k < MTTS_MAX_CONV_WIDTH; ++k)
    if (k < W - 1) win[k] = Io<T>::to_f(st[k + 1]);
  float acc = p.bias ? p.bias[d] : 0.f;
#pragma unroll
  for (int k = 0; k < MTTS_MAX_CONV_WIDTH; ++k) {
    if (k < W - 1) {
      acc = fmaf(p.weight[d * W + k], win[k], acc);
      st[k] = Io<T>::from_f(win[k]);
    }
  }
  acc = fmaf(p.weight[d * W + W - 1], Io<T>::to_f(xin), acc);
  st[W - 1] = xin;
  reinterpret_cast<T*>(p.out)[(int64_t)b * p.out_batch_stride + d] =
      Io<T>::from_f(p.silu ? silu_f(acc) : acc);
}

template <typename T>
static inline bool conv_vec_ok(const void* ptr, int64_t s0, int64_t s1, int L) {
  constexpr int VE = Io<T>::kVecElems;
  return aligned16(ptr) && s0 % VE == 0 && s1 % VE == 0 && L % VE == 0;
}

template <typename T, int W>
static int launch_conv_fwd(const mtts_conv1d_fwd_params& p, cudaStream_t s) {
  const bool vec = conv_vec_ok<T>(p.x, p.x_batch_stride, p.x_dim_stride, p.seqlen) &&
                   conv_vec_ok<T>(p.out, p.out_batch_stride, p.out_dim_stride, p.seqlen);
  const dim3 grid((p.seqlen + kConvTile - 1) / kConvTile, p.dim, p.batch);
  if (vec) conv1d_fwd_kernel<T, W, true><<<grid, kConvThreads, 0, s>>>(p);
  else conv1d_fwd_kernel<T, W, false><<<grid, kConvThreads, 0, s>>>(p);
  return launch_status();
}
template <typename T, int W>
static int launch_conv_bwd(const mtts_conv1d_bwd_params& p, cudaStream_t s) {
  const bool vec = conv_vec_ok<T>(p.x, p.x_batch_stride, p.x_dim_stride, p.seqlen) &&
                   conv_vec_ok<T>(p.dout, p.dout_batch_stride, p.dout_dim_stride, p.seqlen) &&
                   conv_vec_ok<T>(p.dx, p.dx_batch_stride, p.dx_dim_stride, p.seqlen);
  const dim3 grid((p.seqlen + kConvTile - 1) / kConvTile, p.dim, p.batch);
  if (vec) conv1d_bwd_kernel<T, W, true><<<grid, kConvThreads, 0, s>>>(p);
  else conv1d_bwd_kernel<T, W, false><<<grid, kConvThreads, 0, s>>>(p);
  return launch_status();
}
template <typename T>
static int dispatch_conv_fwd(const mtts_conv1d_fwd_params& p, cudaStream_t s) {
  switch (p.width) {
    case 2: return launch_conv_fwd<T, 2>(p, s);
    case 3: return launch_conv_fwd<T, 3>(p, s);
    case 4: return launch_conv_fwd<T, 4>(p, s);
    default: return MTTS_ERR_SHAPE;
  }
}
template <typename T>
static int dispatch_conv_bwd(const mtts_conv1d_bwd_params& p, cudaStream_t s) {
  switch (p.width) {
    case 2: return launch_conv_bwd<T, 2>(p, s);
    case 3: return launch_conv_bwd<T, 3>(p, s);
    case 4: return launch_conv_bwd<T, 4>(p, s);
    default: return MTTS_ERR_SHAPE;
  }
}

}  // namespace mtts

extern "C" int mtts_causal_conv1d_fwd(const mtts_conv1d_fwd_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->weight || !p->out) return MTTS_ERR_NULL;
  if (p->width < 2 || p->width > MTTS_MAX_CONV_WIDTH || p->batch < 0 || p->dim < 0 ||
      p->seqlen < 0 || p->batch > 65535 || p->dim > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->dim == 0 || p->seqlen == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::dispatch_conv_fwd<float>(*p, s);
    case MTTS_BF16: return mtts::dispatch_conv_fwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_causal_conv1d_bwd(const mtts_conv1d_bwd_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->weight || !p->dout || !p->dx || !p->dweight) return MTTS_ERR_NULL;
  if (p->width < 2 || p->width > MTTS_MAX_CONV_WIDTH || p->batch < 0 || p->dim < 0 ||
      p->seqlen < 0 || p->batch > 65535 || p->dim > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->dim == 0 || p->seqlen == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::dispatch_conv_bwd<float>(*p, s);
    case MTTS_BF16: return mtts::dispatch_conv_bwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_causal_conv1d_update(const mtts_conv1d_update_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->conv_state || !p->weight || !p->out) return MTTS_ERR_NULL;
  if (p->width < 2 || p->width > MTTS_MAX_CONV_WIDTH || p->batch < 0 || p->dim < 0 ||
      p->batch > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->dim == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid((p->dim + 127) / 128, p->batch);
  switch (p->io_dtype) {
    case MTTS_F32: mtts::conv1d_update_kernel<float><<<grid, 128, 0, s>>>(*p); break;
    case MTTS_BF16: mtts::conv1d_update_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(*p); break;
    default: return MTTS_ERR_DTYPE;
  }
  return mtts::launch_status();
}
