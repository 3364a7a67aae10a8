// LengthRegulator for sm_100a: phoneme-level rows repeated by their (rounded) durations into frame-level rows.
// Replaces the Python double loop of /root/reference/style_cross_attention.py:185-196 (one .item() host sync
// per phoneme) -- SURVEY.md 8f-3, the caller-side component next to the decoder.  Math: see mamba_tts_b200.h.
//
//   forward   CTA = (64 frames, batch element): the row of durations is rounded (half to even, like
//             torch.round), clamped at 0 and prefix-summed in shared memory (a few hundred values, redone by
//             every CTA of the batch element: cheaper than a second launch); warp = one frame at a time:
//             binary search of the frame in the prefix sums, then a 16-byte-vector copy of the phoneme row
//             (zeros past the end of the sequence).  HBM-bound gather: e * B * max_len * D written, the rows
//             read come from L2 (each is re-read duration times in a row).
//   backward  warp = one phoneme: its frames are contiguous, so d hidden[b, t] is a plain sum over
//             [cum[t-1], cum[t]) of d expanded -- no atomics.
#include "common.cuh"

namespace mtts {
namespace {

constexpr int kLrThreads = 256;
constexpr int kLrFramesPerCta = 64;

// inclusive prefix sums of max(rint(d), 0) over one row of `n` durations, in shared memory; returns the total
__device__ int cumsum_durations(const float* __restrict__ d, int n, int* cum, int* part) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (n + kLrThreads - 1) / kLrThreads;
  const int lo = min(tid * per, n), hi = min(lo + per, n);
  int s = 0;
  for (int i = lo; i < hi; ++i) {
    s += max(__float2int_rn(d[i]), 0);  // round half to even = torch.round
    cum[i] = s;
  }
  // exclusive scan of the per-thread sums
  int v = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) part[warp] = v;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += part[w];
  base += v - s;
  for (int i = lo; i < hi; ++i) cum[i] += base;
  int total = 0;
  for (int w = 0; w < kLrThreads / 32; ++w) total += part[w];
  __syncthreads();
  return total;
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(kLrThreads)
length_regulate_fwd_kernel(const mtts_length_regulate_fwd_params p) {
  extern __shared__ int lr_smem[];
  int* cum = lr_smem;
  __shared__ int part[kLrThreads / 32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Tn = p.t_text, D = p.dim;
  const int total = cumsum_durations(p.durations + (int64_t)b * Tn, Tn, cum, part);
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.output_lengths) p.output_lengths[b] = total;
  const T* hid = reinterpret_cast<const T*>(p.hidden) + (int64_t)b * Tn * D;
  T* out = reinterpret_cast<T*>(p.expanded) + (int64_t)b * p.max_len * D;
  const int f0 = blockIdx.x * kLrFramesPerCta;
  for (int f = f0 + warp; f < min(f0 + kLrFramesPerCta, p.max_len); f += kLrThreads / 32) {
    int t = -1;
    if (f < total) {  // first phoneme whose end (exclusive) lies beyond f
      int lo = 0, hi = Tn - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] > f) hi = mid; else lo = mid + 1;
      }
      t = lo;
    }
    if (lane == 0 && p.frame_index) p.frame_index[(int64_t)b * p.max_len + f] = t;
    T* dst = out + (int64_t)f * D;
    if constexpr (kVec) {
      constexpr int VE = Io<T>::kVecElems;
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      for (int e = lane * VE; e < D; e += 32 * VE)
        stg16_stream(dst + e, t >= 0 ? ldg16(hid + (int64_t)t * D + e) : zero);
    } else {
      for (int e = lane; e < D; e += 32) dst[e] = t >= 0 ? hid[(int64_t)t * D + e] : Io<T>::from_f(0.f);
    }
  }
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(kLrThreads)
length_regulate_bwd_kernel(const mtts_length_regulate_bwd_params p) {
  extern __shared__ int lr_smem[];
  int* cum = lr_smem;
  __shared__ int part[kLrThreads / 32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Tn = p.t_text, D = p.dim;
  cumsum_durations(p.durations + (int64_t)b * Tn, Tn, cum, part);
  const T* go = reinterpret_cast<const T*>(p.dexpanded) + (int64_t)b * p.max_len * D;
  T* gh = reinterpret_cast<T*>(p.dhidden) + (int64_t)b * Tn * D;
  const int t = blockIdx.x * (kLrThreads / 32) + warp;
  if (t >= Tn) return;
  const int fa = t > 0 ? cum[t - 1] : 0, fb = min(cum[t], p.max_len);
  if constexpr (kVec) {
    constexpr int VE = Io<T>::kVecElems;
    for (int e = lane * VE; e < D; e += 32 * VE) {
      float acc[VE];
#pragma unroll
      for (int i = 0; i < VE; ++i) acc[i] = 0.f;
      for (int f = fa; f < fb; ++f) {
        float v[VE];
        Io<T>::unpack(ldg16_stream(go + (int64_t)f * D + e), v);
#pragma unroll
        for (int i = 0; i < VE; ++i) acc[i] += v[i];
      }
      *reinterpret_cast<uint4*>(gh + (int64_t)t * D + e) = Io<T>::pack(acc);
    }
  } else {
    for (int e = lane; e < D; e += 32) {
      float acc = 0.f;
      for (int f = fa; f < fb; ++f) acc += Io<T>::to_f(go[(int64_t)f * D + e]);
      gh[(int64_t)t * D + e] = Io<T>::from_f(acc);
    }
  }
}

template <typename T>
int launch_lr_fwd(const mtts_length_regulate_fwd_params& p, cudaStream_t s) {
  const dim3 grid((p.max_len + kLrFramesPerCta - 1) / kLrFramesPerCta, p.batch);
  const size_t smem = sizeof(int) * (size_t)max(p.t_text, 1);
  const bool vec = p.dim % Io<T>::kVecElems == 0 && aligned16(p.hidden) && aligned16(p.expanded);
  if (vec) length_regulate_fwd_kernel<T, true><<<grid, kLrThreads, smem, s>>>(p);
  else length_regulate_fwd_kernel<T, false><<<grid, kLrThreads, smem, s>>>(p);
  return launch_status();
}

template <typename T>
int launch_lr_bwd(const mtts_length_regulate_bwd_params& p, cudaStream_t s) {
  const dim3 grid((p.t_text + kLrThreads / 32 - 1) / (kLrThreads / 32), p.batch);
  const size_t smem = sizeof(int) * (size_t)max(p.t_text, 1);
  const bool vec = p.dim % Io<T>::kVecElems == 0 && aligned16(p.dexpanded) && aligned16(p.dhidden);
  if (vec) length_regulate_bwd_kernel<T, true><<<grid, kLrThreads, smem, s>>>(p);
  else length_regulate_bwd_kernel<T, false><<<grid, kLrThreads, smem, s>>>(p);
  return launch_status();
}

}  // namespace
}  // namespace mtts

extern "C" int mtts_length_regulate_fwd(const mtts_length_regulate_fwd_params* p, mtts_stream_t stream) {
  if (!p || !p->expanded || (p->t_text > 0 && (!p->hidden || !p->durations))) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->t_text < 0 || p->dim < 1 || p->max_len < 0 || p->batch > 65535 || p->t_text > 8192)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->max_len == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::launch_lr_fwd<float>(*p, s);
    case MTTS_BF16: return mtts::launch_lr_fwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_length_regulate_bwd(const mtts_length_regulate_bwd_params* p, mtts_stream_t stream) {
  if (p && (p->batch == 0 || p->t_text == 0)) return MTTS_OK;  // nothing to write
  if (!p || !p->durations || !p->dhidden || (!p->dexpanded && p->max_len > 0)) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->t_text < 0 || p->dim < 1 || p->max_len < 0 || p->batch > 65535 || p->t_text > 8192)
    return MTTS_ERR_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::launch_lr_bwd<float>(*p, s);
    case MTTS_BF16: return mtts::launch_lr_bwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}
