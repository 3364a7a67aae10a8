// Single-token (decode_step) kernels for sm_100a.
//   mtts_selective_state_update -- replaces mamba_ssm's Triton selective_state_update
//   mtts_mamba_decode_step      -- conv-update + x_proj + dt_proj + state update + gate, ONE launch
//   mtts_cross_attn_decode      -- 1-query attention against the cached K/V of [ref || text]
// All reached from MambaTTSDecoder.decode_step (mamba_decoder.py:188-256) -> layer (:59-89).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mtts {

// ------------------------------------------------------------------------------------------------
// selective_state_update: one thread per (batch, channel), dstate states in a register loop.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) state_update_kernel(const mtts_state_update_params p) {
  extern __shared__ float sm_bc[];  // B then C of this batch element
  const int b = blockIdx.y, N = p.dstate;
  const T* Bp = reinterpret_cast<const T*>(p.B) + (int64_t)b * p.B_batch_stride;
  const T* Cp = reinterpret_cast<const T*>(p.C) + (int64_t)b * p.C_batch_stride;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    sm_bc[n] = Io<T>::to_f(Bp[n]);
    sm_bc[N + n] = Io<T>::to_f(Cp[n]);
  }
  __syncthreads();
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= p.dim) return;
  const float x = Io<T>::to_f(reinterpret_cast<const T*>(p.x)[(int64_t)b * p.x_batch_stride + d]);
  float dt = Io<T>::to_f(reinterpret_cast<const T*>(p.dt)[(int64_t)b * p.dt_batch_stride + d]);
  if (p.dt_bias) dt += p.dt_bias[d];
  if (p.dt_softplus) dt = softplus_f(dt);
  float* st = p.state + ((int64_t)b * p.dim + d) * N;
  const float* Ar = p.A + (int64_t)d * N;
  const float dtx = dt * x, dt2 = dt * kLog2e;
  float y = p.D ? p.D[d] * x : 0.f;
  if ((N & 3) == 0) {
    for (int n = 0; n < N; n += 4) {
      float4 h = *reinterpret_cast<float4*>(st + n);
      const float4 a = __ldg(reinterpret_cast<const float4*>(Ar + n));
      h.x = fmaf(ex2f(dt2 * a.x), h.x, dtx * sm_bc[n]);
      h.y = fmaf(ex2f(dt2 * a.y), h.y, dtx * sm_bc[n + 1]);
      h.z = fmaf(ex2f(dt2 * a.z), h.z, dtx * sm_bc[n + 2]);
      h.w = fmaf(ex2f(dt2 * a.w), h.w, dtx * sm_bc[n + 3]);
      *reinterpret_cast<float4*>(st + n) = h;
      y = fmaf(h.x, sm_bc[N + n], y);
      y = fmaf(h.y, sm_bc[N + n + 1], y);
      y = fmaf(h.z, sm_bc[N + n + 2], y);
      y = fmaf(h.w, sm_bc[N + n + 3], y);
    }
  } else {
    for (int n = 0; n < N; ++n) {
      const float h = fmaf(ex2f(dt2 * Ar[n]), st[n], dtx * sm_bc[n]);
      st[n] = h;
      y = fmaf(h, sm_bc[N + n], y);
    }
  }
  if (p.z) y *= silu_f(Io<T>::to_f(reinterpret_cast<const T*>(p.z)[(int64_t)b * p.z_batch_stride + d]));
  reinterpret_cast<T*>(p.out)[(int64_t)b * p.out_batch_stride + d] = Io<T>::from_f(y);
}

// ------------------------------------------------------------------------------------------------
// Fused decode step.  One thread-block CLUSTER per batch element: the S CTAs of a cluster split the
// d_inner channels; x_proj is a split-K product whose partial sums are exchanged through distributed
// shared memory, so the whole Mamba.step inner part is one launch with no global round trip.
// ------------------------------------------------------------------------------------------------
constexpr int kStepThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kStepThreads) decode_step_kernel(const mtts_decode_step_params p,
                                                                   const int cpc) {
  cg::cluster_group cluster = cg::this_cluster();
  const int S = cluster.num_blocks();
  const int rank = cluster.block_rank();
  const int b = blockIdx.y;
  const int N = p.dstate, R = p.dt_rank, W = p.width, J = R + 2 * N;
  const int c_lo = rank * cpc, c_hi = min(p.dim, c_lo + cpc);
  const int nown = max(0, c_hi - c_lo);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kStepThreads / 32;

  extern __shared__ __align__(16) float sm[];
  // 4-element vector access is possible when every slice start and length is a multiple of 4
  const bool vec4 = (cpc % 4 == 0) && (p.dim % 4 == 0) && (R % 4 == 0);
  float* xs = sm;           // [cpc]  conv output of the own channels
  float* part = xs + cpc;   // [J]    own split-K partial of x_proj
  float* xdbl = part + J;   // [J]    full x_proj output (dt_low | B | C)

  // 1. conv update on the own channels (state rolled in place)
  const T* xz = reinterpret_cast<const T*>(p.xz) + (int64_t)b * p.xz_batch_stride;
  for (int i = threadIdx.x; i < nown; i += kStepThreads) {
    const int d = c_lo + i;
    T* st = reinterpret_cast<T*>(p.conv_state) + ((int64_t)b * p.dim + d) * W;
    const T xin = xz[d];
    float acc = p.conv_bias ? p.conv_bias[d] : 0.f;
#pragma unroll
    for (int k = 0; k < MTTS_MAX_CONV_WIDTH - 1; ++k) {
      if (k < W - 1) {
        const T v = st[k + 1];
        acc = fmaf(p.conv_weight[d * W + k], Io<T>::to_f(v), acc);
        st[k] = v;
      }
    }
    acc = fmaf(p.conv_weight[d * W + W - 1], Io<T>::to_f(xin), acc);
    st[W - 1] = xin;
    xs[i] = Io<T>::to_f(Io<T>::from_f(silu_f(acc)));  // activation dtype, like the reference
  }
  __syncthreads();

  // 2. split-K x_proj: part[j] = sum_{d own} Wx[j, d] * xs[d]
  const T* Wx = reinterpret_cast<const T*>(p.x_proj_w);
  if (vec4) {
    // 4 rows per warp iteration, 4 elements per lane per load: all loads of a row group in flight
    for (int j0 = warp * 4; j0 < J; j0 += kWarps * 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = lane * 4; i < nown; i += 128) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + i);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (j0 + r < J) {
            float wv[4];
            load4<T>(Wx + (int64_t)(j0 + r) * p.dim + c_lo + i, wv);
            acc[r] = fmaf(wv[0], xv.x, fmaf(wv[1], xv.y, fmaf(wv[2], xv.z, fmaf(wv[3], xv.w, acc[r]))));
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float t = warp_sum(acc[r]);
        if (lane == 0 && j0 + r < J) part[j0 + r] = t;
      }
    }
  } else {
    for (int j = warp; j < J; j += kWarps) {
      const T* wr = Wx + (int64_t)j * p.dim + c_lo;
      float acc = 0.f;
      for (int i = lane; i < nown; i += 32) acc = fmaf(Io<T>::to_f(wr[i]), xs[i], acc);
      acc = warp_sum(acc);
      if (lane == 0) part[j] = acc;
    }
  }
  cluster.sync();

  // 3. all-gather-reduce the partials over the cluster through DSMEM
  for (int j = threadIdx.x; j < J; j += kStepThreads) {
    float s = 0.f;
    for (int r = 0; r < S; ++r) s += cluster.map_shared_rank(part, r)[j];
    // the reference rounds x_proj's output to the activation dtype before dt_proj / the update
    xdbl[j] = Io<T>::to_f(Io<T>::from_f(s));
  }
  __syncthreads();

  // 4./5. dt_proj + softplus + state update + gate on the own channels
  const T* Wdt = reinterpret_cast<const T*>(p.dt_proj_w);
  for (int i = threadIdx.x; i < nown; i += kStepThreads) {
    const int d = c_lo + i;
    float dt = 0.f;
    const T* wr = Wdt + (int64_t)d * R;
    if (vec4) {
#pragma unroll 4
      for (int r = 0; r < R; r += 4) {
        float wv[4];
        load4<T>(wr + r, wv);
        dt = fmaf(wv[0], xdbl[r], fmaf(wv[1], xdbl[r + 1], fmaf(wv[2], xdbl[r + 2], fmaf(wv[3], xdbl[r + 3], dt))));
      }
    } else {
      for (int r = 0; r < R; ++r) dt = fmaf(Io<T>::to_f(wr[r]), xdbl[r], dt);
    }
    dt = Io<T>::to_f(Io<T>::from_f(dt));
    dt = softplus_f(dt + p.dt_bias[d]);
    const float x = xs[i];
    const float dtx = dt * x, dt2 = dt * kLog2e;
    float* st = p.ssm_state + ((int64_t)b * p.dim + d) * N;
    const float* Ar = p.A + (int64_t)d * N;
    float y = p.D[d] * x;
    if ((N & 3) == 0) {
      for (int n = 0; n < N; n += 4) {
        float4 h = *reinterpret_cast<float4*>(st + n);
        const float4 a = __ldg(reinterpret_cast<const float4*>(Ar + n));
        h.x = fmaf(ex2f(dt2 * a.x), h.x, dtx * xdbl[R + n]);
        h.y = fmaf(ex2f(dt2 * a.y), h.y, dtx * xdbl[R + n + 1]);
        h.z = fmaf(ex2f(dt2 * a.z), h.z, dtx * xdbl[R + n + 2]);
        h.w = fmaf(ex2f(dt2 * a.w), h.w, dtx * xdbl[R + n + 3]);
        *reinterpret_cast<float4*>(st + n) = h;
        y = fmaf(h.x, xdbl[R + N + n], y);
        y = fmaf(h.y, xdbl[R + N + n + 1], y);
        y = fmaf(h.z, xdbl[R + N + n + 2], y);
        y = fmaf(h.w, xdbl[R + N + n + 3], y);
      }
    } else {
      for (int n = 0; n < N; ++n) {
        const float h = fmaf(ex2f(dt2 * Ar[n]), st[n], dtx * xdbl[R + n]);
        st[n] = h;
        y = fmaf(h, xdbl[R + N + n], y);
      }
    }
    y *= silu_f(Io<T>::to_f(xz[p.dim + d]));
    reinterpret_cast<T*>(p.y)[(int64_t)b * p.y_batch_stride + d] = Io<T>::from_f(y);
  }
  // no CTA may exit while a peer can still read its `part`
  cluster.sync();
}

// ------------------------------------------------------------------------------------------------
// Fused decode step for the decoder's shapes (width 4, dstate 16, dt_rank <= 64, batch >= 32): a cluster of
// S CTAs per batch element, each owning dim / S = CPT * 256 channels.  Against the generic kernel above:
// the x_proj rows are requested BEFORE the conv phase (they do not depend on it), every phase issues all of
// its loads before its first store (the in-place state updates otherwise serialise the round trips), and
// the x_proj exchange is the only cluster barrier.
// ------------------------------------------------------------------------------------------------
constexpr int kStepFastThreads = 256;

template <typename T, int CPT, int S>
__global__ void __launch_bounds__(kStepFastThreads) decode_step_fast_kernel(const mtts_decode_step_params p) {
  constexpr int VE = Io<T>::kVecElems;
  constexpr int kWarps = kStepFastThreads / 32;
  constexpr int Dc = CPT * kStepFastThreads;     // channels of this CTA
  constexpr int NV = Dc / VE / 32;               // 16-byte vectors of one x_proj row slice per lane
  constexpr int N = 16, W = 4;
  constexpr int kMaxJPW = 96 / kWarps;           // x_proj rows per warp (J <= 96)
  __shared__ __align__(16) float xs[Dc];         // conv output of the own channels
  __shared__ float part[96];                     // own split-K partial of x_proj
  __shared__ float xdbl[96];                     // full x_proj output (dt_low | B | C)

  const int b = blockIdx.y;
  int rank = 0;
  if constexpr (S > 1) rank = (int)cg::this_cluster().block_rank();
  const int c_lo = rank * Dc, Dm = S * Dc;
  const int R = p.dt_rank, J = R + 2 * N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jpw = (J + kWarps - 1) / kWarps;

  // x_proj row slices of this warp: j = warp + kWarps * r
  const T* Wx = reinterpret_cast<const T*>(p.x_proj_w);
  uint4 wx[kMaxJPW][NV];
#pragma unroll
  for (int r = 0; r < kMaxJPW; ++r) {
    const int j = warp + kWarps * r;
#pragma unroll
    for (int v = 0; v < NV; ++v)
      wx[r][v] = (r < jpw && j < J) ? ldg16(Wx + (int64_t)j * Dm + c_lo + (v * 32 + lane) * VE)
                                    : make_uint4(0u, 0u, 0u, 0u);
  }

  // 1. conv update (state rolled in place)
  const T* xz = reinterpret_cast<const T*>(p.xz) + (int64_t)b * p.xz_batch_stride;
  {
    float sv[CPT][4], xin[CPT], cb[CPT];
    float4 w[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int d = c_lo + tid + c * kStepFastThreads;
      load4<T>(reinterpret_cast<const T*>(p.conv_state) + ((int64_t)b * Dm + d) * W, sv[c]);
      xin[c] = Io<T>::to_f(xz[d]);
      w[c] = __ldg(reinterpret_cast<const float4*>(p.conv_weight) + d);
      cb[c] = p.conv_bias ? p.conv_bias[d] : 0.f;
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int i = tid + c * kStepFastThreads, d = c_lo + i;
      float acc = cb[c];
      acc = fmaf(w[c].x, sv[c][1], acc);
      acc = fmaf(w[c].y, sv[c][2], acc);
      acc = fmaf(w[c].z, sv[c][3], acc);
      acc = fmaf(w[c].w, xin[c], acc);
      const float nv[4] = {sv[c][1], sv[c][2], sv[c][3], xin[c]};
      store4<T>(reinterpret_cast<T*>(p.conv_state) + ((int64_t)b * Dm + d) * W, nv);
      xs[i] = Io<T>::to_f(Io<T>::from_f(silu_f(acc)));  // activation dtype, like the reference
    }
  }
  __syncthreads();

  // 2. split-K x_proj: each warp owns jpw rows, lanes split the CTA's channels
  {
    float acc[kMaxJPW];
#pragma unroll
    for (int r = 0; r < kMaxJPW; ++r) acc[r] = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float xv[VE];
      const float* xp = xs + (v * 32 + lane) * VE;
#pragma unroll
      for (int i = 0; i < VE; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(xp + i);
        xv[i] = t.x; xv[i + 1] = t.y; xv[i + 2] = t.z; xv[i + 3] = t.w;
      }
#pragma unroll
      for (int r = 0; r < kMaxJPW; ++r) {
        if (r < jpw) {
          float wv[VE];
          Io<T>::unpack(wx[r][v], wv);
#pragma unroll
          for (int i = 0; i < VE; ++i) acc[r] = fmaf(wv[i], xv[i], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kMaxJPW; ++r) {
      if (r < jpw) {
        const float t = warp_sum(acc[r]);
        const int j = warp + kWarps * r;
        if (lane == 0 && j < J) part[j] = t;
      }
    }
  }

  // 3. operands of the last phase do not depend on x_proj: request them before the exchange
  const T* Wdt = reinterpret_cast<const T*>(p.dt_proj_w);
  constexpr int kMaxRV = 64 / VE;  // dt_proj row vectors (dt_rank <= 64)
  uint4 wd[CPT][kMaxRV];
  float4 hst[CPT][4], av[CPT][4];
  float zg[CPT], dtb[CPT], Dv[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int d = c_lo + tid + c * kStepFastThreads;
#pragma unroll
    for (int v = 0; v < kMaxRV; ++v)
      wd[c][v] = v * VE < R ? ldg16(Wdt + (int64_t)d * R + v * VE) : make_uint4(0u, 0u, 0u, 0u);
    const float4* st4 = reinterpret_cast<const float4*>(p.ssm_state + ((int64_t)b * Dm + d) * N);
    const float4* a4 = reinterpret_cast<const float4*>(p.A + (int64_t)d * N);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hst[c][k] = st4[k];
      av[c][k] = __ldg(a4 + k);
    }
    zg[c] = Io<T>::to_f(xz[Dm + d]);
    dtb[c] = p.dt_bias[d];
    Dv[c] = p.D[d];
  }

  // The kernel's own loads are all in flight now and HBM is idle from here to the cross-attention two launches
  // later: pull this batch element's cached K / V into L2.  Requested here rather than at the end, because a
  // kernel does not retire before its prefetches have landed -- this way they overlap the exchange, the state
  // update and the stores.
  if (p.prefetch_a) {
    const char* pa = reinterpret_cast<const char*>(p.prefetch_a) + (int64_t)b * p.prefetch_bytes;
    const char* pb = reinterpret_cast<const char*>(p.prefetch_b) + (int64_t)b * p.prefetch_bytes;
    constexpr int kChunk = 16384;  // bytes per bulk request
    if (p.prefetch_bytes % (S * kChunk) == 0 && ((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 15) == 0) {
      // TMA bulk prefetch: a handful of requests per CTA instead of one LSU request per line
      const int per_cta = (int)(p.prefetch_bytes / S), nchunk = per_cta / kChunk;
      if (tid < 2 * nchunk) {
        const char* src = (tid < nchunk ? pa : pb) + (int64_t)rank * per_cta + (int64_t)(tid % nchunk) * kChunk;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(kChunk) : "memory");
      }
    } else {
      for (int64_t off = (int64_t)(rank * kStepFastThreads + tid) * 128; off < p.prefetch_bytes;
           off += (int64_t)S * kStepFastThreads * 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + off));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + off));
      }
    }
  }
  // all-gather-reduce the partials (over the cluster through DSMEM when S > 1)
  if constexpr (S > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    for (int j = tid; j < J; j += kStepFastThreads) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < S; ++r) t += cluster.map_shared_rank(part, r)[j];
      // the reference rounds x_proj's output to the activation dtype before dt_proj / the update
      xdbl[j] = Io<T>::to_f(Io<T>::from_f(t));
    }
  } else {
    __syncthreads();
    for (int j = tid; j < J; j += kStepFastThreads) xdbl[j] = Io<T>::to_f(Io<T>::from_f(part[j]));
  }
  __syncthreads();

  // 4. dt_proj + softplus + state update + gate
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int i = tid + c * kStepFastThreads, d = c_lo + i;
    float dt = 0.f;
#pragma unroll
    for (int v = 0; v < kMaxRV; ++v) {
      if (v * VE < R) {
        float wv[VE];
        Io<T>::unpack(wd[c][v], wv);
#pragma unroll
        for (int q = 0; q < VE; ++q) dt = fmaf(wv[q], xdbl[v * VE + q], dt);
      }
    }
    dt = Io<T>::to_f(Io<T>::from_f(dt));
    dt = softplus_f(dt + dtb[c]);
    const float x = xs[i];
    const float dtx = dt * x, dt2 = dt * kLog2e;
    float4* st4 = reinterpret_cast<float4*>(p.ssm_state + ((int64_t)b * Dm + d) * N);
    float y = Dv[c] * x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 h = hst[c][k];
      const float4 a = av[c][k];
      const int n = 4 * k;
      h.x = fmaf(ex2f(dt2 * a.x), h.x, dtx * xdbl[R + n]);
      h.y = fmaf(ex2f(dt2 * a.y), h.y, dtx * xdbl[R + n + 1]);
      h.z = fmaf(ex2f(dt2 * a.z), h.z, dtx * xdbl[R + n + 2]);
      h.w = fmaf(ex2f(dt2 * a.w), h.w, dtx * xdbl[R + n + 3]);
      st4[k] = h;
      y = fmaf(h.x, xdbl[R + N + n], y);
      y = fmaf(h.y, xdbl[R + N + n + 1], y);
      y = fmaf(h.z, xdbl[R + N + n + 2], y);
      y = fmaf(h.w, xdbl[R + N + n + 3], y);
    }
    y *= silu_f(zg[c]);
    reinterpret_cast<T*>(p.y)[(int64_t)b * p.y_batch_stride + d] = Io<T>::from_f(y);
  }
  // no CTA may exit while a peer can still read its `part`
  if constexpr (S > 1) cg::this_cluster().sync();
}

// ------------------------------------------------------------------------------------------------
// Cross-attention for one query token per batch element.  CTA = (batch, head).
// Generic fallback (any head_dim); the vectorised kernel below is the hot one.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
cross_attn_decode_generic_kernel(const mtts_cross_attn_decode_params p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int kWarps = kAttnThreads / 32;
  const int h = blockIdx.x, b = blockIdx.y;
  const int dh = p.head_dim, Tk = p.t_kv, Dm = p.heads * dh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sc = sm;                 // [Tk] scores -> probabilities
  float* qs = sc + Tk;            // [dh]
  float* acc_s = qs + dh;         // [kWarps][dh]
  __shared__ float red_s[kWarps];

  const float scale = rsqrtf((float)dh);
  const T* q = reinterpret_cast<const T*>(p.q) + (int64_t)b * Dm + h * dh;
  // torch scales q before QK^T and keeps it in the activation dtype
  for (int e = threadIdx.x; e < dh; e += kAttnThreads)
    qs[e] = Io<T>::to_f(Io<T>::from_f(Io<T>::to_f(q[e]) * scale));
  __syncthreads();

  const T* Kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * Tk * Dm + h * dh;
  const T* Vb = reinterpret_cast<const T*>(p.v) + (int64_t)b * Tk * Dm + h * dh;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * Tk : nullptr;

  float lmax = -INFINITY;
  for (int t = warp; t < Tk; t += kWarps) {
    const T* kr = Kb + (int64_t)t * Dm;
    float s = 0.f;
    for (int e = lane; e < dh; e += 32) s = fmaf(Io<T>::to_f(kr[e]), qs[e], s);
    s = warp_sum(s);
    if (mk && !mk[t]) s = -INFINITY;
    if (lane == 0) sc[t] = s;
    lmax = fmaxf(lmax, s);
  }
  if (lane == 0) red_s[warp] = lmax;
  __syncthreads();
  float gmax = red_s[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) gmax = fmaxf(gmax, red_s[w]);
  __syncthreads();
  float lsum = 0.f;
  for (int t = threadIdx.x; t < Tk; t += kAttnThreads) {
    const float e = ex2f((sc[t] - gmax) * kLog2e);  // all-masked row -> NaN like torch
    sc[t] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red_s[warp] = lsum;
  __syncthreads();
  float gsum = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) gsum += red_s[w];
  const float inv = 1.f / gsum;

  // out[e] = sum_t p_t V[t, e]; warps split t, lanes split e
  for (int e0 = 0; e0 < dh; e0 += 32) {
    const int e = e0 + lane;
    float a = 0.f;
    if (e < dh)
      for (int t = warp; t < Tk; t += kWarps) a = fmaf(sc[t], Io<T>::to_f(Vb[(int64_t)t * Dm + e]), a);
    if (e < dh) acc_s[warp * dh + e] = a;
  }
  __syncthreads();
  T* out = reinterpret_cast<T*>(p.out) + (int64_t)b * Dm + h * dh;
  for (int e = threadIdx.x; e < dh; e += kAttnThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) a += acc_s[w * dh + e];
    out[e] = Io<T>::from_f(a * inv);
  }
}

// Vectorised variant: head_dim = kLPR * (16 bytes of T); a group of kLPR lanes reads one K/V row with
// one 16-byte load each, so a warp load instruction covers 32 / kLPR whole rows (full sectors).
constexpr int kAttnVecThreads = 128;

template <typename T, int kLPR>
__global__ void __launch_bounds__(kAttnVecThreads)
cross_attn_decode_vec_kernel(const mtts_cross_attn_decode_params p) {
  constexpr int VE = Io<T>::kVecElems;
  constexpr int kWarps = kAttnVecThreads / 32;
  constexpr int kRPW = 32 / kLPR;          // rows per warp load
  constexpr int kDh = kLPR * VE;
  extern __shared__ __align__(16) float sm[];
  float* sc = sm;                          // [Tk]
  float* acc_s = sc + p.t_kv;              // [kWarps][kDh]
  __shared__ float red_s[kWarps];

  const int h = blockIdx.x, b = blockIdx.y;
  const int Tk = p.t_kv, Dm = p.heads * kDh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % kLPR, rsel = lane / kLPR;
  const float scale = rsqrtf((float)kDh);

  float qv[VE];
  {
    const T* q = reinterpret_cast<const T*>(p.q) + (int64_t)b * Dm + h * kDh + sub * VE;
    Io<T>::unpack(ldg16(q), qv);
#pragma unroll
    for (int j = 0; j < VE; ++j) qv[j] = Io<T>::to_f(Io<T>::from_f(qv[j] * scale));
  }
  const T* Kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * Tk * Dm + h * kDh + sub * VE;
  const T* Vb = reinterpret_cast<const T*>(p.v) + (int64_t)b * Tk * Dm + h * kDh + sub * VE;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * Tk : nullptr;

  float lmax = -INFINITY;
#pragma unroll 4
  for (int t0 = warp * kRPW; t0 < Tk; t0 += kWarps * kRPW) {
    const int t = t0 + rsel;
    float s = 0.f;
    if (t < Tk) {
      float kv[VE];
      Io<T>::unpack(ldg16_stream(Kb + (int64_t)t * Dm), kv);
#pragma unroll
      for (int j = 0; j < VE; ++j) s = fmaf(kv[j], qv[j], s);
    }
#pragma unroll
    for (int o = kLPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (t < Tk) {
      if (mk && !mk[t]) s = -INFINITY;
      if (sub == 0) sc[t] = s;
      lmax = fmaxf(lmax, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red_s[warp] = lmax;
  __syncthreads();
  float gmax = red_s[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) gmax = fmaxf(gmax, red_s[w]);
  __syncthreads();
  float lsum = 0.f;
  for (int t = threadIdx.x; t < Tk; t += kAttnVecThreads) {
    const float e = ex2f((sc[t] - gmax) * kLog2e);
    sc[t] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red_s[warp] = lsum;
  __syncthreads();
  float gsum = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) gsum += red_s[w];
  const float inv = 1.f / gsum;

  float acc[VE];
#pragma unroll
  for (int j = 0; j < VE; ++j) acc[j] = 0.f;
#pragma unroll 4
  for (int t0 = warp * kRPW; t0 < Tk; t0 += kWarps * kRPW) {
    const int t = t0 + rsel;
    if (t < Tk) {
      float vv[VE];
      Io<T>::unpack(ldg16_stream(Vb + (int64_t)t * Dm), vv);
      const float pt = sc[t];
#pragma unroll
      for (int j = 0; j < VE; ++j) acc[j] = fmaf(pt, vv[j], acc[j]);
    }
  }
#pragma unroll
  for (int o = kLPR; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < VE; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (rsel == 0) {
#pragma unroll
    for (int j = 0; j < VE; ++j) acc_s[warp * kDh + sub * VE + j] = acc[j];
  }
  __syncthreads();
  T* out = reinterpret_cast<T*>(p.out) + (int64_t)b * Dm + h * kDh;
  for (int e = threadIdx.x; e < kDh; e += kAttnVecThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) a += acc_s[w * kDh + e];
    out[e] = Io<T>::from_f(a * inv);
  }
}

// Register-cached variant for t_kv <= kIt * (rows per CTA pass): every K and V row this lane will
// ever touch is loaded up front (one memory round trip for the whole kernel), scores and V stay in
// registers, two block reductions (max, then sum + output) finish the softmax.
constexpr int kAttnCachedThreads = 256;

template <typename T, int kLPR, int kIt>
__global__ void __launch_bounds__(kAttnCachedThreads)
cross_attn_decode_cached_kernel(const mtts_cross_attn_decode_params p) {
  constexpr int VE = Io<T>::kVecElems;
  constexpr int kWarps = kAttnCachedThreads / 32;
  constexpr int kRPW = 32 / kLPR;
  constexpr int kDh = kLPR * VE;
  __shared__ float red_s[kWarps];
  __shared__ float sum_s[kWarps];
  __shared__ __align__(16) float acc_s[kWarps * kDh];

  const int h = blockIdx.x, b = blockIdx.y;
  const int Tk = p.t_kv, Dm = p.heads * kDh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % kLPR, rsel = lane / kLPR;
  const float scale = rsqrtf((float)kDh);

  const T* Kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * Tk * Dm + h * kDh + sub * VE;
  const T* Vb = reinterpret_cast<const T*>(p.v) + (int64_t)b * Tk * Dm + h * kDh + sub * VE;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * Tk : nullptr;

  uint4 kraw[kIt], vraw[kIt];
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int t = (it * kWarps + warp) * kRPW + rsel;
    if (t < Tk) {
      kraw[it] = ldg16_stream(Kb + (int64_t)t * Dm);
      vraw[it] = ldg16_stream(Vb + (int64_t)t * Dm);
    } else {
      kraw[it] = make_uint4(0, 0, 0, 0);
      vraw[it] = make_uint4(0, 0, 0, 0);
    }
  }
  float qv[VE];
  {
    const T* q = reinterpret_cast<const T*>(p.q) + (int64_t)b * Dm + h * kDh + sub * VE;
    Io<T>::unpack(ldg16(q), qv);
#pragma unroll
    for (int j = 0; j < VE; ++j) qv[j] = Io<T>::to_f(Io<T>::from_f(qv[j] * scale));
  }
  float sc[kIt];
  float lmax = -INFINITY;
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int t = (it * kWarps + warp) * kRPW + rsel;
    float kv[VE];
    Io<T>::unpack(kraw[it], kv);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VE; ++j) s = fmaf(kv[j], qv[j], s);
#pragma unroll
    for (int o = kLPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (t >= Tk || (mk && !mk[t])) s = -INFINITY;
    sc[it] = s;
    lmax = fmaxf(lmax, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red_s[warp] = lmax;
  __syncthreads();
  float gmax = red_s[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) gmax = fmaxf(gmax, red_s[w]);

  float acc[VE];
#pragma unroll
  for (int j = 0; j < VE; ++j) acc[j] = 0.f;
  float lsum = 0.f;
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int t = (it * kWarps + warp) * kRPW + rsel;
    // rows past t_kv contribute nothing; an all-masked row gives exp(nan) = nan like torch
    const float pt = (t < Tk) ? ex2f((sc[it] - gmax) * kLog2e) : 0.f;
    float vv[VE];
    Io<T>::unpack(vraw[it], vv);
#pragma unroll
    for (int j = 0; j < VE; ++j) acc[j] = fmaf(pt, vv[j], acc[j]);
    lsum += pt;  // identical on the kLPR lanes of a row group
  }
  // reduce over the row groups of the warp (lanes with equal `sub`)
#pragma unroll
  for (int o = kLPR; o < 32; o <<= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
#pragma unroll
    for (int j = 0; j < VE; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (rsel == 0) {
#pragma unroll
    for (int j = 0; j < VE; ++j) acc_s[warp * kDh + sub * VE + j] = acc[j];
  }
  if (lane == 0) sum_s[warp] = lsum;
  __syncthreads();
  float gsum = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) gsum += sum_s[w];
  const float inv = 1.f / gsum;
  T* out = reinterpret_cast<T*>(p.out) + (int64_t)b * Dm + h * kDh;
  for (int e = threadIdx.x; e < kDh; e += kAttnCachedThreads) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) a += acc_s[w * kDh + e];
    out[e] = Io<T>::from_f(a * inv);
  }
}

template <typename T, int kLPR>
static int launch_attn_vec(const mtts_cross_attn_decode_params& p, cudaStream_t s) {
  constexpr int kDh = kLPR * Io<T>::kVecElems;
  constexpr int kIt = 8;
  constexpr int kRowsCached = kIt * (kAttnCachedThreads / 32) * (32 / kLPR);
  if (p.t_kv <= kRowsCached) {
    cross_attn_decode_cached_kernel<T, kLPR, kIt>
        <<<dim3(p.heads, p.batch), kAttnCachedThreads, 0, s>>>(p);
    return launch_status();
  }
  const size_t smem = sizeof(float) * ((size_t)p.t_kv + (kAttnVecThreads / 32) * kDh);
  auto kern = cross_attn_decode_vec_kernel<T, kLPR>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -static_cast<int>(e);
  }
  kern<<<dim3(p.heads, p.batch), kAttnVecThreads, smem, s>>>(p);
  return launch_status();
}

// returns 1 when no vectorised variant matches
template <typename T>
static int dispatch_attn_vec(const mtts_cross_attn_decode_params& p, cudaStream_t s, int* rc) {
  constexpr int VE = Io<T>::kVecElems;
  if (p.head_dim % VE != 0 || !aligned16(p.q) || !aligned16(p.k) || !aligned16(p.v)) return 1;
  switch (p.head_dim / VE) {
    case 4: *rc = launch_attn_vec<T, 4>(p, s); return 0;
    case 8: *rc = launch_attn_vec<T, 8>(p, s); return 0;
    case 16: *rc = launch_attn_vec<T, 16>(p, s); return 0;
    case 32: *rc = launch_attn_vec<T, 32>(p, s); return 0;
    default: return 1;
  }
}

template <typename T, int CPT, int S>
static int launch_decode_step_fast(const mtts_decode_step_params& p, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, p.batch);
  cfg.blockDim = dim3(kStepFastThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_step_fast_kernel<T, CPT, S>, p);
  if (e != cudaSuccess) return -static_cast<int>(e);
  return launch_status();
}

template <typename T>
static int launch_decode_step(const mtts_decode_step_params& p, cudaStream_t s) {
  constexpr int VE = Io<T>::kVecElems;
  // the decoder's shapes, batch large enough that a few CTAs per batch element fill the machine
  // (measured at B 64, dim 1024 inside the replayed decode graph: S = 4 8.5 us, S = 2 9.4 us, generic
  // cluster-of-8 kernel 11.7 us)
  if (p.batch >= 32 && p.width == 4 && p.dstate == 16 && p.dt_rank % VE == 0 && p.dt_rank <= 64 &&
      aligned16(p.x_proj_w) && aligned16(p.dt_proj_w) && aligned16(p.conv_state) && aligned16(p.conv_weight)) {
    if (p.dim == 512) return launch_decode_step_fast<T, 1, 2>(p, s);
    if (p.dim == 1024) return launch_decode_step_fast<T, 1, 4>(p, s);
    if (p.dim == 2048) return launch_decode_step_fast<T, 1, 8>(p, s);
  }
  // cluster size: enough CTAs to cover the chip, at least 32 channels per CTA
  int S = 8;
  while (S > 1 && (p.batch * S > 2 * kNumSMs * 2 || p.dim / S < 32)) S >>= 1;
  const int cpc = (p.dim + S - 1) / S;
  const int J = p.dt_rank + 2 * p.dstate;
  const size_t smem = sizeof(float) * ((size_t)cpc + 2 * J);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, p.batch);
  cfg.blockDim = dim3(kStepThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  auto kern = decode_step_kernel<T>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -static_cast<int>(e);
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, cpc);
  if (e != cudaSuccess) return -static_cast<int>(e);
  return launch_status();
}


// ---------------------------------------------------------------------------------------------------------
// The whole cross-attention branch of one decode step as ONE launch (mamba_decoder.py:67-81 with T_q = 1):
//   x1 = x + delta;  hq = LN2(x1);  q = Wq hq + bq;  a = softmax(q K^T / sqrt(dh) + mask) V;
//   o = Wo a + bo;   x2 = x1 + o;   out = FiLM(LN3(x2))
// A thread-block cluster of `heads` CTAs owns one batch element, CTA = one head.  Every CTA normalises
// the (tiny) row itself, projects its own 64 query features straight from L2-resident Wq rows while its
// K / V rows are in flight from HBM, and after the attention multiplies its head's slice of Wo; the
// per-head partial rows are summed through distributed shared memory, after which every CTA holds the
// complete x2 row, normalises it and writes its own 64-column slice.  Replaces LayerNorm + q GEMM +
// attention + o GEMM + LayerNorm (5 launches, 4 of them a few microseconds of pure latency) per layer.
// Rounding points mirror the unfused path: hq, q, q*scale, a, o are rounded to the io dtype.
constexpr int kXBlockThreads = 256;

__device__ __forceinline__ float block_sum256(float v, float* red, int lane, int warp) {
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < kXBlockThreads / 32; ++w) s += red[w];
  __syncthreads();
  return s;
}

template <typename T, int kLPR, int kIt, int kHeads, int kQBatch, bool kFull>
__global__ void __launch_bounds__(kXBlockThreads, sizeof(T) == 2 ? 4 : 2)
cross_attn_block_decode_kernel(const mtts_cross_attn_block_params p) {
  constexpr int VE = Io<T>::kVecElems;
  constexpr int kWarps = kXBlockThreads / 32;
  constexpr int kRPW = 32 / kLPR;
  constexpr int kDh = kLPR * VE;
  constexpr int Dm = kHeads * kDh;
  constexpr int EPT = Dm / kXBlockThreads;      // row elements per thread
  constexpr int kRowVecs = Dm / VE;             // 16-byte vectors per Wq row
  constexpr int kQPass = (kRowVecs + 31) / 32;  // ... per lane
  constexpr int kQRows = kDh / kWarps;          // query features per warp
  constexpr int kSegVecs = kDh / VE;            // vectors per Wo row segment (= kLPR)
  constexpr int kORowsPerPass = kXBlockThreads / kSegVecs;
  constexpr int kOPass = Dm / kORowsPerPass;
  static_assert(Dm % kXBlockThreads == 0 && kDh % kWarps == 0 && kQRows % kQBatch == 0, "shape");
  static_assert(Dm % kORowsPerPass == 0 && kRowVecs % 32 == 0, "shape");

  __shared__ __align__(16) float hq_s[Dm];
  __shared__ __align__(16) float opart_s[Dm];
  __shared__ __align__(16) float q_s[kDh];
  __shared__ __align__(16) float a_s[kDh];
  __shared__ __align__(16) float acc_s[kWarps * kDh];
  __shared__ float red_s[kWarps];
  __shared__ float sum_s[kWarps];

  const int h = blockIdx.x, b = blockIdx.y;  // cluster = the kHeads CTAs of one batch element, rank == head
  const int Tk = p.t_kv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane % kLPR, rsel = lane / kLPR;
  const float scale = rsqrtf((float)kDh);

  // ---- K / V rows of this head: pulled from HBM into L2 now (no registers), loaded after the query
  //      projection -- four CTAs per SM stay resident and the whole grid is one wave ----------------------
  const T* Kb = reinterpret_cast<const T*>(p.k) + (int64_t)b * Tk * Dm + h * kDh + sub * VE;
  const T* Vb = reinterpret_cast<const T*>(p.v) + (int64_t)b * Tk * Dm + h * kDh + sub * VE;
  const uint8_t* mk = p.mask ? p.mask + (int64_t)b * Tk : nullptr;
  if (sub == 0) {  // one prefetch per 128-byte row segment (kDh * sizeof(T) bytes, line aligned for bf16 x 64)
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int t = (it * kWarps + warp) * kRPW + rsel;
      if (t < Tk) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(Kb + (int64_t)t * Dm));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(Vb + (int64_t)t * Dm));
        if (kDh * sizeof(T) > 128) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(Kb + (int64_t)t * Dm) + 128));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(Vb + (int64_t)t * Dm) + 128));
        }
      }
    }
  }

  // ---- x1 = x + delta, hq = LN2(x1) (two-pass statistics, like add_layernorm_fwd) ------------------------
  const int e0 = tid * EPT;
  float x1[EPT];
  {
    const float* xr = p.x + (int64_t)b * Dm + e0;
    const T* dl = p.delta ? reinterpret_cast<const T*>(p.delta) + (int64_t)b * Dm + e0 : nullptr;
    float s = 0.f, lw[EPT], lb[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      x1[i] = xr[i] + (dl ? Io<T>::to_f(dl[i]) : 0.f);
      lw[i] = p.lnq_weight[e0 + i];  // requested before the reductions: one round trip less
      lb[i] = p.lnq_bias[e0 + i];
      s += x1[i];
    }
    const float mean = block_sum256(s, red_s, lane, warp) / (float)Dm;
    float qd = 0.f;
#pragma unroll
    for (int i = 0; i < EPT; ++i) qd = fmaf(x1[i] - mean, x1[i] - mean, qd);
    const float rstd = rsqrtf(block_sum256(qd, red_s, lane, warp) / (float)Dm + p.eps_q);
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const float y = fmaf((x1[i] - mean) * rstd, lw[i], lb[i]);
      hq_s[e0 + i] = Io<T>::to_f(Io<T>::from_f(y));
    }
  }
  __syncthreads();

  // ---- q = Wq[h*dh .. , :] hq + bq: warp = kQRows features, lanes along the row ---------------------------
  {
    const T* wq = reinterpret_cast<const T*>(p.wq) + (int64_t)(h * kDh + warp * kQRows) * Dm;
#pragma unroll
    for (int r0 = 0; r0 < kQRows; r0 += kQBatch) {
      uint4 wv[kQBatch][kQPass];
#pragma unroll
      for (int r = 0; r < kQBatch; ++r)
#pragma unroll
        for (int ps = 0; ps < kQPass; ++ps)
          wv[r][ps] = ldg16(wq + (int64_t)(r0 + r) * Dm + (ps * 32 + lane) * VE);
#pragma unroll
      for (int r = 0; r < kQBatch; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int ps = 0; ps < kQPass; ++ps) {
          float wf[VE];
          Io<T>::unpack(wv[r][ps], wf);
          const float* hv = hq_s + (ps * 32 + lane) * VE;
#pragma unroll
          for (int i = 0; i < VE; ++i) acc = fmaf(wf[i], hv[i], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
          const int j = warp * kQRows + r0 + r;
          const float qb = acc + Io<T>::to_f(reinterpret_cast<const T*>(p.bq)[h * kDh + j]);
          q_s[j] = Io<T>::to_f(Io<T>::from_f(Io<T>::to_f(Io<T>::from_f(qb)) * scale));
        }
      }
    }
  }
  __syncthreads();

  // ---- attention over the cached rows (cross_attn_decode_cached_kernel) -----------------------------------
  uint4 kraw[kIt], vraw[kIt];
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int t = (it * kWarps + warp) * kRPW + rsel;
    if (t < Tk) {
      kraw[it] = ldg16_stream(Kb + (int64_t)t * Dm);
      vraw[it] = ldg16_stream(Vb + (int64_t)t * Dm);
    } else {
      kraw[it] = make_uint4(0, 0, 0, 0);
      vraw[it] = make_uint4(0, 0, 0, 0);
    }
  }
  float qv[VE];
#pragma unroll
  for (int j = 0; j < VE; ++j) qv[j] = q_s[sub * VE + j];
  float sc[kIt];
  float lmax = -INFINITY;
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int t = (it * kWarps + warp) * kRPW + rsel;
    float kv[VE];
    Io<T>::unpack(kraw[it], kv);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VE; ++j) s = fmaf(kv[j], qv[j], s);
#pragma unroll
    for (int o = kLPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (t >= Tk || (mk && !mk[t])) s = -INFINITY;
    sc[it] = s;
    lmax = fmaxf(lmax, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) red_s[warp] = lmax;
  __syncthreads();
  float gmax = red_s[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) gmax = fmaxf(gmax, red_s[w]);
  float acc[VE];
#pragma unroll
  for (int j = 0; j < VE; ++j) acc[j] = 0.f;
  float lsum = 0.f;
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int t = (it * kWarps + warp) * kRPW + rsel;
    const float pt = (t < Tk) ? ex2f((sc[it] - gmax) * kLog2e) : 0.f;
    float vv[VE];
    Io<T>::unpack(vraw[it], vv);
#pragma unroll
    for (int j = 0; j < VE; ++j) acc[j] = fmaf(pt, vv[j], acc[j]);
    lsum += pt;
  }
#pragma unroll
  for (int o = kLPR; o < 32; o <<= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
#pragma unroll
    for (int j = 0; j < VE; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (rsel == 0) {
#pragma unroll
    for (int j = 0; j < VE; ++j) acc_s[warp * kDh + sub * VE + j] = acc[j];
  }
  if (lane == 0) sum_s[warp] = lsum;
  __syncthreads();
  if (tid < kDh) {
    float gsum = 0.f, a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      gsum += sum_s[w];
      a += acc_s[w * kDh + tid];
    }
    const T aT = Io<T>::from_f(a * (1.f / gsum));
    a_s[tid] = Io<T>::to_f(aT);
    if constexpr (!kFull) reinterpret_cast<T*>(p.out)[(int64_t)b * Dm + h * kDh + tid] = aT;
  }
  const bool mine = (e0 / kDh) == h;  // this CTA writes its own head-sized slice of the row
  if constexpr (!kFull) {  // front half only: x_out = x1, out = attention output; o-proj and LN3 follow separately
    if (mine && p.x_out) {
#pragma unroll
      for (int i = 0; i < EPT; ++i) p.x_out[(int64_t)b * Dm + e0 + i] = x1[i];
    }
    return;
  }
  cg::cluster_group cluster = cg::this_cluster();
  __syncthreads();

  // ---- this head's share of o = Wo a: kSegVecs lanes per output row, 128-byte row segments ---------------
  {
    const int seg = tid % kSegVecs, rgrp = tid / kSegVecs;
    const T* wo = reinterpret_cast<const T*>(p.wo) + h * kDh + seg * VE;
    float av[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) av[i] = a_s[seg * VE + i];
    constexpr int kOBatch = kOPass < 8 ? kOPass : 8;
#pragma unroll
    for (int p0 = 0; p0 < kOPass; p0 += kOBatch) {
      uint4 wv[kOBatch];
#pragma unroll
      for (int ps = 0; ps < kOBatch; ++ps) wv[ps] = ldg16(wo + (int64_t)((p0 + ps) * kORowsPerPass + rgrp) * Dm);
#pragma unroll
      for (int ps = 0; ps < kOBatch; ++ps) {
        float wf[VE];
        Io<T>::unpack(wv[ps], wf);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VE; ++i) s = fmaf(wf[i], av[i], s);
#pragma unroll
        for (int o = kSegVecs / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (seg == 0) opart_s[(p0 + ps) * kORowsPerPass + rgrp] = s;
      }
    }
  }
  cluster.sync();

  // ---- x2 = x1 + (sum of the heads' partial rows + bo), complete row in every CTA ---------------------------
  float x2[EPT];
  {
    float o[EPT];
#pragma unroll
    for (int i = 0; i < EPT; ++i) o[i] = 0.f;
#pragma unroll
    for (int r = 0; r < kHeads; ++r) {
      const float* rp = cluster.map_shared_rank(opart_s, r) + e0;
#pragma unroll
      for (int i = 0; i < EPT; ++i) o[i] += rp[i];
    }
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      const float ob = o[i] + Io<T>::to_f(reinterpret_cast<const T*>(p.bo)[e0 + i]);
      x2[i] = x1[i] + Io<T>::to_f(Io<T>::from_f(ob));
    }
  }
  if (mine && p.x_out) {
#pragma unroll
    for (int i = 0; i < EPT; ++i) p.x_out[(int64_t)b * Dm + e0 + i] = x2[i];
  }
  // ---- out = FiLM(LN3(x2)) ------------------------------------------------------------------------------------
  {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPT; ++i) s += x2[i];
    const float mean = block_sum256(s, red_s, lane, warp) / (float)Dm;
    float qd = 0.f;
#pragma unroll
    for (int i = 0; i < EPT; ++i) qd = fmaf(x2[i] - mean, x2[i] - mean, qd);
    const float rstd = rsqrtf(block_sum256(qd, red_s, lane, warp) / (float)Dm + p.eps_o);
    if (mine) {
      T* out = reinterpret_cast<T*>(p.out) + (int64_t)b * Dm + e0;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        float w = p.lno_weight[e0 + i], bb = p.lno_bias[e0 + i];
        if (p.film_gamma) {  // out = gamma (xhat w + b) + beta, folded like add_layernorm_fwd
          const float gm = p.film_gamma[(int64_t)b * Dm + e0 + i], bt = p.film_beta[(int64_t)b * Dm + e0 + i];
          w *= gm;
          bb = fmaf(gm, bb, bt);
        }
        out[i] = Io<T>::from_f(fmaf((x2[i] - mean) * rstd, w, bb));
      }
    }
  }
  cluster.sync();  // nobody leaves while its partial row may still be read
}

template <typename T, int kLPR, int kIt, int kHeads, int kQBatch>
static int launch_xattn_block(const mtts_cross_attn_block_params& p, cudaStream_t s) {
  if (!p.wo) {  // front half: independent CTAs
    cross_attn_block_decode_kernel<T, kLPR, kIt, kHeads, kQBatch, false>
        <<<dim3(kHeads, p.batch), kXBlockThreads, 0, s>>>(p);
    return launch_status();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kHeads, p.batch);
  cfg.blockDim = dim3(kXBlockThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kHeads;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, cross_attn_block_decode_kernel<T, kLPR, kIt, kHeads, kQBatch, true>, p);
  if (e != cudaSuccess) return -static_cast<int>(e);
  return launch_status();
}


// ---------------------------------------------------------------------------------------------------------
// Token plumbing of the generation loop (the caller side of mamba_decoder.py:218-221,254-256): two launches
// instead of the nine ATen ones (2 gathers, add, cast, argmax, 2 copies, 2 counter increments).
//   embed:  x[b, :] = tok_embed[tok[b], :] + pos_embed[*pos, :];  *step += 1   (step = column being generated)
//   greedy: tok[b] = out[b, *step] = argmax_v logits[b, v] (lowest index on ties);  *pos += 1
// Each counter is bumped by the kernel that does NOT read it, so stream order alone makes it race-free.
__global__ void __launch_bounds__(128)
decode_embed_kernel(const mtts_decode_embed_params p) {
  const int b = blockIdx.x;
  const int64_t tk = p.tok[b], ps = *p.pos;
  const float* te = p.tok_embed + tk * p.dim;
  const float* pe = p.pos_embed + ps * p.dim;
  float* x = p.x + (int64_t)b * p.dim;
  for (int e = threadIdx.x * 4; e < p.dim; e += 128 * 4) {
    const float4 a = *reinterpret_cast<const float4*>(te + e);
    const float4 c = *reinterpret_cast<const float4*>(pe + e);
    *reinterpret_cast<float4*>(x + e) = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
  if (b == 0 && threadIdx.x == 0 && p.step) *p.step += 1;
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_greedy_kernel(const mtts_decode_greedy_params p) {
  __shared__ float bv_s[8];
  __shared__ int bi_s[8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* lg = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.vocab;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < p.vocab; v += 256) {
    const float x = Io<T>::to_f(lg[v]);
    if (x > bv || (x == bv && v < bi)) { bv = x; bi = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { bv_s[warp] = bv; bi_s[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w)
      if (bv_s[w] > bv || (bv_s[w] == bv && bi_s[w] < bi)) { bv = bv_s[w]; bi = bi_s[w]; }
    if (bi == 0x7fffffff) bi = 0;  // all NaN / -inf row
    int64_t t = bi;
    if (p.eos_id >= 0) {
      if (p.lengths[b] >= 0) t = p.pad_id;                       // finished earlier
      else if (t == p.eos_id) p.lengths[b] = (p.step ? *p.step : 0) + 1;  // finishes now, eos kept
    }
    p.tok[b] = t;
    if (p.out) p.out[(int64_t)b * p.out_stride + *p.step] = t;
    if (b == 0 && p.pos) *p.pos += 1;
  }
}
}  // namespace mtts

extern "C" int mtts_selective_state_update(const mtts_state_update_params* p, mtts_stream_t stream) {
  if (!p || !p->state || !p->x || !p->dt || !p->A || !p->B || !p->C || !p->out) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->dim < 0 || p->dstate < 1 || p->dstate > MTTS_MAX_DSTATE || p->batch > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->dim == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid((p->dim + 127) / 128, p->batch);
  const size_t smem = sizeof(float) * 2 * p->dstate;
  switch (p->io_dtype) {
    case MTTS_F32: mtts::state_update_kernel<float><<<grid, 128, smem, s>>>(*p); break;
    case MTTS_BF16: mtts::state_update_kernel<__nv_bfloat16><<<grid, 128, smem, s>>>(*p); break;
    default: return MTTS_ERR_DTYPE;
  }
  return mtts::launch_status();
}

extern "C" int mtts_mamba_decode_step(const mtts_decode_step_params* p, mtts_stream_t stream) {
  if (!p || !p->xz || !p->conv_state || !p->ssm_state || !p->conv_weight || !p->x_proj_w ||
      !p->dt_proj_w || !p->dt_bias || !p->A || !p->D || !p->y)
    return MTTS_ERR_NULL;
  if (p->batch < 0 || p->dim < 1 || p->dstate < 1 || p->dstate > MTTS_MAX_DSTATE || p->dt_rank < 1 ||
      p->width < 2 || p->width > MTTS_MAX_CONV_WIDTH || p->batch > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::launch_decode_step<float>(*p, s);
    case MTTS_BF16: return mtts::launch_decode_step<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_cross_attn_decode(const mtts_cross_attn_decode_params* p, mtts_stream_t stream) {
  if (!p || !p->q || !p->k || !p->v || !p->out) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->heads < 1 || p->head_dim < 1 || p->t_kv < 1 || p->batch > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = sizeof(float) * ((size_t)p->t_kv + p->head_dim + (mtts::kAttnThreads / 32) * p->head_dim);
  if (smem > 200 * 1024) return MTTS_ERR_SHAPE;
  int rc = 0;
  if (p->io_dtype == MTTS_F32) {
    if (mtts::dispatch_attn_vec<float>(*p, s, &rc) == 0) return rc;
  } else if (p->io_dtype == MTTS_BF16) {
    if (mtts::dispatch_attn_vec<__nv_bfloat16>(*p, s, &rc) == 0) return rc;
  } else {
    return MTTS_ERR_DTYPE;
  }
  const dim3 grid(p->heads, p->batch);
  cudaError_t e;
  if (p->io_dtype == MTTS_F32) {
    if (smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(mtts::cross_attn_decode_generic_kernel<float>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return -static_cast<int>(e);
    mtts::cross_attn_decode_generic_kernel<float><<<grid, mtts::kAttnThreads, smem, s>>>(*p);
  } else {
    if (smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(mtts::cross_attn_decode_generic_kernel<__nv_bfloat16>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return -static_cast<int>(e);
    mtts::cross_attn_decode_generic_kernel<__nv_bfloat16><<<grid, mtts::kAttnThreads, smem, s>>>(*p);
  }
  return mtts::launch_status();
}

extern "C" int mtts_cross_attn_block_decode(const mtts_cross_attn_block_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->lnq_weight || !p->lnq_bias || !p->wq || !p->bq || !p->k || !p->v || !p->out)
    return MTTS_ERR_NULL;
  if (p->wo && (!p->bo || !p->lno_weight || !p->lno_bias)) return MTTS_ERR_NULL;
  // front half: the head CTAs of a batch element are independent and each reads the whole row of x
  if (!p->wo && p->x_out == p->x) return MTTS_ERR_UNSUPPORTED;
  if ((p->film_gamma == nullptr) != (p->film_beta == nullptr)) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->t_kv < 1 || p->batch > 65535) return MTTS_ERR_SHAPE;
  // one instantiation per supported decoder shape: 8 heads x 64 features, cached rows
  if (p->heads != 8 || p->head_dim != 64) return MTTS_ERR_UNSUPPORTED;
  if (!mtts::aligned16(p->k) || !mtts::aligned16(p->v) || !mtts::aligned16(p->wq) ||
      (p->wo && !mtts::aligned16(p->wo)))
    return MTTS_ERR_ALIGN;
  if (p->batch == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_BF16:
      if (p->t_kv > 8 * 8 * 4) return MTTS_ERR_UNSUPPORTED;
      return mtts::launch_xattn_block<__nv_bfloat16, 8, 8, 8, 8>(*p, s);
    case MTTS_F32:
      if (p->t_kv > 16 * 8 * 2) return MTTS_ERR_UNSUPPORTED;
      return mtts::launch_xattn_block<float, 16, 16, 8, 2>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_decode_embed(const mtts_decode_embed_params* p, mtts_stream_t stream) {
  if (!p || !p->tok || !p->pos || !p->tok_embed || !p->pos_embed || !p->x) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->dim < 4 || p->dim % 4 != 0 || p->batch > 65535) return MTTS_ERR_SHAPE;
  if (!mtts::aligned16(p->tok_embed) || !mtts::aligned16(p->pos_embed) || !mtts::aligned16(p->x)) return MTTS_ERR_ALIGN;
  if (p->batch == 0) return MTTS_OK;
  mtts::decode_embed_kernel<<<p->batch, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return mtts::launch_status();
}

extern "C" int mtts_decode_greedy(const mtts_decode_greedy_params* p, mtts_stream_t stream) {
  if (!p || !p->logits || !p->tok) return MTTS_ERR_NULL;
  if (p->out && !p->step) return MTTS_ERR_NULL;
  if (p->eos_id >= 0 && !p->lengths) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->vocab < 1 || p->batch > 65535) return MTTS_ERR_SHAPE;
  if (p->batch == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: mtts::decode_greedy_kernel<float><<<p->batch, 256, 0, s>>>(*p); break;
    case MTTS_BF16: mtts::decode_greedy_kernel<__nv_bfloat16><<<p->batch, 256, 0, s>>>(*p); break;
    default: return MTTS_ERR_DTYPE;
  }
  return mtts::launch_status();
}
