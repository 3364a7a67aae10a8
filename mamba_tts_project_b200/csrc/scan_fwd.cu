// selective_scan forward for sm_100a.  Replaces selective_scan_cuda.fwd (mamba_ssm), which the
// reference reaches through Mamba.forward at mamba_decoder.py:61.  Math: see mamba_tts_b200.h.
//
// Design ("time-sequential, state-sliced"):
//   thread = one channel x G consecutive dstate rows; it walks the sequence IN ORDER with the G states
//            in registers, so there is no scan across time at all: per state update exactly one
//            MUFU.EX2 (the binding unit on B200, 16/clk/SM) and 2 packed FFMA2/FMUL2 issue slots;
//   warp   = 32/NG channels x NG state slices of the same channel in adjacent lanes; the only
//            cross-lane traffic is the <C, h> partial sum over the NG slices (a transposing
//            butterfly: 3 SHFL per 4 timesteps at NG = 4);
//   CTA    = kChan channels of one batch element, sequence walked in tiles of TT timesteps:
//            P  every thread turns its 16-byte vectors of u / delta (prefetched into registers one
//               tile ahead) into fp32 dt = softplus(delta + bias) and dt*u rows in shared memory, and
//               transposes the tile of B / C to fp32 [t][n] (16-byte chunk XOR-swizzled);
//            M  the scan proper: LDS.128 of dt, dt*u (per 4 timesteps) and B, C (per timestep, warp
//               broadcast), y partials reduced over the NG lanes and stored to a shared y tile;
//            E  out = (y + D u) * silu(z), packed and streamed out with 16-byte stores.
//   Parallelism comes from channels x state slices (B*Di*NG threads); several small CTAs per SM overlap
//   each other's P/E phases with M.
#include "scan_common.cuh"

namespace mtts {

int dispatch_scan_fwd_wide(const mtts_scan_fwd_params& p, cudaStream_t stream);  // scan_fwd_wide.cu

namespace {

__device__ __forceinline__ float4 lds128(const float* p) { return *reinterpret_cast<const float4*>(p); }

// 16 bytes of a row starting at element t: vector path or bounds-checked scalar gather (zeros >= len)
template <typename T, bool kVec>
__device__ __forceinline__ uint4 load_raw(const T* __restrict__ row, int t, int len) {
  constexpr int VE = Io<T>::kVecElems;
  if constexpr (kVec) {
    if (t < len) return ldg16_stream(row + t);
    return make_uint4(0u, 0u, 0u, 0u);
  } else {
    float v[VE];
#pragma unroll
    for (int i = 0; i < VE; ++i) v[i] = (t + i < len) ? Io<T>::to_f(row[t + i]) : 0.f;
    return Io<T>::pack(v);
  }
}
template <typename T, bool kVec>
__device__ __forceinline__ void store_raw(T* __restrict__ row, int t, int len, const float* v) {
  constexpr int VE = Io<T>::kVecElems;
  if constexpr (kVec) {
    if (t < len) stg16_stream(row + t, Io<T>::pack(v));
  } else {
#pragma unroll
    for (int i = 0; i < VE; ++i)
      if (t + i < len) row[t + i] = Io<T>::from_f(v[i]);
  }
}

}  // namespace

// CC = channels per thread: every LDS of a B / C chunk then feeds CC recurrences (register tiling of
// the shared operands -- the kernel is bound by shared-memory wavefronts).
template <typename T, int G, int NG, int CC, int kChan, int TT, bool kVec>
struct ScanFwdCfg {
  static constexpr int VE = Io<T>::kVecElems;
  static constexpr int kThreads = kChan / CC * NG;
  static constexpr int NP = G * NG;            // padded dstate
  static constexpr int kChunks = NP / 4;       // 16-byte chunks per B/C row
  static constexpr int kSwz = kChunks >= 4 ? 3 : kChunks - 1;
  static constexpr int RS = TT + 4;            // dt / dtu / y row stride (floats)
  static constexpr int kVecPerRow = TT / VE;
  static constexpr int kItems = kChan * kVecPerRow;                   // u/delta/z vectors per tile
  static constexpr int kIt = (kItems + kThreads - 1) / kThreads;      // ... per thread
  static constexpr int kBCItems = NP * kVecPerRow;                    // B (or C) vectors per tile
  static constexpr int kBC = (kBCItems + kThreads - 1) / kThreads;
  static constexpr size_t kSmemFloats = 3 * (size_t)kChan * RS + 2 * (size_t)TT * NP;
  static_assert(G % 4 == 0 && (NG & (NG - 1)) == 0 && NG <= 32 && TT % VE == 0 && TT % 4 == 0, "cfg");
  static_assert(kChan % CC == 0 && kThreads % 32 == 0, "whole warps");
};

template <typename T, int G, int NG, int CC, int kChan, int TT, bool kVec>
__global__ void __launch_bounds__(kChan / CC * NG, (G <= 4 && CC == 1) ? 512 / (kChan * NG) : 1)
scan_fwd_kernel(const mtts_scan_fwd_params p, const int nchunks) {
  using Cfg = ScanFwdCfg<T, G, NG, CC, kChan, TT, kVec>;
  constexpr int VE = Cfg::VE, kThreads = Cfg::kThreads, NP = Cfg::NP, RS = Cfg::RS;
  constexpr int kIt = Cfg::kIt, kBC = Cfg::kBC, kVecPerRow = Cfg::kVecPerRow;
  constexpr int Q = G / 4;

  extern __shared__ __align__(16) float smem[];
  float* dts = smem;                  // [kChan][RS]  dt
  float* dtus = dts + kChan * RS;     // [kChan][RS]  dt * u
  float* ys = dtus + kChan * RS;      // [kChan][RS]  <C, h>
  float* Bs = ys + kChan * RS;        // [TT][NP]     swizzled
  float* Cs = Bs + TT * NP;

  const int N = p.dstate, L = p.seqlen;
  const int b = blockIdx.y, c0 = blockIdx.x * kChan;
  const int tid = threadIdx.x;
  const int chl = tid / NG * CC, g = tid % NG;   // first of this thread's CC channels
  const int c = c0 + chl;

  // ---- per-thread constants: A*log2(e) and the running state of the G rows of this slice ----------
  float2 A2[CC][G / 2], h[CC][G / 2];
#pragma unroll
  for (int k = 0; k < CC; ++k) {
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int n = g * G + i;
      float a = 0.f, hv = 0.f;
      if (c + k < p.dim && n < N) {
        a = p.A[(int64_t)(c + k) * N + n] * kLog2e;
        if (p.initial_state) hv = p.initial_state[((int64_t)b * p.dim + c + k) * N + n];
      }
      reinterpret_cast<float*>(A2[k])[i] = a;
      reinterpret_cast<float*>(h[k])[i] = hv;
    }
  }

  // ---- P/E item bookkeeping: item = (channel row, 16-byte vector) -----------------------------------
  const T* ub = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_batch_stride;
  const T* db = reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_batch_stride;
  const T* zb = p.z ? reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_batch_stride : nullptr;
  T* ob = reinterpret_cast<T*>(p.out) + (int64_t)b * p.out_batch_stride;
  const T* Bb = reinterpret_cast<const T*>(p.B) + (int64_t)b * p.B_batch_stride;
  const T* Cb = reinterpret_cast<const T*>(p.C) + (int64_t)b * p.C_batch_stride;

  int it_ch[kIt], it_t[kIt];
  float it_bias[kIt], it_D[kIt];
  bool it_ok[kIt];
#pragma unroll
  for (int k = 0; k < kIt; ++k) {
    const int idx = tid + k * kThreads;
    it_ch[k] = idx / kVecPerRow;
    it_t[k] = (idx % kVecPerRow) * VE;
    it_ok[k] = idx < Cfg::kItems && c0 + it_ch[k] < p.dim;
    it_bias[k] = (it_ok[k] && p.delta_bias) ? p.delta_bias[c0 + it_ch[k]] : 0.f;
    it_D[k] = (it_ok[k] && p.D) ? p.D[c0 + it_ch[k]] : 0.f;
  }
  int bc_n[kBC], bc_t[kBC];
  bool bc_ok[kBC];
#pragma unroll
  for (int k = 0; k < kBC; ++k) {
    const int idx = tid + k * kThreads;
    bc_n[k] = idx / kVecPerRow;
    bc_t[k] = (idx % kVecPerRow) * VE;
    bc_ok[k] = idx < Cfg::kBCItems;
  }

  uint4 u_nx[kIt], d_nx[kIt], z_cur[kIt], u_cur[kIt], B_nx[kBC], C_nx[kBC];
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  auto prefetch = [&](int t0) {
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      u_nx[k] = d_nx[k] = zero4;
      if (it_ok[k]) {
        const int cc = c0 + it_ch[k];
        u_nx[k] = load_raw<T, kVec>(ub + (int64_t)cc * p.u_dim_stride, t0 + it_t[k], L);
        d_nx[k] = load_raw<T, kVec>(db + (int64_t)cc * p.delta_dim_stride, t0 + it_t[k], L);
      }
    }
#pragma unroll
    for (int k = 0; k < kBC; ++k) {
      B_nx[k] = C_nx[k] = zero4;
      if (bc_ok[k] && bc_n[k] < N) {
        B_nx[k] = load_raw<T, kVec>(Bb + (int64_t)bc_n[k] * p.B_state_stride, t0 + bc_t[k], L);
        C_nx[k] = load_raw<T, kVec>(Cb + (int64_t)bc_n[k] * p.C_state_stride, t0 + bc_t[k], L);
      }
    }
  };

  prefetch(0);
  const int ntiles = (L + TT - 1) / TT;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int t0 = tile * TT;

    // ---- P: registers -> shared fp32 tiles --------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      if (tid + k * kThreads < Cfg::kItems) {
        float uv[VE], dv[VE];
        Io<T>::unpack(u_nx[k], uv);
        Io<T>::unpack(d_nx[k], dv);
        u_cur[k] = u_nx[k];
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          float x = dv[i] + it_bias[k];
          if (p.delta_softplus) x = softplus_f(x);
          if (!it_ok[k] || t0 + it_t[k] + i >= L) x = 0.f;  // identity step: decay 1, input 0
          dv[i] = x;
          uv[i] *= x;
        }
        float* d0 = dts + it_ch[k] * RS + it_t[k];
        float* d1 = dtus + it_ch[k] * RS + it_t[k];
#pragma unroll
        for (int i = 0; i < VE; i += 4) {
          *reinterpret_cast<float4*>(d0 + i) = make_float4(dv[i], dv[i + 1], dv[i + 2], dv[i + 3]);
          *reinterpret_cast<float4*>(d1 + i) = make_float4(uv[i], uv[i + 1], uv[i + 2], uv[i + 3]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kBC; ++k) {
      if (bc_ok[k]) {
        float bv[VE], cv[VE];
        Io<T>::unpack(B_nx[k], bv);
        Io<T>::unpack(C_nx[k], cv);
        const int n = bc_n[k];
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          const int t = bc_t[k] + i;
          const int pos = t * NP + (((n >> 2) ^ ((t >> 3) & Cfg::kSwz)) << 2) + (n & 3);
          Bs[pos] = bv[i];
          Cs[pos] = cv[i];
        }
      }
    }
    __syncthreads();

    // ---- prefetch the next tile (and this tile's z) while M runs -----------------------------------
    if (zb) {
#pragma unroll
      for (int k = 0; k < kIt; ++k) {
        z_cur[k] = zero4;
        if (it_ok[k])
          z_cur[k] = load_raw<T, kVec>(zb + (int64_t)(c0 + it_ch[k]) * p.z_dim_stride, t0 + it_t[k], L);
      }
    }
    if (tile + 1 < ntiles) prefetch(t0 + TT);

    // ---- M: the recurrence -------------------------------------------------------------------------
    // (issuing the exps a group ahead was measured and lost to plain unrolling: the kernel is bound by
    //  shared-memory wavefronts, not by MUFU latency)
    {
      const float* dtr = dts + chl * RS;
      const float* dur = dtus + chl * RS;
      float* yr = ys + chl * RS;
#pragma unroll(G * CC <= 8 ? 2 : 1)
      for (int t4 = 0; t4 < TT; t4 += 4) {
        // checkpoint: state at the start of every MTTS_SCAN_CHUNK timesteps (what the backward restarts from)
        if (p.checkpoints && ((t0 + t4) % MTTS_SCAN_CHUNK) == 0 && t0 + t4 < L) {
#pragma unroll
          for (int k = 0; k < CC; ++k) {
            if (c + k < p.dim) {
              float* ck = p.checkpoints +
                          ((((int64_t)b * p.dim + c + k) * nchunks) + (t0 + t4) / MTTS_SCAN_CHUNK) * N + g * G;
              if ((N & 3) == 0) {
#pragma unroll
                for (int i = 0; i < G; i += 4)
                  if (g * G + i < N)
                    *reinterpret_cast<float4*>(ck + i) =
                        make_float4(h[k][i / 2].x, h[k][i / 2].y, h[k][i / 2 + 1].x, h[k][i / 2 + 1].y);
              } else {
#pragma unroll
                for (int i = 0; i < G; ++i)
                  if (g * G + i < N) ck[i] = reinterpret_cast<const float*>(h[k])[i];
              }
            }
          }
        }
        float dtv[CC][4], duv[CC][4], yp[CC][4];
#pragma unroll
        for (int k = 0; k < CC; ++k) {
          const float4 d4 = lds128(dtr + k * RS + t4);
          const float4 x4 = lds128(dur + k * RS + t4);
          dtv[k][0] = d4.x; dtv[k][1] = d4.y; dtv[k][2] = d4.z; dtv[k][3] = d4.w;
          duv[k][0] = x4.x; duv[k][1] = x4.y; duv[k][2] = x4.z; duv[k][3] = x4.w;
        }
        const int swz = (t4 >> 3) & Cfg::kSwz;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* Bt = Bs + (t4 + j) * NP;
          const float* Ct = Cs + (t4 + j) * NP;
          float2 acc[CC];
#pragma unroll
          for (int k = 0; k < CC; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const int off = ((g * Q + q) ^ swz) << 2;
            const float4 Bv = lds128(Bt + off);
            const float4 Cv = lds128(Ct + off);
#pragma unroll
            for (int k = 0; k < CC; ++k) {
              const float2 dt2 = dup2(dtv[k][j]), du2 = dup2(duv[k][j]);
              const float2 e0 = ex2f2(fmul2(dt2, A2[k][2 * q]));
              const float2 e1 = ex2f2(fmul2(dt2, A2[k][2 * q + 1]));
              h[k][2 * q] = ffma2(e0, h[k][2 * q], fmul2(du2, make_float2(Bv.x, Bv.y)));
              h[k][2 * q + 1] = ffma2(e1, h[k][2 * q + 1], fmul2(du2, make_float2(Bv.z, Bv.w)));
              acc[k] = ffma2(h[k][2 * q], make_float2(Cv.x, Cv.y), acc[k]);
              acc[k] = ffma2(h[k][2 * q + 1], make_float2(Cv.z, Cv.w), acc[k]);
            }
          }
#pragma unroll
          for (int k = 0; k < CC; ++k) yp[k][j] = acc[k].x + acc[k].y;
        }
        // sum the partials over the NG lanes of each channel
#pragma unroll
        for (int k = 0; k < CC; ++k) slice_reduce_store<NG>(yp[k], g, yr + k * RS + t4);
      }
    }
    __syncthreads();

    // ---- E: gate and stream out --------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      if (it_ok[k]) {
        float uv[VE], yv[VE];
        Io<T>::unpack(u_cur[k], uv);
        const float* y0 = ys + it_ch[k] * RS + it_t[k];
#pragma unroll
        for (int i = 0; i < VE; i += 4) {
          const float4 v = lds128(y0 + i);
          yv[i] = v.x; yv[i + 1] = v.y; yv[i + 2] = v.z; yv[i + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < VE; ++i) yv[i] = fmaf(it_D[k], uv[i], yv[i]);
        if (p.y_pre)  // the backward's dz needs y before the gate
          store_raw<T, kVec>(reinterpret_cast<T*>(p.y_pre) + (int64_t)b * p.y_batch_stride +
                                 (int64_t)(c0 + it_ch[k]) * p.y_dim_stride, t0 + it_t[k], L, yv);
        if (zb) {
          float zv[VE];
          Io<T>::unpack(z_cur[k], zv);
#pragma unroll
          for (int i = 0; i < VE; ++i) yv[i] *= silu_f(zv[i]);
        }
        store_raw<T, kVec>(ob + (int64_t)(c0 + it_ch[k]) * p.out_dim_stride, t0 + it_t[k], L, yv);
      }
    }
    // no barrier needed here: the next P writes dts/dtus/Bs/Cs (last read in M, before the barrier
    // above) and the next M writes ys only after the next P's barrier.
  }

  if (p.last_state) {
#pragma unroll
    for (int k = 0; k < CC; ++k) {
      if (c + k < p.dim) {
#pragma unroll
        for (int i = 0; i < G; ++i)
          if (g * G + i < N)
            p.last_state[((int64_t)b * p.dim + c + k) * N + g * G + i] = reinterpret_cast<const float*>(h[k])[i];
      }
    }
  }
}

template <typename T, int G, int NG, int CC, int kChan, int TT, bool kVec>
static int launch_scan_fwd(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  using Cfg = ScanFwdCfg<T, G, NG, CC, kChan, TT, kVec>;
  const int nchunks = (p.seqlen + MTTS_SCAN_CHUNK - 1) / MTTS_SCAN_CHUNK;
  const size_t smem = sizeof(float) * Cfg::kSmemFloats;
  auto kern = scan_fwd_kernel<T, G, NG, CC, kChan, TT, kVec>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  const dim3 grid((p.dim + kChan - 1) / kChan, p.batch);
  kern<<<grid, Cfg::kThreads, smem, stream>>>(p, nchunks);
  return launch_status();
}

// G states per thread x NG slices cover the (padded) dstate; TT scales with the bytes per element so
// that the register-resident prefetch stays at two vectors per thread and stream.
template <typename T, bool kVec>
static int dispatch_scan_fwd_n(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  constexpr int TT = Io<T>::kVecElems * 8;  // 64 (bf16) / 32 (fp32)
  const int N = p.dstate;
  if (N <= 4) return launch_scan_fwd<T, 4, 1, 1, 64, TT / 4, kVec>(p, stream);
  if (N <= 8) return launch_scan_fwd<T, 4, 2, 1, 32, TT / 2, kVec>(p, stream);
  // CC = 2 (<T, 4, 4, 2, 16, TT / 2>) was measured at the C4 shape: 1.97 ms vs 1.87 ms for CC = 1 -- the
  // halved LDS traffic is paid for with half the resident warps, as in the backward.
  if (N <= 16) return launch_scan_fwd<T, 4, 4, 1, 16, TT, kVec>(p, stream);
  // wider states: still 4 rows per thread, more slices per channel (measured at N = 64, C4 shape: 4 x 16
  // slices 7.9 ms, 8 x 8 10.5 ms, 16 x 4 9.1 ms -- registers, hence resident warps, decide)
  if (N <= 32) return launch_scan_fwd<T, 4, 8, 1, 8, TT, kVec>(p, stream);
  if (N <= 64) return launch_scan_fwd<T, 4, 16, 1, 8, TT, kVec>(p, stream);
  if (N <= 128) return launch_scan_fwd<T, 4, 32, 1, 4, TT, kVec>(p, stream);
  return launch_scan_fwd<T, 8, 32, 1, 4, TT, kVec>(p, stream);
}

template <typename T>
static int dispatch_scan_fwd(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  if (scan_use_wide(p.batch, p.dim, p.seqlen)) return dispatch_scan_fwd_wide(p, stream);
  const bool vec = vec_ok<T>(p.u, p.u_batch_stride, p.u_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.delta, p.delta_batch_stride, p.delta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.B, p.B_batch_stride, p.B_state_stride, p.seqlen) &&
                   vec_ok<T>(p.C, p.C_batch_stride, p.C_state_stride, p.seqlen) &&
                   vec_ok<T>(p.z, p.z_batch_stride, p.z_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.out, p.out_batch_stride, p.out_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.y_pre, p.y_batch_stride, p.y_dim_stride, p.seqlen);
  return vec ? dispatch_scan_fwd_n<T, true>(p, stream) : dispatch_scan_fwd_n<T, false>(p, stream);
}

}  // namespace mtts

extern "C" int mtts_selective_scan_fwd(const mtts_scan_fwd_params* p, mtts_stream_t stream) {
  if (!p || !p->u || !p->delta || !p->A || !p->B || !p->C || !p->out) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->dim < 0 || p->seqlen < 0 || p->dstate < 1 ||
      p->dstate > MTTS_MAX_DSTATE || p->batch > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->dim == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p->seqlen == 0) {
    // nothing to scan: the state passes through
    if (p->last_state) {
      const size_t bytes = sizeof(float) * (size_t)p->batch * p->dim * p->dstate;
      cudaError_t e = p->initial_state
                          ? cudaMemcpyAsync(p->last_state, p->initial_state, bytes,
                                            cudaMemcpyDeviceToDevice, s)
                          : cudaMemsetAsync(p->last_state, 0, bytes, s);
      if (e != cudaSuccess) return -static_cast<int>(e);
    }
    return MTTS_OK;
  }
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::dispatch_scan_fwd<float>(*p, s);
    case MTTS_BF16: return mtts::dispatch_scan_fwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}
