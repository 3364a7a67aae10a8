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
//            raw 16-byte vectors of u / delta / z / B / C travel global -> shared with cp.async
//            (LDGSTS), one tile ahead, into slots PRIVATE to the thread that will consume them (no
//            registers are held across M, no barrier is needed for them -- only cp.async.wait_group);
//            P  every thread turns its vectors of u / delta into fp32 dt = softplus(delta + bias) and
//               dt*u rows in shared memory, and transposes the tile of B / C to fp32 [t][n] (16-byte
//               chunk XOR-swizzled);
//            M  the scan proper: LDS.128 of dt, dt*u (per 4 timesteps) and B, C (per timestep, warp
//               broadcast), y partials reduced over the NG lanes and stored over the dt row (all its
//               readers are the lanes of that reduction);
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

// 16 bytes of row[t ..) into *slot: asynchronous when the row is vector-aligned, zeros when absent
template <typename T, bool kVec>
__device__ __forceinline__ void stage_raw(uint4* slot, const T* __restrict__ row, int t, int len, bool ok) {
  if (!ok || t >= len) {
    *slot = make_uint4(0u, 0u, 0u, 0u);
  } else if constexpr (kVec) {
    cp_async16(slot, row + t);
  } else {
    *slot = load_raw<T, false>(row, t, len);
  }
}

}  // namespace

// CC = channels per thread: every LDS of a B / C chunk then feeds CC recurrences (register tiling of
// the shared operands -- shared-memory wavefronts and MUFU are the two busy pipes).
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
  static constexpr int kBCItems = 2 * kChunks * kVecPerRow;           // (tensor, 4-row chunk, vector of timesteps)
  static constexpr int kBC = (kBCItems + kThreads - 1) / kThreads;
  // fp32 rows dt (later y), dt*u; fp32 B, C tiles; raw slots: u (two tiles), delta, z, B, C
  static constexpr size_t kSmemFloats = 2 * (size_t)kChan * RS + 2 * (size_t)TT * NP;
  static constexpr size_t kRawVecs = (size_t)kThreads * (4 * kIt + 4 * kBC);
  static constexpr size_t kSmemBytes = 4 * kSmemFloats + 16 * kRawVecs;
  static_assert(G % 4 == 0 && (NG & (NG - 1)) == 0 && NG <= 32 && TT % VE == 0 && TT % 4 == 0, "cfg");
  static_assert(kChan % CC == 0 && kThreads % 32 == 0, "whole warps");
};

template <typename T, int G, int NG, int CC, int kChan, int TT, int kMinBlocks, bool kVec>
__global__ void __launch_bounds__(kChan / CC * NG, kMinBlocks)
scan_fwd_kernel(const mtts_scan_fwd_params p, const int nchunks) {
  using Cfg = ScanFwdCfg<T, G, NG, CC, kChan, TT, kVec>;
  constexpr int VE = Cfg::VE, kThreads = Cfg::kThreads, NP = Cfg::NP, RS = Cfg::RS;
  constexpr int kIt = Cfg::kIt, kBC = Cfg::kBC, kVecPerRow = Cfg::kVecPerRow;
  constexpr int Q = G / 4;

  extern __shared__ __align__(16) float smem[];
  float* dts = smem;                  // [kChan][RS]  dt, overwritten by <C, h> group by group
  float* dtus = dts + kChan * RS;     // [kChan][RS]  dt * u
  float* ys = dts;
  float* Bs = dtus + kChan * RS;      // [TT][NP]     swizzled
  float* Cs = Bs + TT * NP;
  // raw slots, [item][thread]: a thread only ever touches its own
  uint4* rawU = reinterpret_cast<uint4*>(Cs + TT * NP) + threadIdx.x;  // [2][kIt]
  uint4* rawD = rawU + 2 * kIt * kThreads;                             // [kIt]
  uint4* rawZ = rawD + kIt * kThreads;                                 // [kIt]
  uint4* rawBC = rawZ + kIt * kThreads;                                // [kBC][4 rows]

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

  auto item_ch = [&](int k) { return (tid + k * kThreads) / kVecPerRow; };
  auto item_t = [&](int k) { return ((tid + k * kThreads) % kVecPerRow) * VE; };
  auto item_ok = [&](int k) { return tid + k * kThreads < Cfg::kItems && c0 + item_ch(k) < p.dim; };
  float it_bias[kIt], it_D[kIt];
#pragma unroll
  for (int k = 0; k < kIt; ++k) {
    it_bias[k] = (item_ok(k) && p.delta_bias) ? p.delta_bias[c0 + item_ch(k)] : 0.f;
    it_D[k] = (item_ok(k) && p.D) ? p.D[c0 + item_ch(k)] : 0.f;
  }
  // B / C item = (tensor, chunk of 4 dstate rows, 16-byte vector of timesteps): 4 raw vectors -> VE STS.128
  constexpr int kBCPer = Cfg::kChunks * kVecPerRow;
  auto bc_which = [&](int k) { return (tid + k * kThreads) / kBCPer; };
  auto bc_chunk = [&](int k) { return ((tid + k * kThreads) % kBCPer) / kVecPerRow; };
  auto bc_t = [&](int k) { return ((tid + k * kThreads) % kVecPerRow) * VE; };

  // tile at t0: u -> rawU[ub], delta, B, C (consumed by the next P); zt0 >= 0: z of that tile (for E)
  auto stage = [&](int t0, int ubuf, int zt0) {
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      const bool ok = item_ok(k);
      const int64_t cc = ok ? c0 + item_ch(k) : 0;
      stage_raw<T, kVec>(rawU + (ubuf * kIt + k) * kThreads, ub + cc * p.u_dim_stride, t0 + item_t(k), L,
                         ok && t0 >= 0);
      stage_raw<T, kVec>(rawD + k * kThreads, db + cc * p.delta_dim_stride, t0 + item_t(k), L, ok && t0 >= 0);
      if (zb && zt0 >= 0)
        stage_raw<T, kVec>(rawZ + k * kThreads, zb + cc * p.z_dim_stride, zt0 + item_t(k), L, ok);
    }
    if (t0 >= 0) {
#pragma unroll
      for (int k = 0; k < kBC; ++k) {
        if (tid + k * kThreads < Cfg::kBCItems) {
          const bool isC = bc_which(k) != 0;
          const T* src = isC ? Cb : Bb;
          const int64_t rs = isC ? p.C_state_stride : p.B_state_stride;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = bc_chunk(k) * 4 + i;
            stage_raw<T, kVec>(rawBC + (k * 4 + i) * kThreads, src + (int64_t)(n < N ? n : 0) * rs, t0 + bc_t(k), L,
                               n < N);
          }
        }
      }
    }
    cp_async_commit();
  };

  stage(0, 0, -1);
  cp_async_wait_all();
  const int ntiles = (L + TT - 1) / TT;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int t0 = tile * TT;
    const int ubuf = tile & 1;

    // ---- P: raw slots -> shared fp32 tiles --------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      if (tid + k * kThreads < Cfg::kItems) {
        const int ich = item_ch(k), it = item_t(k);
        // vector path: a vector is wholly inside or outside the sequence (seqlen % VE == 0)
        const bool live = item_ok(k) && (!kVec || t0 + it < L);
        float uv[VE], dv[VE];
        Io<T>::unpack(rawU[(ubuf * kIt + k) * kThreads], uv);
        Io<T>::unpack(rawD[k * kThreads], dv);
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          float x = dv[i] + it_bias[k];
          if (p.delta_softplus) x = softplus_f(x);
          if (!live || (!kVec && t0 + it + i >= L)) x = 0.f;  // identity step: decay 1, input 0
          dv[i] = x;
          uv[i] *= x;
        }
        float* d0 = dts + ich * RS + it;
        float* d1 = dtus + ich * RS + it;
#pragma unroll
        for (int i = 0; i < VE; i += 4) {
          *reinterpret_cast<float4*>(d0 + i) = make_float4(dv[i], dv[i + 1], dv[i + 2], dv[i + 3]);
          *reinterpret_cast<float4*>(d1 + i) = make_float4(uv[i], uv[i + 1], uv[i + 2], uv[i + 3]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kBC; ++k) {
      if (tid + k * kThreads < Cfg::kBCItems) {
        float rows[4][VE];
#pragma unroll
        for (int i = 0; i < 4; ++i) Io<T>::unpack(rawBC[(k * 4 + i) * kThreads], rows[i]);
        float* dst = bc_which(k) ? Cs : Bs;
        const int chunk = bc_chunk(k), tv = bc_t(k);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const int t = tv + e;
          *reinterpret_cast<float4*>(dst + t * NP + ((chunk ^ ((t >> 3) & Cfg::kSwz)) << 2)) =
              make_float4(rows[0][e], rows[1][e], rows[2][e], rows[3][e]);
        }
      }
    }
    __syncthreads();

    // ---- request the next tile (and this tile's z) while M runs ------------------------------------
    stage(tile + 1 < ntiles ? t0 + TT : -1, ubuf ^ 1, t0);

    // ---- M: the recurrence -------------------------------------------------------------------------
    // (issuing the exps a group ahead was measured and lost to plain unrolling: the kernel is bound by
    //  shared-memory wavefronts, not by MUFU latency)
    {
      const float* dtr = dts + chl * RS;
      const float* dur = dtus + chl * RS;
      float* yr = ys + chl * RS;
      // checkpoint: state at the start of every MTTS_SCAN_CHUNK timesteps (what the backward restarts from);
      // tiles are exactly one chunk long
      static_assert(TT == MTTS_SCAN_CHUNK, "one checkpoint per tile");
      if (p.checkpoints) {
#pragma unroll
        for (int k = 0; k < CC; ++k) {
          if (c + k < p.dim) {
            float* ck = p.checkpoints + ((((int64_t)b * p.dim + c + k) * nchunks) + tile) * N + g * G;
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
#pragma unroll(G * CC <= 8 ? 2 : 1)
      for (int t4 = 0; t4 < TT; t4 += 4) {
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
        // sum the partials over the NG lanes of each channel; y replaces dt of the same 4 timesteps, whose
        // only readers were these lanes (same warp, before the shuffles)
#pragma unroll
        for (int k = 0; k < CC; ++k) slice_reduce_store<NG>(yp[k], g, yr + k * RS + t4);
      }
    }
    cp_async_wait_all();  // own slots: z of this tile, raw inputs of the next
    __syncthreads();

    // ---- E: gate and stream out --------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      if (item_ok(k)) {
        const int ich = item_ch(k), it = item_t(k);
        const float Dv = it_D[k];
        float uv[VE], yv[VE];
        Io<T>::unpack(rawU[(ubuf * kIt + k) * kThreads], uv);
        const float* y0 = ys + ich * RS + it;
#pragma unroll
        for (int i = 0; i < VE; i += 4) {
          const float4 v = lds128(y0 + i);
          yv[i] = v.x; yv[i + 1] = v.y; yv[i + 2] = v.z; yv[i + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < VE; ++i) yv[i] = fmaf(Dv, uv[i], yv[i]);
        if (p.y_pre)  // the backward's dz needs y before the gate
          store_raw<T, kVec>(reinterpret_cast<T*>(p.y_pre) + (int64_t)b * p.y_batch_stride +
                                 (int64_t)(c0 + ich) * p.y_dim_stride, t0 + it, L, yv);
        if (zb) {
          float zv[VE];
          Io<T>::unpack(rawZ[k * kThreads], zv);
#pragma unroll
          for (int i = 0; i < VE; ++i) yv[i] *= silu_f(zv[i]);
        }
        store_raw<T, kVec>(ob + (int64_t)(c0 + ich) * p.out_dim_stride, t0 + it, L, yv);
      }
    }
    // no barrier needed here: the next P writes the dt rows of its OWN items (read as y by the same thread
    // in E), dtus/Bs/Cs were last read in M (before the barrier above), raw slots are thread-private.
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

template <typename T, int G, int NG, int CC, int kChan, int TT, int kMinBlocks, bool kVec>
static int launch_scan_fwd(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  using Cfg = ScanFwdCfg<T, G, NG, CC, kChan, TT, kVec>;
  const int nchunks = (p.seqlen + MTTS_SCAN_CHUNK - 1) / MTTS_SCAN_CHUNK;
  const size_t smem = Cfg::kSmemBytes;
  auto kern = scan_fwd_kernel<T, G, NG, CC, kChan, TT, kMinBlocks, kVec>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  const dim3 grid((p.dim + kChan - 1) / kChan, p.batch);
  kern<<<grid, Cfg::kThreads, smem, stream>>>(p, nchunks);
  return launch_status();
}

// G states per thread x NG slices cover the (padded) dstate.  Tiles are 32 timesteps (= the checkpoint
// interval): shared memory per channel -- raw slots plus two fp32 rows -- is what limits the resident warps.
// Measured on B200 (bf16, C4 shape B 32 x d_inner 2048 x T 4096; C2 layer shape B 16 x 1024 x 2048):
//   N = 16: 8 states x 2 slices, 1 channel per thread   1.57 ms / 0.272 ms   <- selected
//           4 x 4, 2 channels per thread                1.61 ms / 0.285 ms
//           4 x 4, 1 channel (the round-1a mapping)     1.92 ms / 0.290 ms
//           8 x 2 or 16 x 1 with 1-warp CTAs            1.97-2.0 ms / 0.35 ms
//   N = 64: 8 x 8, 2 channels, 128 threads              5.69 ms              <- selected
//           8 x 8, 2 channels, 64 threads  5.93;  4 x 16, 2 channels  6.65;  4 x 16, 1 channel  8.2-10.6
// i.e. fewer, fatter threads win as long as >= 12 warps stay resident: every LDS of dt / dt*u feeds 8 states,
// every LDS of a B / C chunk feeds 2 channels (wide states), and the slice butterfly shrinks.
template <typename T, bool kVec>
static int dispatch_scan_fwd_n(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  constexpr int TT = 32;
  const int N = p.dstate;
  if (N <= 4) return launch_scan_fwd<T, 4, 1, 1, 64, TT, 4, kVec>(p, stream);
  if (N <= 8) return launch_scan_fwd<T, 4, 2, 1, 32, TT, 4, kVec>(p, stream);
  if (N <= 16) return launch_scan_fwd<T, 8, 2, 1, 32, TT, 8, kVec>(p, stream);
  if (N <= 32) return launch_scan_fwd<T, 8, 4, 2, 32, TT, 4, kVec>(p, stream);
  if (N <= 64) return launch_scan_fwd<T, 8, 8, 2, 32, TT, 3, kVec>(p, stream);
  if (N <= 128) return launch_scan_fwd<T, 8, 16, 2, 16, TT, 3, kVec>(p, stream);
  return launch_scan_fwd<T, 8, 32, 1, 4, TT, 1, kVec>(p, stream);
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
