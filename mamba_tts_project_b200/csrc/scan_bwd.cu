// selective_scan backward for sm_100a.  Replaces selective_scan_cuda.bwd (mamba_ssm), reached from
// loss.backward() at train.py:231 through MambaInnerFn.backward.
//
// Same decomposition as the forward (scan_fwd.cu): thread = one channel x 4 consecutive dstate rows,
// walking time sequentially with the states in registers; a warp = 32/NG channels x NG state slices.
// Recompute-based: the forward saved the state at the start of every MTTS_SCAN_CHUNK (= 32 = one tile)
// timesteps.  Tiles are walked last to first; per tile
//   P   dt = softplus(delta + bias), dt*u, gy = dout * silu(z) -> fp32 shared tiles; B / C tile -> fp32
//       [t][n] (16-byte chunk XOR-swizzled);
//   M   pre-pass: the recurrence over the first 28 timesteps from the checkpoint, leaving the state at the
//       start of each group of 4 timesteps in shared memory (16 bytes per thread and group);
//       then the groups last to first: re-run the 4 steps keeping decays a_t and states h_t in registers
//       (y_t = <C_t, h_t> for dz falls out), then the reverse recurrence
//           G_t = C_t gy_t + a_{t+1} G_{t+1}
//           dC_t = gy_t h_t              dB_t = (dt_t u_t) G_t                   (summed over channels)
//           sGB_t = <G_t, B_t>           w_t = G_t (h_t - dt_t u_t B_t) = G_t a_t h_{t-1}
//           ddtA_t = <w_t, A>            dA += w_t dt_t
//       y / sGB / ddtA are summed over the NG slice lanes with a transposing butterfly (3 SHFL per 4
//       values at NG = 4); dB / dC over the warp's channel lanes with a halving butterfly (14 SHFL per
//       16 values at 8 channel lanes) whose survivors -- two consecutive timesteps of one state row per
//       lane -- go straight to global memory with 8-byte vector REDs (no shared tile, no flush loop);
//   E   du = gy D + dt sGB,  ddelta = (ddtA + u sGB) softplus'(.),  dz = dout y silu'(z), streamed out; the
//       raw u / delta / dout / z vectors come back from thread-private shared slots filled in P.
// Two MUFU.EX2 per state update (pre-pass + re-run), everything else packed fp32x2.
#include "scan_common.cuh"

// Timing experiments only (results are wrong when set): bit 0 skips the dB / dC channel reduction + REDs, bit 1 the
// slice reductions, bit 2 the E phase, bit 3 the pre-pass, bit 4 the P phase math.
#ifndef MTTS_SCAN_DBG_SKIP
#define MTTS_SCAN_DBG_SKIP 0
#endif

namespace mtts {

int dispatch_scan_bwd_wide(const mtts_scan_bwd_params& p, cudaStream_t stream);  // scan_bwd_wide.cu

namespace {

constexpr int kTT = MTTS_SCAN_CHUNK;  // tile = checkpoint interval
static_assert(kTT == 32, "tile bookkeeping below assumes 32-timestep tiles");

__device__ __forceinline__ float4 lds128(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }

template <typename T, bool kVec>
__device__ __forceinline__ uint4 load_raw(const T* __restrict__ row, int t, int len) {
  constexpr int VE = Io<T>::kVecElems;
  if constexpr (kVec) {
    if (t < len) return ldg16(row + t);
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

// Halving butterfly over the channel lanes (lane bits [log2 NG, 5)): v[i*4 + j] (state i of the slice,
// timestep j) summed over the warp's 32/NG channels.
template <int NG, int CNT, int BIT>
__device__ __forceinline__ void chan_reduce_step(float (&v)[16], int lane, int& prefix) {
  if constexpr (BIT >= NG) {
    if constexpr (CNT > 1) {
      constexpr int H = CNT / 2;
      const bool up = lane & BIT;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float send = up ? v[i] : v[i + H];
        const float keep = up ? v[i + H] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
      }
      prefix = prefix * 2 + (up ? 1 : 0);
      chan_reduce_step<NG, H, BIT / 2>(v, lane, prefix);
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], BIT);
      chan_reduce_step<NG, 1, BIT / 2>(v, lane, prefix);
    }
  }
}
// After the butterfly a writer lane holds R = 16 >> stages values v[0..R): idx = prefix * R + r with state
// i = idx >> 2 of the slice and timestep j = idx & 3 of the group; prefix depends on the lane only.
template <int NG>
struct ChanReduce {
  static constexpr int CL = 32 / NG;                                           // channel lanes
  static constexpr int S = CL >= 16 ? 4 : (CL == 8 ? 3 : (CL == 4 ? 2 : 1));   // halving stages
  static constexpr int R = 16 >> S;                                            // values left per lane
  __device__ static __forceinline__ bool writer(int lane) { return CL <= 16 || (lane & NG) == 0; }
  __device__ static __forceinline__ int prefix(int lane) {
    int pf = 0;
#pragma unroll
    for (int bit = 16, cnt = 16; bit >= NG && cnt > 1; bit >>= 1, cnt >>= 1) pf = pf * 2 + ((lane & bit) ? 1 : 0);
    return pf;
  }
};
// Sum v over the channel lanes and add the survivors to row[0..R) (R consecutive timesteps of one state row
// of dB / dC in global memory); nvalid = how many of them lie inside the sequence.
template <int NG, bool kVec>
__device__ __forceinline__ void chan_reduce_red(float (&v)[16], int lane, float* row, int nvalid) {
  using CR = ChanReduce<NG>;
  int prefix = 0;
  chan_reduce_step<NG, 16, 16>(v, lane, prefix);
  if (CR::writer(lane) && row != nullptr) {
    if (CR::R == 2 && kVec) {
      if (nvalid > 0)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(row), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
#pragma unroll
      for (int r = 0; r < CR::R; ++r)
        if (r < nvalid) atomicAdd(row + r, v[r]);
    }
  }
}

}  // namespace

// CC = channels per thread (register tiling of the B / C operands: every LDS of a B / C chunk feeds CC
// recurrences, and the dB / dC contributions of the CC channels are summed in-thread before the
// butterfly), kWarps = warps per CTA along the channels, NS = warps along the state dimension: NS chunks of
// 4 * NG states share one CTA's P / E phases and its dt / dt*u / gy rows (wide states).
template <typename T, int NG, int CC, int kWarps, bool kVec, int NS = 1>
struct ScanBwdCfg {
  static constexpr int VE = Io<T>::kVecElems;
  static constexpr int kCThreads = 32 * kWarps;  // threads of one state chunk
  static constexpr int kThreads = kCThreads * NS;
  static constexpr int kChan = kCThreads / NG * CC;
  static constexpr int NP = 4 * NG * NS;
  static constexpr int kSwz = NG >= 4 ? 3 : NG - 1;
  static constexpr int RS = kTT + 4;
  static constexpr int kVecPerRow = kTT / VE;
  static constexpr int kItems = kChan * kVecPerRow;
  static constexpr int kIt = (kItems + kThreads - 1) / kThreads;
  static constexpr int kBCItems = NG * NS * kVecPerRow;  // (4-row chunk, 16-byte vector) per tensor
  // dt, dtu (-> sGB when NS == 1), gy rows; per state chunk: ddtA (, y) (, sGB when NS > 1) rows; B, C tiles;
  // group-start states; raw u / delta / dout / z slots + two tiles of the forward's y
  static constexpr size_t smem_floats(bool recompute_y) {
    return (3 + NS * ((recompute_y ? 2 : 1) + (NS > 1 ? 1 : 0))) * (size_t)kChan * RS + 2 * (size_t)kTT * NP +
           7 * 4 * CC * (size_t)kThreads + 6 * 4 * kIt * (size_t)kThreads;
  }
};

// kRecomputeY: dz needs y = <C, h> and the forward did not hand over its y_pre.
template <typename T, int NG, int CC, int kWarps, bool kVec, bool kRecomputeY, int NS>
__global__ void __launch_bounds__(32 * kWarps * NS, (NG == 4 && CC == 1) ? 7 : 1)
scan_bwd_kernel(const mtts_scan_bwd_params p, const int nchunks) {
  using Cfg = ScanBwdCfg<T, NG, CC, kWarps, kVec, NS>;
  constexpr int VE = Cfg::VE, kChan = Cfg::kChan, NP = Cfg::NP, RS = Cfg::RS;
  constexpr int kIt = Cfg::kIt, kVecPerRow = Cfg::kVecPerRow, kThreads = Cfg::kThreads;

  extern __shared__ __align__(16) float smem[];
  float* dts = smem;                      // [kChan][RS] dt
  float* dtus = dts + kChan * RS;         // dt*u, overwritten by sGB
  float* gys = dtus + kChan * RS;         // dout * silu(z)
  float* das = gys + kChan * RS;          // [NS][kChan][RS] <w, A*log2e> over the chunk's states
  float* ys = das + NS * kChan * RS;      // [NS] <C, h> (only when recomputed here)
  float* sgs = ys + (kRecomputeY ? NS * kChan * RS : 0);  // [NS] <G, B>; a single chunk reuses the dt*u rows
  float* Bs = sgs + (NS > 1 ? NS * kChan * RS : 0);      // [kTT][NP] swizzled
  float* Cs = Bs + kTT * NP;
  float* hbs = Cs + kTT * NP;             // [groups 1..7][kThreads][CC][4] state at the start of the group
  // raw u / delta / dout / z vectors of the tile, [tensor][item][thread]: written in P, read back in E by
  // the same thread (instead of a second trip to L2 with its address arithmetic)
  uint4* raws = reinterpret_cast<uint4*>(hbs + 7 * 4 * CC * kThreads) + threadIdx.x;

  const int N = p.dstate, L = p.seqlen;
  const int b = blockIdx.y, c0 = blockIdx.x * kChan;
  const int tid = threadIdx.x, lane = tid & 31;
  const int sc = tid / Cfg::kCThreads, tic = tid % Cfg::kCThreads;  // state chunk, thread within it
  const int chl = tic / NG * CC, g = tic % NG;  // first of this thread's CC channels
  const int n0 = (sc * NG + g) * 4;              // first of this thread's 4 states
  const int c = c0 + chl;

  float2 A2[CC][2], Gc[CC][2], dAacc[CC][2];
#pragma unroll
  for (int k = 0; k < CC; ++k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + i;
      reinterpret_cast<float*>(A2[k])[i] = (c + k < p.dim && n < N) ? p.A[(int64_t)(c + k) * N + n] * kLog2e : 0.f;
      reinterpret_cast<float*>(Gc[k])[i] = 0.f;
      reinterpret_cast<float*>(dAacc[k])[i] = 0.f;
    }
  }

  // dB / dC: the state row and the first timestep (within a group of 4) this lane's butterfly survivors
  // belong to
  using CR = ChanReduce<NG>;
  const int red_idx = CR::prefix(lane) * CR::R;
  const int red_n = n0 + (red_idx >> 2), red_j = red_idx & 3;
  const int64_t red_off = ((int64_t)b * N + red_n) * L + red_j;
  const bool red_ok = red_n < N;

  const T* ub = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_batch_stride;
  const T* db = reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_batch_stride;
  const T* gob = reinterpret_cast<const T*>(p.dout) + (int64_t)b * p.dout_batch_stride;
  const T* zb = p.z ? reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_batch_stride : nullptr;
  const T* Bb = reinterpret_cast<const T*>(p.B) + (int64_t)b * p.B_batch_stride;
  const T* Cb = reinterpret_cast<const T*>(p.C) + (int64_t)b * p.C_batch_stride;

  float dD_acc[kIt], dbias_acc[kIt];
#pragma unroll
  for (int k = 0; k < kIt; ++k) dD_acc[k] = dbias_acc[k] = 0.f;

  // Raw inputs of a tile (u, delta, dout, z vectors of this thread's items) and its checkpoints are fetched
  // one tile ahead, during the previous tile's E phase, so they are never live across M.
  uint4 ur[kIt], dr[kIt], gr[kIt], zr[kIt];
  float4 ck4[CC];
  // the forward's y of a tile (for dz) goes global -> shared by cp.async into a slot of its own parity: it is
  // requested a whole tile before E reads it and never occupies a register
  const T* yb = (!kRecomputeY && zb && p.y_pre) ? reinterpret_cast<const T*>(p.y_pre) + (int64_t)b * p.y_batch_stride
                                                 : nullptr;
  auto fetch_tile = [&](int tile) {
    const int t0 = tile * kTT;
    if (yb) {
#pragma unroll
      for (int k = 0; k < kIt; ++k) {
        const int idx = tid + k * kThreads;
        const int ich = idx / kVecPerRow, it = (idx % kVecPerRow) * VE;
        const int cc = c0 + ich;
        uint4* slot = raws + ((4 + (tile & 1)) * kIt + k) * kThreads;
        if (idx < Cfg::kItems && cc < p.dim && t0 + it < L) {
          if constexpr (kVec) cp_async16(slot, yb + (int64_t)cc * p.y_dim_stride + t0 + it);
          else *slot = load_raw<T, false>(yb + (int64_t)cc * p.y_dim_stride, t0 + it, L);
        } else {
          *slot = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      cp_async_commit();
    }
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      const int idx = tid + k * kThreads;
      const int ich = idx / kVecPerRow, it = (idx % kVecPerRow) * VE;
      const int cc = c0 + ich;
      ur[k] = dr[k] = gr[k] = zr[k] = make_uint4(0u, 0u, 0u, 0u);
      if (idx < Cfg::kItems && cc < p.dim) {
        ur[k] = load_raw<T, kVec>(ub + (int64_t)cc * p.u_dim_stride, t0 + it, L);
        dr[k] = load_raw<T, kVec>(db + (int64_t)cc * p.delta_dim_stride, t0 + it, L);
        gr[k] = load_raw<T, kVec>(gob + (int64_t)cc * p.dout_dim_stride, t0 + it, L);
        if (zb) zr[k] = load_raw<T, kVec>(zb + (int64_t)cc * p.z_dim_stride, t0 + it, L);
      }
    }
#pragma unroll
    for (int k = 0; k < CC; ++k) {
      float hv[4] = {0.f, 0.f, 0.f, 0.f};
      if (c + k < p.dim) {
        const float* ck = p.checkpoints + (((int64_t)b * p.dim + c + k) * nchunks + tile) * N + n0;
        if ((N & 3) == 0) {
          if (n0 < N) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(ck));
            hv[0] = v.x; hv[1] = v.y; hv[2] = v.z; hv[3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (n0 + i < N) hv[i] = ck[i];
        }
      }
      ck4[k] = make_float4(hv[0], hv[1], hv[2], hv[3]);
    }
  };

  const int ntiles = (L + kTT - 1) / kTT;
  fetch_tile(ntiles - 1);
  for (int tile = ntiles - 1; tile >= 0; --tile) {
    const int t0 = tile * kTT;

    // ---- P ----------------------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < kIt; ++k) {
      const int idx = tid + k * kThreads;
      if (idx < Cfg::kItems) {
        const int ich = idx / kVecPerRow, it = (idx % kVecPerRow) * VE;
        const int cc = c0 + ich;
        const bool ok = cc < p.dim;
        float uv[VE], dv[VE], gv[VE];
        raws[(0 * kIt + k) * kThreads] = ur[k];
        raws[(1 * kIt + k) * kThreads] = dr[k];
        raws[(2 * kIt + k) * kThreads] = gr[k];
        if (zb) raws[(3 * kIt + k) * kThreads] = zr[k];
        {
          Io<T>::unpack(ur[k], uv);
          Io<T>::unpack(dr[k], dv);
          Io<T>::unpack(gr[k], gv);
          if (zb) {
            float zv[VE];
            Io<T>::unpack(zr[k], zv);
#pragma unroll
            for (int i = 0; i < VE; ++i) gv[i] *= silu_f(zv[i]);
          }
        }
        const float bias = (ok && p.delta_bias) ? p.delta_bias[cc] : 0.f;
        // vector path: a vector is wholly inside or outside the sequence (seqlen % VE == 0)
        const bool live = ok && (!kVec || t0 + it < L);
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          float x = dv[i] + bias;
          if (p.delta_softplus) x = softplus_f(x);
          if (!live || (!kVec && t0 + it + i >= L)) x = 0.f;  // identity step (gy is 0 there: dout loads as 0)
          dv[i] = x;
          uv[i] *= x;
        }
        float* r0 = dts + ich * RS + it;
        float* r1 = dtus + ich * RS + it;
        float* r2 = gys + ich * RS + it;
#pragma unroll
        for (int i = 0; i < VE; i += 4) {
          *reinterpret_cast<float4*>(r0 + i) = make_float4(dv[i], dv[i + 1], dv[i + 2], dv[i + 3]);
          *reinterpret_cast<float4*>(r1 + i) = make_float4(uv[i], uv[i + 1], uv[i + 2], uv[i + 3]);
          *reinterpret_cast<float4*>(r2 + i) = make_float4(gv[i], gv[i + 1], gv[i + 2], gv[i + 3]);
        }
      }
    }
    // B / C: item = (tensor, chunk of 4 dstate rows, 16-byte vector of timesteps) -> VE STS.128
    for (int idx = tid; idx < 2 * Cfg::kBCItems; idx += kThreads) {
      const int which = idx / Cfg::kBCItems, r = idx % Cfg::kBCItems;
      const int chunk = r / kVecPerRow, tv = (r % kVecPerRow) * VE;
      const T* src = which ? Cb : Bb;
      const int64_t rs = which ? p.C_state_stride : p.B_state_stride;
      float* dst = which ? Cs : Bs;
      float rows[4][VE];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = chunk * 4 + i;
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);
        if (n < N) raw = load_raw<T, kVec>(src + (int64_t)n * rs, t0 + tv, L);
        Io<T>::unpack(raw, rows[i]);
      }
#pragma unroll
      for (int e = 0; e < VE; ++e) {
        const int t = tv + e;
        *reinterpret_cast<float4*>(dst + t * NP + ((chunk ^ ((t >> 3) & Cfg::kSwz)) << 2)) =
            make_float4(rows[0][e], rows[1][e], rows[2][e], rows[3][e]);
      }
    }
    __syncthreads();

    // ---- M ----------------------------------------------------------------------------------------
    {
      const float* dtr = dts + chl * RS;
      const float* dur = dtus + chl * RS;
      const float* gyr = gys + chl * RS;
      float* yr = ys + (sc * kChan + chl) * RS;
      float* dar = das + (sc * kChan + chl) * RS;
      float* sgr = NS > 1 ? sgs + (sc * kChan + chl) * RS : dtus + chl * RS;
      float* hb = hbs + tid * (4 * CC);

      float2 h[CC][2];
      float4 h0[CC];
#pragma unroll
      for (int k = 0; k < CC; ++k) {
        h[k][0] = make_float2(ck4[k].x, ck4[k].y);
        h[k][1] = make_float2(ck4[k].z, ck4[k].w);
        h0[k] = ck4[k];
      }
      // pre-pass: state at the start of every group of 4 timesteps
#pragma unroll 1
      for (int s = 0; s < ((MTTS_SCAN_DBG_SKIP & 8) ? 0 : 7); ++s) {
        float dtv[CC][4], duv[CC][4];
#pragma unroll
        for (int k = 0; k < CC; ++k) {
          const float4 d4 = lds128(dtr + k * RS + 4 * s);
          const float4 x4 = lds128(dur + k * RS + 4 * s);
          dtv[k][0] = d4.x; dtv[k][1] = d4.y; dtv[k][2] = d4.z; dtv[k][3] = d4.w;
          duv[k][0] = x4.x; duv[k][1] = x4.y; duv[k][2] = x4.z; duv[k][3] = x4.w;
        }
        const int off = ((sc * NG + g) ^ ((s >> 1) & Cfg::kSwz)) << 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 Bv = lds128(Bs + (4 * s + j) * NP + off);
#pragma unroll
          for (int k = 0; k < CC; ++k) {
            const float2 dt2 = dup2(dtv[k][j]), du2 = dup2(duv[k][j]);
            h[k][0] = ffma2(ex2f2(fmul2(dt2, A2[k][0])), h[k][0], fmul2(du2, lo2(Bv)));
            h[k][1] = ffma2(ex2f2(fmul2(dt2, A2[k][1])), h[k][1], fmul2(du2, hi2(Bv)));
          }
        }
#pragma unroll
        for (int k = 0; k < CC; ++k)
          *reinterpret_cast<float4*>(hb + s * 4 * CC * kThreads + 4 * k) =
              make_float4(h[k][0].x, h[k][0].y, h[k][1].x, h[k][1].y);
      }

#pragma unroll 1
      for (int s = 7; s >= 0; --s) {
        float dtv[CC][4], duv[CC][4], gyv[CC][4];
#pragma unroll
        for (int k = 0; k < CC; ++k) {
          const float4 d4 = lds128(dtr + k * RS + 4 * s);
          const float4 x4 = lds128(dur + k * RS + 4 * s);
          const float4 g4 = lds128(gyr + k * RS + 4 * s);
          dtv[k][0] = d4.x; dtv[k][1] = d4.y; dtv[k][2] = d4.z; dtv[k][3] = d4.w;
          duv[k][0] = x4.x; duv[k][1] = x4.y; duv[k][2] = x4.z; duv[k][3] = x4.w;
          gyv[k][0] = g4.x; gyv[k][1] = g4.y; gyv[k][2] = g4.z; gyv[k][3] = g4.w;
        }
        const int off = ((sc * NG + g) ^ ((s >> 1) & Cfg::kSwz)) << 2;
        const float* Bt = Bs + 4 * s * NP + off;
        const float* Ct = Cs + 4 * s * NP + off;

        // re-run the 4 steps, keeping decays and states
        float2 a[CC][4][2], hs[CC][4][2];
        float4 Bkeep[4];  // B_t of the group: the reverse sweep needs it again
        {
          float2 hc[CC][2];
          float yp[CC][4];
#pragma unroll
          for (int k = 0; k < CC; ++k) {
            const float4 h4 = s == 0 ? h0[k] : lds128(hb + (s - 1) * 4 * CC * kThreads + 4 * k);
            hc[k][0] = lo2(h4);
            hc[k][1] = hi2(h4);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 Bv = lds128(Bt + j * NP);
            Bkeep[j] = Bv;
            float4 Cv = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (kRecomputeY) Cv = lds128(Ct + j * NP);
#pragma unroll
            for (int k = 0; k < CC; ++k) {
              const float2 dt2 = dup2(dtv[k][j]), du2 = dup2(duv[k][j]);
              a[k][j][0] = ex2f2(fmul2(dt2, A2[k][0]));
              a[k][j][1] = ex2f2(fmul2(dt2, A2[k][1]));
              hc[k][0] = ffma2(a[k][j][0], hc[k][0], fmul2(du2, lo2(Bv)));
              hc[k][1] = ffma2(a[k][j][1], hc[k][1], fmul2(du2, hi2(Bv)));
              hs[k][j][0] = hc[k][0];
              hs[k][j][1] = hc[k][1];
              if constexpr (kRecomputeY) {
                const float2 acc = ffma2(hc[k][1], hi2(Cv), fmul2(hc[k][0], lo2(Cv)));
                yp[k][j] = acc.x + acc.y;
              }
            }
          }
          if constexpr (kRecomputeY) {
#pragma unroll
            for (int k = 0; k < CC; ++k) slice_reduce_store<NG>(yp[k], g, yr + k * RS + 4 * s);
          }
        }

        // reverse recurrence
        float dBv[16], dCv[16], sgb[CC][4], dda[CC][4];
#pragma unroll
        for (int j = 3; j >= 0; --j) {
          const float4 Bv = Bkeep[j];
          const float4 Cv = lds128(Ct + j * NP);
          float2 dB0, dB1, dC0, dC1;
#pragma unroll
          for (int k = 0; k < CC; ++k) {
            const float2 gy2 = dup2(gyv[k][j]), du2 = dup2(duv[k][j]), dt2 = dup2(dtv[k][j]);
            const float2 G0 = ffma2(lo2(Cv), gy2, Gc[k][0]);
            const float2 G1 = ffma2(hi2(Cv), gy2, Gc[k][1]);
            if (k == 0) {
              dC0 = fmul2(gy2, hs[k][j][0]); dC1 = fmul2(gy2, hs[k][j][1]);
              dB0 = fmul2(du2, G0); dB1 = fmul2(du2, G1);
            } else {
              dC0 = ffma2(gy2, hs[k][j][0], dC0); dC1 = ffma2(gy2, hs[k][j][1], dC1);
              dB0 = ffma2(du2, G0, dB0); dB1 = ffma2(du2, G1, dB1);
            }
            const float2 sg = ffma2(G1, hi2(Bv), fmul2(G0, lo2(Bv)));
            sgb[k][j] = sg.x + sg.y;
            // a_t h_{t-1} = h_t - dt u B_t
            const float2 ndu = make_float2(-du2.x, -du2.y);
            const float2 ah0 = ffma2(ndu, lo2(Bv), hs[k][j][0]);
            const float2 ah1 = ffma2(ndu, hi2(Bv), hs[k][j][1]);
            const float2 w0 = fmul2(G0, ah0), w1 = fmul2(G1, ah1);
            const float2 da = ffma2(w1, A2[k][1], fmul2(w0, A2[k][0]));
            dda[k][j] = da.x + da.y;
            dAacc[k][0] = ffma2(w0, dt2, dAacc[k][0]);
            dAacc[k][1] = ffma2(w1, dt2, dAacc[k][1]);
            Gc[k][0] = fmul2(a[k][j][0], G0);
            Gc[k][1] = fmul2(a[k][j][1], G1);
          }
          // state-major (idx = i*4 + j): the first butterfly stages split on the state bits, so they can
          // start as soon as this timestep is done instead of after the whole group
          dCv[j] = dC0.x; dCv[4 + j] = dC0.y; dCv[8 + j] = dC1.x; dCv[12 + j] = dC1.y;
          dBv[j] = dB0.x; dBv[4 + j] = dB0.y; dBv[8 + j] = dB1.x; dBv[12 + j] = dB1.y;
        }
        if constexpr (!(MTTS_SCAN_DBG_SKIP & 2)) {
#pragma unroll
          for (int k = 0; k < CC; ++k) {
            slice_reduce_store<NG>(sgb[k], g, sgr + k * RS + 4 * s);
            slice_reduce_store<NG>(dda[k], g, dar + k * RS + 4 * s);
          }
        } else {
          if (sgb[0][0] + dda[0][0] + sgb[CC - 1][3] + dda[CC - 1][3] == 123.f) sgr[0] = 0.f;
        }
        if constexpr (MTTS_SCAN_DBG_SKIP & 1) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += dBv[i] + dCv[i];
          if (acc == 123.f) p.dB[0] = acc;
        } else {
          const int tg = t0 + 4 * s + red_j;
          const int nvalid = red_ok ? L - tg : 0;
          chan_reduce_red<NG, kVec>(dBv, lane, p.dB + red_off + t0 + 4 * s, nvalid);
          chan_reduce_red<NG, kVec>(dCv, lane, p.dC + red_off + t0 + 4 * s, nvalid);
        }
      }
    }
    __syncthreads();

    // ---- E ----------------------------------------------------------------------------------------
    // the next tile's raw inputs / checkpoints / y are requested now (this tile's y sits in the other slot)
    cp_async_wait_all();
    if (tile > 0) fetch_tile(tile - 1);
#pragma unroll
    for (int k = 0; k < ((MTTS_SCAN_DBG_SKIP & 4) ? 0 : kIt); ++k) {
      const int idx = tid + k * kThreads;
      const int ich = idx / kVecPerRow, it = (idx % kVecPerRow) * VE;
      const int cc = c0 + ich;
      if (idx < Cfg::kItems && cc < p.dim) {
        float uv[VE], dv[VE], gv[VE], zv[VE];
        Io<T>::unpack(raws[(0 * kIt + k) * kThreads], uv);
        Io<T>::unpack(raws[(1 * kIt + k) * kThreads], dv);
        Io<T>::unpack(raws[(2 * kIt + k) * kThreads], gv);
        if (zb) Io<T>::unpack(raws[(3 * kIt + k) * kThreads], zv);
        const float bias = p.delta_bias ? p.delta_bias[cc] : 0.f;
        const float Dv = p.D ? p.D[cc] : 0.f;
        float dtv[VE], sgv[VE], yv[VE], dav[VE];
#pragma unroll
        for (int i = 0; i < VE; i += 4) {
          const float4 q0 = lds128(dts + ich * RS + it + i);
          float4 q1 = lds128((NS > 1 ? sgs : dtus) + ich * RS + it + i);
          float4 q2 = kRecomputeY ? lds128(ys + ich * RS + it + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 q3 = lds128(das + ich * RS + it + i);
#pragma unroll
          for (int w = 1; w < NS; ++w) {  // the chunks' partial sums over their states
            const float4 a1 = lds128(sgs + (w * kChan + ich) * RS + it + i);
            const float4 a3 = lds128(das + (w * kChan + ich) * RS + it + i);
            q1.x += a1.x; q1.y += a1.y; q1.z += a1.z; q1.w += a1.w;
            q3.x += a3.x; q3.y += a3.y; q3.z += a3.z; q3.w += a3.w;
            if constexpr (kRecomputeY) {
              const float4 a2 = lds128(ys + (w * kChan + ich) * RS + it + i);
              q2.x += a2.x; q2.y += a2.y; q2.z += a2.z; q2.w += a2.w;
            }
          }
          dtv[i] = q0.x; dtv[i + 1] = q0.y; dtv[i + 2] = q0.z; dtv[i + 3] = q0.w;
          sgv[i] = q1.x; sgv[i + 1] = q1.y; sgv[i + 2] = q1.z; sgv[i + 3] = q1.w;
          if constexpr (kRecomputeY) {
            yv[i] = q2.x; yv[i + 1] = q2.y; yv[i + 2] = q2.z; yv[i + 3] = q2.w;
          }
          dav[i] = q3.x; dav[i + 1] = q3.y; dav[i + 2] = q3.z; dav[i + 3] = q3.w;
        }
        if constexpr (!kRecomputeY) {  // the forward's y already holds D u
          if (yb) Io<T>::unpack(raws[((4 + (tile & 1)) * kIt + k) * kThreads], yv);
          else
#pragma unroll
            for (int i = 0; i < VE; ++i) yv[i] = 0.f;
        }
        float o_du[VE], o_dd[VE], o_dz[VE];
#pragma unroll
        for (int i = 0; i < VE; ++i) {
          float gy = gv[i];
          if (zb) {
            const float sig = sigmoid_f(zv[i]);
            gy *= zv[i] * sig;
            const float yfull = kRecomputeY ? fmaf(Dv, uv[i], yv[i]) : yv[i];
            o_dz[i] = gv[i] * yfull * sig * fmaf(zv[i], 1.f - sig, 1.f);
          }
          const float x = dv[i] + bias;
          const float sp = (p.delta_softplus && x <= 20.f) ? sigmoid_f(x) : 1.f;
          o_du[i] = fmaf(dtv[i], sgv[i], gy * Dv);
          const float dd = (kVec ? t0 + it < L : t0 + it + i < L) ? fmaf(uv[i], sgv[i], dav[i] * kLn2) * sp : 0.f;
          o_dd[i] = dd;
          dbias_acc[k] += dd;
          dD_acc[k] = fmaf(gy, uv[i], dD_acc[k]);
        }
        store_raw<T, kVec>(reinterpret_cast<T*>(p.du) + (int64_t)b * p.du_batch_stride +
                               (int64_t)cc * p.du_dim_stride, t0 + it, L, o_du);
        store_raw<T, kVec>(reinterpret_cast<T*>(p.ddelta) + (int64_t)b * p.ddelta_batch_stride +
                               (int64_t)cc * p.ddelta_dim_stride, t0 + it, L, o_dd);
        if (zb)
          store_raw<T, kVec>(reinterpret_cast<T*>(p.dz) + (int64_t)b * p.dz_batch_stride +
                                 (int64_t)cc * p.dz_dim_stride, t0 + it, L, o_dz);
      }
    }
    __syncthreads();  // tiles are single-buffered
  }

  // ---- per-channel reductions over time ------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < kIt; ++k) {
    float sD = dD_acc[k], sb = dbias_acc[k];
#pragma unroll
    for (int o = kVecPerRow / 2; o > 0; o >>= 1) {
      sD += __shfl_xor_sync(0xffffffffu, sD, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    const int idx = tid + k * kThreads;
    const int cc = c0 + idx / kVecPerRow;
    if (idx < Cfg::kItems && cc < p.dim && (idx % kVecPerRow) == 0) {
      if (p.dD) atomicAdd(p.dD + cc, sD);
      if (p.ddelta_bias) atomicAdd(p.ddelta_bias + cc, sb);
    }
  }
#pragma unroll
  for (int k = 0; k < CC; ++k) {
    if (c + k < p.dim) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (n0 + i < N)
          atomicAdd(p.dA + (int64_t)(c + k) * N + n0 + i, reinterpret_cast<const float*>(dAacc[k])[i]);
    }
  }
}

template <typename T, int NG, int CC, int kWarps, bool kVec, bool kRecomputeY, int NS>
static int launch_scan_bwd_y(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  using Cfg = ScanBwdCfg<T, NG, CC, kWarps, kVec, NS>;
  const int nchunks = (p.seqlen + MTTS_SCAN_CHUNK - 1) / MTTS_SCAN_CHUNK;
  const size_t smem = sizeof(float) * Cfg::smem_floats(kRecomputeY);
  auto kern = scan_bwd_kernel<T, NG, CC, kWarps, kVec, kRecomputeY, NS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  const dim3 grid((p.dim + Cfg::kChan - 1) / Cfg::kChan, p.batch);
  kern<<<grid, Cfg::kThreads, smem, stream>>>(p, nchunks);
  return launch_status();
}

template <typename T, int NG, int CC, int kWarps, bool kVec, int NS = 1>
static int launch_scan_bwd(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  // y is only needed for dz; with the forward's y_pre at hand the <C, h> recompute is skipped
  if (p.z && !p.y_pre) return launch_scan_bwd_y<T, NG, CC, kWarps, kVec, true, NS>(p, stream);
  return launch_scan_bwd_y<T, NG, CC, kWarps, kVec, false, NS>(p, stream);
}

template <typename T, bool kVec>
static int dispatch_scan_bwd_n(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  const int N = p.dstate;
  if (N <= 4) return launch_scan_bwd<T, 1, 1, 2, kVec>(p, stream);
  if (N <= 8) return launch_scan_bwd<T, 2, 1, 2, kVec>(p, stream);
  if (N <= 16) return launch_scan_bwd<T, 4, 2, 1, kVec>(p, stream);
  // wide states: 16-state chunks in separate warps of one CTA, sharing its P / E phases and dt / dt*u / gy rows
  // (C4 shape, bf16: N 32 9.8 ms, N 64 20.1 ms; the time-parallel kernel: 19.2 / 36.6 ms)
  if (N <= 32) return launch_scan_bwd<T, 4, 2, 1, kVec, 2>(p, stream);
  return launch_scan_bwd<T, 4, 2, 1, kVec, 4>(p, stream);
}

template <typename T>
static int dispatch_scan_bwd(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  // very wide states or too few channels: the time-parallel kernel
  if (p.dstate > 64 || scan_use_wide(p.batch, p.dim, p.seqlen)) return dispatch_scan_bwd_wide(p, stream);
  const bool vec = vec_ok<T>(p.u, p.u_batch_stride, p.u_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.delta, p.delta_batch_stride, p.delta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.B, p.B_batch_stride, p.B_state_stride, p.seqlen) &&
                   vec_ok<T>(p.C, p.C_batch_stride, p.C_state_stride, p.seqlen) &&
                   vec_ok<T>(p.z, p.z_batch_stride, p.z_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.dout, p.dout_batch_stride, p.dout_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.du, p.du_batch_stride, p.du_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.ddelta, p.ddelta_batch_stride, p.ddelta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.dz, p.dz_batch_stride, p.dz_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.y_pre, p.y_batch_stride, p.y_dim_stride, p.seqlen) &&
                   aligned16(p.dB) && aligned16(p.dC);  // 8-byte vector REDs
  return vec ? dispatch_scan_bwd_n<T, true>(p, stream) : dispatch_scan_bwd_n<T, false>(p, stream);
}

}  // namespace mtts

extern "C" int mtts_selective_scan_bwd(const mtts_scan_bwd_params* p, mtts_stream_t stream) {
  if (!p || !p->u || !p->delta || !p->A || !p->B || !p->C || !p->dout || !p->checkpoints ||
      !p->du || !p->ddelta || !p->dA || !p->dB || !p->dC)
    return MTTS_ERR_NULL;
  if ((p->z && !p->dz) || (p->D && !p->dD) || (p->delta_bias && !p->ddelta_bias))
    return MTTS_ERR_NULL;
  if (p->batch < 0 || p->dim < 0 || p->seqlen < 0 || p->dstate < 1 ||
      p->dstate > MTTS_MAX_DSTATE || p->batch > 65535)
    return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->dim == 0 || p->seqlen == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::dispatch_scan_bwd<float>(*p, s);
    case MTTS_BF16: return mtts::dispatch_scan_bwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}
