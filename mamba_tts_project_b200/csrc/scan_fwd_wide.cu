// selective_scan forward, time-parallel variant: used when batch x dim is too small to fill the GPU with
// the time-sequential kernel of scan_fwd.cu (which holds the C entry point).  Math: see mamba_tts_b200.h.
//
// Per (lane, pair of dstate rows): one MUFU.EX2 per timestep-state (the binding unit on B200:
// 16/clk/SM); everything else is packed fp32x2 (FFMA2/FMUL2: the two rows of a pair share an issue
// slot).  Sweep 1 builds the lane-local affine map, a 5-step shuffle scan stitches the 32 lanes,
// sweep 2 replays the recurrence from the true incoming state and contracts with C.
#include "scan_common.cuh"

namespace mtts {

template <typename T, int kItems, int kWarps, int kCPW, bool kVec>
__global__ void __launch_bounds__(kWarps * 32, 2)
scan_fwd_wide_kernel(const mtts_scan_fwd_params p, const int nchunks) {
  using Tile = PairTile<kItems>;
  constexpr int kThreads = kWarps * 32;
  constexpr int G = kWarps * kCPW;
  constexpr int kLanesPerChunk = MTTS_SCAN_CHUNK / kItems;
  constexpr int kChunksPerTile = Tile::kLen / MTTS_SCAN_CHUNK;

  extern __shared__ __align__(16) float smem[];
  const int N = p.dstate, L = p.seqlen;
  const int NP = (N + 1) >> 1;  // dstate row pairs
  float* Bs = smem;
  float* Cs = Bs + Tile::kPairs * Tile::kRow;
  float2* A2s = reinterpret_cast<float2*>(Cs + Tile::kPairs * Tile::kRow);  // A*log2(e), [G][NP]
  float2* hs = A2s + G * NP;                                                // running state [G][NP]

  const int b = blockIdx.y, c0 = blockIdx.x * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int idx = threadIdx.x; idx < G * NP * 2; idx += kThreads) {
    const int cl = idx / (2 * NP), n = idx - cl * 2 * NP, c = c0 + cl;
    float a2 = 0.f, h = 0.f;
    if (c < p.dim && n < N) {
      a2 = p.A[(int64_t)c * N + n] * kLog2e;
      const int64_t bc = (int64_t)b * p.dim + c;
      if (p.initial_state) h = p.initial_state[bc * N + n];
      if (p.checkpoints) p.checkpoints[bc * nchunks * N + n] = h;
    }
    reinterpret_cast<float*>(A2s)[idx] = a2;
    reinterpret_cast<float*>(hs)[idx] = h;
  }
  __syncthreads();

  const T* Bb = reinterpret_cast<const T*>(p.B) + (int64_t)b * p.B_batch_stride;
  const T* Cb = reinterpret_cast<const T*>(p.C) + (int64_t)b * p.C_batch_stride;
  const int ntiles = (L + Tile::kLen - 1) / Tile::kLen;
  const bool restage_per_pass = N > kScanNChunk;

  for (int tile = 0; tile < ntiles; ++tile) {
    const int t0 = tile * Tile::kLen;
    const int tl = t0 + lane * kItems;
    const bool partial = t0 + Tile::kLen > L;
#pragma unroll 1
    for (int pass = 0; pass < kCPW; ++pass) {
      const int cl = pass * kWarps + warp;
      const int c = c0 + cl;
      const bool cvalid = c < p.dim;  // warp-uniform

      float dl[kItems], du[kItems], y[kItems];
      float dsum = 0.f;
      if (cvalid) {
        const T* urow = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_batch_stride +
                        (int64_t)c * p.u_dim_stride;
        const T* drow = reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_batch_stride +
                        (int64_t)c * p.delta_dim_stride;
        load_items<T, kItems, kVec>(urow, tl, L, du);
        load_items<T, kItems, kVec>(drow, tl, L, dl);
        const float bias = p.delta_bias ? p.delta_bias[c] : 0.f;
        const float Dv = p.D ? p.D[c] : 0.f;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
          float x = dl[i] + bias;
          if (p.delta_softplus) x = softplus_f(x);
          if (partial && tl + i >= L) x = 0.f;  // padding: decay 1, input 0 = identity step
          const float uu = du[i];
          dl[i] = x;
          y[i] = Dv * uu;
          du[i] = x * uu;
          dsum += x;
        }
      } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i) dl[i] = du[i] = y[i] = 0.f;
      }

      for (int n0 = 0; n0 < N; n0 += kScanNChunk) {
        const int ncnt = min(kScanNChunk, N - n0);
        if (restage_per_pass || pass == 0) {
          __syncthreads();  // every warp is done reading the previous B/C tile
          stage_pairs<T, kItems, kVec, kThreads>(Bb, p.B_state_stride, n0, ncnt, t0, L, Bs);
          stage_pairs<T, kItems, kVec, kThreads>(Cb, p.C_state_stride, n0, ncnt, t0, L, Cs);
          __syncthreads();
        }
        if (!cvalid) continue;
        const int npairs = (ncnt + 1) >> 1;
#pragma unroll 1
        for (int pp = 0; pp < npairs; ++pp) {
          const int pg = (n0 >> 1) + pp;  // global pair index
          const float2 A2 = A2s[cl * NP + pg];
          const float2 h_in = hs[cl * NP + pg];
          const float* Bl = Bs + pp * Tile::kRow + lane * Tile::kSeg;
          const float* Cl = Cs + pp * Tile::kRow + lane * Tile::kSeg;
          float2 a[kItems], tmp[kItems];
          lane_pairs<kItems>(Bl, tmp);
          float2 hl = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < kItems; ++i) {
            a[i] = ex2f2(fmul2(dup2(dl[i]), A2));
            hl = ffma2(a[i], hl, fmul2(dup2(du[i]), tmp[i]));
          }
          float2 P = ex2f2(fmul2(dup2(dsum), A2));  // product of the lane's decays
          warp_scan_affine_up2(P, hl, lane);
          float2 Pe = shfl_up2(P, 1);
          float2 he = shfl_up2(hl, 1);
          if (lane == 0) {
            Pe = make_float2(1.f, 1.f);
            he = make_float2(0.f, 0.f);
          }
          float2 h = ffma2(Pe, h_in, he);  // state entering this lane's first timestep
          {
            float2 cv[kItems];
            lane_pairs<kItems>(Bl, tmp);  // b is recomputed rather than kept: 32 registers saved
            lane_pairs<kItems>(Cl, cv);
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
              h = ffma2(a[i], h, fmul2(dup2(du[i]), tmp[i]));
              y[i] = fmaf(h.y, cv[i].y, fmaf(h.x, cv[i].x, y[i]));
            }
          }
          // h = state after this lane's last timestep
          if (lane == 31) hs[cl * NP + pg] = h;
          if (p.checkpoints && ((lane + 1) % kLanesPerChunk) == 0) {
            const int k = tile * kChunksPerTile + (lane + 1) / kLanesPerChunk;
            if (k < nchunks) {
              float* ck = p.checkpoints + (((int64_t)b * p.dim + c) * nchunks + k) * N + 2 * pg;
              ck[0] = h.x;
              if (2 * pg + 1 < N) ck[1] = h.y;
            }
          }
        }
        __syncwarp();
      }

      if (cvalid) {
        if (p.y_pre) {
          T* yrow = reinterpret_cast<T*>(p.y_pre) + (int64_t)b * p.y_batch_stride + (int64_t)c * p.y_dim_stride;
          store_items<T, kItems, kVec>(yrow, tl, L, y);
        }
        if (p.z) {
          const T* zrow = reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_batch_stride +
                          (int64_t)c * p.z_dim_stride;
          float zv[kItems];
          load_items<T, kItems, kVec>(zrow, tl, L, zv);
#pragma unroll
          for (int i = 0; i < kItems; ++i) y[i] *= silu_f(zv[i]);
        }
        T* orow = reinterpret_cast<T*>(p.out) + (int64_t)b * p.out_batch_stride +
                  (int64_t)c * p.out_dim_stride;
        store_items<T, kItems, kVec>(orow, tl, L, y);
      }
    }
  }

  if (p.last_state) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * NP * 2; idx += kThreads) {
      const int cl = idx / (2 * NP), n = idx - cl * 2 * NP, c = c0 + cl;
      if (c < p.dim && n < N)
        p.last_state[((int64_t)b * p.dim + c) * N + n] = reinterpret_cast<const float*>(hs)[idx];
    }
  }
}

template <typename T, int kItems, int kWarps, int kCPW, bool kVec>
static int launch_scan_fwd_wide(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  using Tile = PairTile<kItems>;
  constexpr int G = kWarps * kCPW;
  const int nchunks = (p.seqlen + MTTS_SCAN_CHUNK - 1) / MTTS_SCAN_CHUNK;
  const int NP = (p.dstate + 1) / 2;
  const size_t smem = sizeof(float) * (2 * Tile::kPairs * Tile::kRow + 4 * (size_t)G * NP);
  auto kern = scan_fwd_wide_kernel<T, kItems, kWarps, kCPW, kVec>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  const dim3 grid((p.dim + G - 1) / G, p.batch);
  kern<<<grid, kWarps * 32, smem, stream>>>(p, nchunks);
  return launch_status();
}

template <typename T>
static int dispatch_scan_fwd_wide_t(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  const bool vec = vec_ok<T>(p.u, p.u_batch_stride, p.u_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.delta, p.delta_batch_stride, p.delta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.B, p.B_batch_stride, p.B_state_stride, p.seqlen) &&
                   vec_ok<T>(p.C, p.C_batch_stride, p.C_state_stride, p.seqlen) &&
                   vec_ok<T>(p.z, p.z_batch_stride, p.z_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.out, p.out_batch_stride, p.out_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.y_pre, p.y_batch_stride, p.y_dim_stride, p.seqlen);
  const bool two = p.dstate <= kScanNChunk && p.dim >= 16;
  if (vec) {
    return two ? launch_scan_fwd_wide<T, 16, 8, 2, true>(p, stream)
               : launch_scan_fwd_wide<T, 16, 8, 1, true>(p, stream);
  }
  return two ? launch_scan_fwd_wide<T, 16, 8, 2, false>(p, stream)
             : launch_scan_fwd_wide<T, 16, 8, 1, false>(p, stream);
}

int dispatch_scan_fwd_wide(const mtts_scan_fwd_params& p, cudaStream_t stream) {
  switch (p.io_dtype) {
    case MTTS_F32: return dispatch_scan_fwd_wide_t<float>(p, stream);
    case MTTS_BF16: return dispatch_scan_fwd_wide_t<__nv_bfloat16>(p, stream);
    default: return MTTS_ERR_DTYPE;
  }
}

}  // namespace mtts
