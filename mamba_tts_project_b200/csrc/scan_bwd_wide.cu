// selective_scan backward, wide-state variant (dstate > 64; never reached by the decoder, kept so the
// operator covers upstream's full dstate range).  scan_bwd.cu holds the main kernel and the C entry.
//
// Recompute-based: the forward saved only the state at the start of every MTTS_SCAN_CHUNK (=256)
// timesteps.  Tiles are walked last-to-first; inside a tile each warp (one channel) re-runs the
// forward recurrence for one dstate row, then runs the reverse-time recurrence
//     g_t = a_t * (C_t dy_t + g_{t+1}),   dh_t = C_t dy_t + g_{t+1}
// with the same lane-local sweep + warp-shuffle stitch as the forward, and accumulates
//     dC_t += dy_t h_t          dB_t += dh_t (dl_t u_t)          (summed over channels)
//     d(dl_t u_t) += dh_t B_t   d dl_t += g_t h_{t-1} A          dA += g_t h_{t-1} dl_t
// The cross-channel sums for dB/dC are reduced over the CTA's channels in shared memory first
// (one fp32 RED per (state, timestep) per CTA instead of upstream's one per channel).
#include "scan_common.cuh"

namespace mtts {

template <typename T, int kWarps, bool kVec>
__global__ void __launch_bounds__(kWarps * 32, 2)
scan_bwd_wide_kernel(const mtts_scan_bwd_params p, const int nchunks) {
  constexpr int kItems = 8;
  using Tile = PairTile<kItems>;
  constexpr int kChunksPerTile = Tile::kLen / MTTS_SCAN_CHUNK;  // the forward checkpoints every 32 steps
  constexpr int kThreads = kWarps * 32;
  constexpr int G = kWarps;
  constexpr int kRed = 4 * Tile::kLen;  // per warp: {dB, dC} x {row 2p, row 2p+1} x timesteps

  extern __shared__ __align__(16) float smem[];
  const int N = p.dstate, L = p.seqlen;
  const int NP = (N + 1) >> 1;
  float* Bs = smem;
  float* Cs = Bs + Tile::kPairs * Tile::kRow;
  float* red = Cs + Tile::kPairs * Tile::kRow;                 // [kWarps][kRed]
  float2* A2s = reinterpret_cast<float2*>(red + kWarps * kRed);  // [G][NP] A * log2(e)
  float2* hs = A2s + G * NP;                                   // [G][NP] state at tile start
  float2* gs = hs + G * NP;                                    // [G][NP] reverse carry
  // dA partials: per lane ([G][NP][32]) when dstate <= 16, else already warp-reduced ([G][NP])
  float2* dAs = gs + G * NP;
  const bool lane_da = N <= kScanNChunk;
  const int da_stride = lane_da ? 32 : 1;

  const int b = blockIdx.y, c0 = blockIdx.x * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = c0 + warp;
  const bool cvalid = c < p.dim;

  for (int idx = threadIdx.x; idx < G * NP * 2; idx += kThreads) {
    const int cl = idx / (2 * NP), n = idx - cl * 2 * NP;
    reinterpret_cast<float*>(A2s)[idx] =
        (c0 + cl < p.dim && n < N) ? p.A[(int64_t)(c0 + cl) * N + n] * kLog2e : 0.f;
    reinterpret_cast<float*>(gs)[idx] = 0.f;
  }
  for (int idx = threadIdx.x; idx < G * NP * da_stride; idx += kThreads) dAs[idx] = make_float2(0.f, 0.f);

  const T* Bb = reinterpret_cast<const T*>(p.B) + (int64_t)b * p.B_batch_stride;
  const T* Cb = reinterpret_cast<const T*>(p.C) + (int64_t)b * p.C_batch_stride;
  const int64_t cc = cvalid ? c : 0;
  const T* urow = reinterpret_cast<const T*>(p.u) + (int64_t)b * p.u_batch_stride + cc * p.u_dim_stride;
  const T* drow = reinterpret_cast<const T*>(p.delta) + (int64_t)b * p.delta_batch_stride +
                  cc * p.delta_dim_stride;
  const T* gorow = reinterpret_cast<const T*>(p.dout) + (int64_t)b * p.dout_batch_stride +
                   cc * p.dout_dim_stride;
  const T* zrow = p.z ? reinterpret_cast<const T*>(p.z) + (int64_t)b * p.z_batch_stride +
                            cc * p.z_dim_stride
                      : nullptr;
  const float bias = (cvalid && p.delta_bias) ? p.delta_bias[c] : 0.f;
  const float Dv = (cvalid && p.D) ? p.D[c] : 0.f;

  float dD_acc = 0.f, dbias_acc = 0.f;
  const int ntiles = (L + Tile::kLen - 1) / Tile::kLen;

  for (int tile = ntiles - 1; tile >= 0; --tile) {
    const int t0 = tile * Tile::kLen;
    const int tl = t0 + lane * kItems;

    // state at the start of this tile for every channel of the group
    for (int idx = threadIdx.x; idx < G * NP * 2; idx += kThreads) {
      const int cl = idx / (2 * NP), n = idx - cl * 2 * NP;
      reinterpret_cast<float*>(hs)[idx] =
          (c0 + cl < p.dim && n < N)
              ? p.checkpoints[(((int64_t)b * p.dim + c0 + cl) * nchunks + tile * kChunksPerTile) * N + n]
              : 0.f;
    }

    float dl[kItems], du[kItems], dy[kItems], y[kItems], ddu[kItems], ddl[kItems];
    float dsum = 0.f;
    if (cvalid) {
      float u[kItems];
      load_items<T, kItems, kVec>(urow, tl, L, u);
      load_items<T, kItems, kVec>(drow, tl, L, dl);
      load_items<T, kItems, kVec>(gorow, tl, L, dy);
      if (zrow) {
        float zv[kItems];
        load_items<T, kItems, kVec>(zrow, tl, L, zv);
#pragma unroll
        for (int i = 0; i < kItems; ++i) dy[i] *= silu_f(zv[i]);
      }
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        float x = dl[i] + bias;
        if (p.delta_softplus) x = softplus_f(x);
        if (tl + i >= L) x = 0.f;
        dl[i] = x;
        du[i] = x * u[i];
        y[i] = Dv * u[i];
        dsum += x;
        ddu[i] = 0.f;
        ddl[i] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kItems; ++i) dl[i] = du[i] = dy[i] = y[i] = ddu[i] = ddl[i] = 0.f;
    }

    for (int n0 = 0; n0 < N; n0 += kScanNChunk) {
      const int ncnt = min(kScanNChunk, N - n0);
      __syncthreads();
      stage_pairs<T, kItems, kVec, kThreads>(Bb, p.B_state_stride, n0, ncnt, t0, L, Bs);
      stage_pairs<T, kItems, kVec, kThreads>(Cb, p.C_state_stride, n0, ncnt, t0, L, Cs);
      __syncthreads();
      const int npairs = (ncnt + 1) >> 1;

#pragma unroll 1
      for (int pp = 0; pp < npairs; ++pp) {
        const int pg = (n0 >> 1) + pp;
        const float2 A2 = A2s[warp * NP + pg];
        const float2 An = fmul2(A2, dup2(kLn2));
        const float2 h_in = hs[warp * NP + pg];
        const float2 g_in = gs[warp * NP + pg];
        const float* Bl = Bs + pp * Tile::kRow + lane * Tile::kSeg;
        const float* Cl = Cs + pp * Tile::kRow + lane * Tile::kSeg;

        float2 a[kItems], h[kItems], cd[kItems];
        {
          float2 bv[kItems];
          lane_pairs<kItems>(Bl, bv);
          float2 hl = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < kItems; ++i) {
            a[i] = ex2f2(fmul2(dup2(dl[i]), A2));
            h[i] = fmul2(dup2(du[i]), bv[i]);
            hl = ffma2(a[i], hl, h[i]);
          }
          const float2 P0 = ex2f2(fmul2(dup2(dsum), A2));
          float2 Pf = P0;
          warp_scan_affine_up2(Pf, hl, lane);
          float2 Pe = shfl_up2(Pf, 1);
          float2 he = shfl_up2(hl, 1);
          if (lane == 0) {
            Pe = make_float2(1.f, 1.f);
            he = make_float2(0.f, 0.f);
          }
          const float2 hstart = ffma2(Pe, h_in, he);
          lane_pairs<kItems>(Cl, cd);
          float2 hp = hstart;
#pragma unroll
          for (int i = 0; i < kItems; ++i) {
            h[i] = ffma2(a[i], hp, h[i]);
            hp = h[i];
            y[i] = fmaf(h[i].y, cd[i].y, fmaf(h[i].x, cd[i].x, y[i]));
            cd[i] = fmul2(cd[i], dup2(dy[i]));  // C_t * dy_t
          }
          // reverse-time lane-local sweep
          float2 gl = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = kItems - 1; i >= 0; --i) gl = fmul2(a[i], fadd2(gl, cd[i]));
          float2 Pr = P0;
          warp_scan_affine_down2(Pr, gl, lane);
          float2 Pn = shfl_down2(Pr, 1);
          float2 gn = shfl_down2(gl, 1);
          if (lane == 31) {
            Pn = make_float2(1.f, 1.f);
            gn = make_float2(0.f, 0.f);
          }
          float2 g = ffma2(Pn, g_in, gn);  // g entering this lane's last timestep from the future
          if (lane == 0) gs[warp * NP + pg] = ffma2(Pr, g_in, gl);  // carry for the previous tile

          lane_pairs<kItems>(Bl, bv);
          float2 dA_acc = make_float2(0.f, 0.f);
          float* rb = red + warp * kRed + lane * kItems;
          float dBx[kItems], dBy[kItems], dCx[kItems], dCy[kItems];
#pragma unroll
          for (int i = kItems - 1; i >= 0; --i) {
            const float2 dh = fadd2(cd[i], g);
            g = fmul2(a[i], dh);
            const float2 dCc = fmul2(dup2(dy[i]), h[i]);
            const float2 dBc = fmul2(dh, dup2(du[i]));
            dBx[i] = dBc.x; dBy[i] = dBc.y; dCx[i] = dCc.x; dCy[i] = dCc.y;
            ddu[i] = fmaf(dh.y, bv[i].y, fmaf(dh.x, bv[i].x, ddu[i]));
            const float2 hprev = (i > 0) ? h[i - 1] : hstart;
            const float2 w = fmul2(g, hprev);
            ddl[i] = fmaf(w.y, An.y, fmaf(w.x, An.x, ddl[i]));
            dA_acc = ffma2(w, dup2(dl[i]), dA_acc);
          }
          if (lane_da) {
            float2* dap = dAs + (warp * NP + pg) * 32 + lane;
            *dap = fadd2(*dap, dA_acc);
          } else {
            const float sx = warp_sum(dA_acc.x), sy = warp_sum(dA_acc.y);
            if (lane == 0) {
              float2* dap = dAs + warp * NP + pg;
              *dap = fadd2(*dap, make_float2(sx, sy));
            }
          }
#pragma unroll
          for (int j = 0; j < kItems; j += 4) {
            *reinterpret_cast<float4*>(rb + j) = make_float4(dBx[j], dBx[j + 1], dBx[j + 2], dBx[j + 3]);
            *reinterpret_cast<float4*>(rb + Tile::kLen + j) =
                make_float4(dBy[j], dBy[j + 1], dBy[j + 2], dBy[j + 3]);
            *reinterpret_cast<float4*>(rb + 2 * Tile::kLen + j) =
                make_float4(dCx[j], dCx[j + 1], dCx[j + 2], dCx[j + 3]);
            *reinterpret_cast<float4*>(rb + 3 * Tile::kLen + j) =
                make_float4(dCy[j], dCy[j + 1], dCy[j + 2], dCy[j + 3]);
          }
        }
        __syncthreads();
        // sum over the CTA's channels, then one RED per (state, timestep)
        {
          // thread -> (slot s in [0, 4): {dB row0, dB row1, dC row0, dC row1}, 4 timesteps)
          const int s4 = threadIdx.x * 4;          // kThreads * 4 == kRed
          const int slot = s4 / Tile::kLen, tp = s4 - slot * Tile::kLen;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < kWarps; ++w) {
            const float4 v = *reinterpret_cast<const float4*>(red + w * kRed + s4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          const int n = 2 * pg + (slot & 1);
          if (n < N) {
            float* dst = ((slot >> 1) == 0 ? p.dB : p.dC) + ((int64_t)b * N + n) * L + t0 + tp;
            if (kVec && t0 + tp + 3 < L) {
              // one 16-byte reduction instead of four (L % 4 == 0 on the vector path)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc.x),
                           "f"(acc.y), "f"(acc.z), "f"(acc.w)
                           : "memory");
            } else {
              if (t0 + tp < L) atomicAdd(dst, acc.x);
              if (t0 + tp + 1 < L) atomicAdd(dst + 1, acc.y);
              if (t0 + tp + 2 < L) atomicAdd(dst + 2, acc.z);
              if (t0 + tp + 3 < L) atomicAdd(dst + 3, acc.w);
            }
          }
        }
        __syncthreads();  // red is single-buffered (keeps two CTAs per SM)
      }
    }

    // per-timestep outputs of this tile
    if (cvalid) {
      float tmp[kItems], u[kItems];
      load_items<T, kItems, kVec>(drow, tl, L, tmp);
      load_items<T, kItems, kVec>(urow, tl, L, u);
      float o_du[kItems], o_dd[kItems];
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const float x = tmp[i] + bias;
        const float sg = (p.delta_softplus && x <= 20.f) ? sigmoid_f(x) : 1.f;
        const bool in = tl + i < L;
        o_du[i] = fmaf(ddu[i], dl[i], dy[i] * Dv);
        const float dd = in ? fmaf(ddu[i], u[i], ddl[i]) * sg : 0.f;
        o_dd[i] = dd;
        dbias_acc += dd;
        dD_acc = fmaf(dy[i], u[i], dD_acc);
      }
      T* du_row = reinterpret_cast<T*>(p.du) + (int64_t)b * p.du_batch_stride + (int64_t)c * p.du_dim_stride;
      T* dd_row = reinterpret_cast<T*>(p.ddelta) + (int64_t)b * p.ddelta_batch_stride +
                  (int64_t)c * p.ddelta_dim_stride;
      store_items<T, kItems, kVec>(du_row, tl, L, o_du);
      store_items<T, kItems, kVec>(dd_row, tl, L, o_dd);
      if (zrow) {
        float zv[kItems], go[kItems];
        load_items<T, kItems, kVec>(zrow, tl, L, zv);
        load_items<T, kItems, kVec>(gorow, tl, L, go);
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
          const float sig = sigmoid_f(zv[i]);
          go[i] = go[i] * y[i] * sig * fmaf(zv[i], 1.f - sig, 1.f);
        }
        T* dz_row = reinterpret_cast<T*>(p.dz) + (int64_t)b * p.dz_batch_stride + (int64_t)c * p.dz_dim_stride;
        store_items<T, kItems, kVec>(dz_row, tl, L, go);
      }
    }
  }

  __syncthreads();
  if (cvalid) {
    dD_acc = warp_sum(dD_acc);
    dbias_acc = warp_sum(dbias_acc);
    if (lane == 0) {
      if (p.dD) atomicAdd(p.dD + c, dD_acc);
      if (p.ddelta_bias) atomicAdd(p.ddelta_bias + c, dbias_acc);
    }
    for (int pg = 0; pg < NP; ++pg) {
      float sx, sy;
      if (lane_da) {
        const float2 v = dAs[(warp * NP + pg) * 32 + lane];
        sx = warp_sum(v.x);
        sy = warp_sum(v.y);
      } else {
        const float2 v = dAs[warp * NP + pg];
        sx = v.x;
        sy = v.y;
      }
      if (lane == 0) {
        atomicAdd(p.dA + (int64_t)c * N + 2 * pg, sx);
        if (2 * pg + 1 < N) atomicAdd(p.dA + (int64_t)c * N + 2 * pg + 1, sy);
      }
    }
  }
}

template <typename T, bool kVec>
static int launch_scan_bwd_wide(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  constexpr int kWarps = 8;
  using Tile = PairTile<8>;
  const int nchunks = (p.seqlen + MTTS_SCAN_CHUNK - 1) / MTTS_SCAN_CHUNK;
  const int NP = (p.dstate + 1) / 2;
  const size_t smem = sizeof(float) * (2 * Tile::kPairs * Tile::kRow + kWarps * 4 * Tile::kLen +
                                       (size_t)kWarps * NP * (6 + (p.dstate <= kScanNChunk ? 64 : 2)));
  auto kern = scan_bwd_wide_kernel<T, kWarps, kVec>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  const dim3 grid((p.dim + kWarps - 1) / kWarps, p.batch);
  kern<<<grid, kWarps * 32, smem, stream>>>(p, nchunks);
  return launch_status();
}

template <typename T>
static int dispatch_scan_bwd_wide_t(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  const bool vec = vec_ok<T>(p.u, p.u_batch_stride, p.u_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.delta, p.delta_batch_stride, p.delta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.B, p.B_batch_stride, p.B_state_stride, p.seqlen) &&
                   vec_ok<T>(p.C, p.C_batch_stride, p.C_state_stride, p.seqlen) &&
                   vec_ok<T>(p.z, p.z_batch_stride, p.z_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.dout, p.dout_batch_stride, p.dout_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.du, p.du_batch_stride, p.du_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.ddelta, p.ddelta_batch_stride, p.ddelta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.dz, p.dz_batch_stride, p.dz_dim_stride, p.seqlen);
  return vec ? launch_scan_bwd_wide<T, true>(p, stream) : launch_scan_bwd_wide<T, false>(p, stream);
}

int dispatch_scan_bwd_wide(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  switch (p.io_dtype) {
    case MTTS_F32: return dispatch_scan_bwd_wide_t<float>(p, stream);
    case MTTS_BF16: return dispatch_scan_bwd_wide_t<__nv_bfloat16>(p, stream);
    default: return MTTS_ERR_DTYPE;
  }
}

}  // namespace mtts
