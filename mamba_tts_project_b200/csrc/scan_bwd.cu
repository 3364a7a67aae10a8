// selective_scan backward for sm_100a.  Replaces selective_scan_cuda.bwd (mamba_ssm), reached from
// loss.backward() at train.py:231 through MambaInnerFn.backward.
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
scan_bwd_kernel(const mtts_scan_bwd_params p, const int nchunks) {
  constexpr int kItems = 8;
  using Tile = ScanTile<kItems>;
  static_assert(Tile::kLen == MTTS_SCAN_CHUNK, "backward tile == checkpoint chunk");
  constexpr int kThreads = kWarps * 32;
  constexpr int G = kWarps;

  extern __shared__ __align__(16) float smem[];
  const int N = p.dstate, L = p.seqlen;
  float* Bs = smem;
  float* Cs = Bs + kScanNChunk * Tile::kRow;
  float* red = Cs + kScanNChunk * Tile::kRow;        // [2][kWarps][2][kLen]
  float* A2s = red + 2 * kWarps * 2 * Tile::kLen;    // [G][N]  A * log2(e)
  float* hs = A2s + G * N;                           // [G][N]  state at tile start
  float* gs = hs + G * N;                            // [G][N]  reverse carry
  float* dAs = gs + G * N;                           // [G][N]  dA accumulator

  const int b = blockIdx.y, c0 = blockIdx.x * G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = c0 + warp;
  const bool cvalid = c < p.dim;

  for (int idx = threadIdx.x; idx < G * N; idx += kThreads) {
    const int cl = idx / N, n = idx - cl * N;
    A2s[idx] = (c0 + cl < p.dim) ? p.A[(int64_t)(c0 + cl) * N + n] * kLog2e : 0.f;
    gs[idx] = 0.f;
    dAs[idx] = 0.f;
  }

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
  int buf = 0;
  const int ntiles = (L + Tile::kLen - 1) / Tile::kLen;

  for (int tile = ntiles - 1; tile >= 0; --tile) {
    const int t0 = tile * Tile::kLen;
    const int tl = t0 + lane * kItems;

    // state at the start of this tile for every channel of the group
    for (int idx = threadIdx.x; idx < G * N; idx += kThreads) {
      const int cl = idx / N, n = idx - cl * N;
      hs[idx] = (c0 + cl < p.dim)
                    ? p.checkpoints[(((int64_t)b * p.dim + c0 + cl) * nchunks + tile) * N + n]
                    : 0.f;
    }

    float u[kItems], dl[kItems], du[kItems], dy[kItems], y[kItems], ddu[kItems], ddl[kItems];
    float dsum = 0.f;
    if (cvalid) {
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
      for (int i = 0; i < kItems; ++i) u[i] = dl[i] = du[i] = dy[i] = y[i] = ddu[i] = ddl[i] = 0.f;
    }

    for (int n0 = 0; n0 < N; n0 += kScanNChunk) {
      const int ncnt = min(kScanNChunk, N - n0);
      __syncthreads();
      stage_rows<T, kItems, kVec, kThreads>(Bb, p.B_state_stride, n0, ncnt, t0, L, Bs);
      stage_rows<T, kItems, kVec, kThreads>(Cb, p.C_state_stride, n0, ncnt, t0, L, Cs);
      __syncthreads();

#pragma unroll 1
      for (int nn = 0; nn < ncnt; ++nn) {
        const int n = n0 + nn;
        const float A2 = A2s[warp * N + n];
        const float An = A2 * kLn2;
        const float h_in = hs[warp * N + n];
        const float g_in = gs[warp * N + n];

        float a[kItems], h[kItems], bv[kItems], cd[kItems];
        lane_row<kItems>(Bs + nn * Tile::kRow + lane * Tile::kSeg, bv);
        float hl = 0.f;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
          a[i] = ex2f(dl[i] * A2);
          h[i] = du[i] * bv[i];
          hl = fmaf(a[i], hl, h[i]);
        }
        const float P = ex2f(A2 * dsum);
        float Pf = P;
        warp_scan_affine_up(Pf, hl, lane);
        float Pe = __shfl_up_sync(0xffffffffu, Pf, 1);
        float he = __shfl_up_sync(0xffffffffu, hl, 1);
        if (lane == 0) {
          Pe = 1.f;
          he = 0.f;
        }
        const float hstart = fmaf(Pe, h_in, he);
        lane_row<kItems>(Cs + nn * Tile::kRow + lane * Tile::kSeg, cd);
        {
          float hp = hstart;
#pragma unroll
          for (int i = 0; i < kItems; ++i) {
            h[i] = fmaf(a[i], hp, h[i]);
            hp = h[i];
            y[i] = fmaf(h[i], cd[i], y[i]);
            cd[i] *= dy[i];  // C_t * dy_t
          }
        }
        // reverse-time lane-local sweep
        float gl = 0.f;
#pragma unroll
        for (int i = kItems - 1; i >= 0; --i) gl = a[i] * (gl + cd[i]);
        float Pr = P;
        warp_scan_affine_down(Pr, gl, lane);
        float Pn = __shfl_down_sync(0xffffffffu, Pr, 1);
        float gn = __shfl_down_sync(0xffffffffu, gl, 1);
        if (lane == 31) {
          Pn = 1.f;
          gn = 0.f;
        }
        float g = fmaf(Pn, g_in, gn);  // g entering this lane's last timestep from the future
        const float g_out = fmaf(Pr, g_in, gl);  // lane 0: carry for the previous tile
        if (lane == 0) gs[warp * N + n] = g_out;

        float dBc[kItems], dCc[kItems];
        float dA_acc = 0.f;
#pragma unroll
        for (int i = kItems - 1; i >= 0; --i) {
          const float dh = cd[i] + g;
          g = a[i] * dh;
          dCc[i] = dy[i] * h[i];
          dBc[i] = dh * du[i];
          ddu[i] = fmaf(dh, bv[i], ddu[i]);
          const float hprev = (i > 0) ? h[i - 1] : hstart;
          const float w = g * hprev;
          ddl[i] = fmaf(w, An, ddl[i]);
          dA_acc = fmaf(w, dl[i], dA_acc);
        }
        dA_acc = warp_sum(dA_acc);
        if (lane == 0) dAs[warp * N + n] += dA_acc;

        float* rb = red + ((buf * kWarps + warp) * 2) * Tile::kLen + lane * kItems;
#pragma unroll
        for (int j = 0; j < kItems; j += 4) {
          *reinterpret_cast<float4*>(rb + j) = make_float4(dBc[j], dBc[j + 1], dBc[j + 2], dBc[j + 3]);
          *reinterpret_cast<float4*>(rb + Tile::kLen + j) =
              make_float4(dCc[j], dCc[j + 1], dCc[j + 2], dCc[j + 3]);
        }
        __syncthreads();
        // sum over the CTA's channels, then one RED per (state, timestep)
        {
          constexpr int kPairs = 2 * Tile::kLen / 2;  // float2 slots over both tensors
          for (int s = threadIdx.x; s < kPairs; s += kThreads) {
            const int tensor = s / (Tile::kLen / 2);
            const int tp = (s - tensor * (Tile::kLen / 2)) * 2;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
              const float2 v = *reinterpret_cast<const float2*>(
                  red + ((buf * kWarps + w) * 2 + tensor) * Tile::kLen + tp);
              acc.x += v.x;
              acc.y += v.y;
            }
            float* dst = (tensor == 0 ? p.dB : p.dC) + ((int64_t)b * N + n) * L + t0 + tp;
            if (t0 + tp < L) atomicAdd(dst, acc.x);
            if (t0 + tp + 1 < L) atomicAdd(dst + 1, acc.y);
          }
        }
        buf ^= 1;
      }
    }

    // per-timestep outputs of this tile
    if (cvalid) {
      float tmp[kItems], sg[kItems];
      load_items<T, kItems, kVec>(drow, tl, L, tmp);
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const float x = tmp[i] + bias;
        sg[i] = (p.delta_softplus && x <= 20.f) ? sigmoid_f(x) : 1.f;
      }
      float o_du[kItems], o_dd[kItems];
#pragma unroll
      for (int i = 0; i < kItems; ++i) {
        const bool in = tl + i < L;
        o_du[i] = fmaf(ddu[i], dl[i], dy[i] * Dv);
        const float dd = in ? fmaf(ddu[i], u[i], ddl[i]) * sg[i] : 0.f;
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
    for (int n = lane; n < N; n += 32) atomicAdd(p.dA + (int64_t)c * N + n, dAs[warp * N + n]);
  }
}

template <typename T, bool kVec>
static int launch_scan_bwd(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  constexpr int kWarps = 8;
  using Tile = ScanTile<8>;
  const int nchunks = (p.seqlen + MTTS_SCAN_CHUNK - 1) / MTTS_SCAN_CHUNK;
  const size_t smem = sizeof(float) * (2 * kScanNChunk * Tile::kRow + 2 * kWarps * 2 * Tile::kLen +
                                       4 * (size_t)kWarps * p.dstate);
  auto kern = scan_bwd_kernel<T, kWarps, kVec>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  const dim3 grid((p.dim + kWarps - 1) / kWarps, p.batch);
  kern<<<grid, kWarps * 32, smem, stream>>>(p, nchunks);
  return launch_status();
}

template <typename T>
static int dispatch_scan_bwd(const mtts_scan_bwd_params& p, cudaStream_t stream) {
  const bool vec = vec_ok<T>(p.u, p.u_batch_stride, p.u_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.delta, p.delta_batch_stride, p.delta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.B, p.B_batch_stride, p.B_state_stride, p.seqlen) &&
                   vec_ok<T>(p.C, p.C_batch_stride, p.C_state_stride, p.seqlen) &&
                   vec_ok<T>(p.z, p.z_batch_stride, p.z_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.dout, p.dout_batch_stride, p.dout_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.du, p.du_batch_stride, p.du_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.ddelta, p.ddelta_batch_stride, p.ddelta_dim_stride, p.seqlen) &&
                   vec_ok<T>(p.dz, p.dz_batch_stride, p.dz_dim_stride, p.seqlen);
  return vec ? launch_scan_bwd<T, true>(p, stream) : launch_scan_bwd<T, false>(p, stream);
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
