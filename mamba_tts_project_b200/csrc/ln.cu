// Fused residual-add + LayerNorm + FiLM, forward and backward, for sm_100a.
//
// The reference layer (mamba_decoder.py:59-89) is three times  x = x + branch(...) ; h = LN(x)
// [; h = gamma * h + beta].  Each "x + branch" is folded into the LayerNorm that consumes it:
//     x_out = x + delta            (fp32 residual stream, delta = previous branch output, io dtype)
//     out   = FiLM(LN(x_out))      (io dtype: the next GEMM's operand)
// One warp owns one row; a lane owns 4-element column groups {lane, lane+32, ...}, so every row
// statistic is a shuffle reduction and every column statistic (backward) stays in the lane's
// registers across the rows the warp walks.  HBM-bound: each tensor is touched exactly once.
#include "common.cuh"

namespace mtts {

constexpr int kLnWarps = 4;

// kG = 4-element column groups per lane (dim <= 128 * kG)
template <typename T, int kG>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layernorm_fwd_kernel(const mtts_add_layernorm_fwd_params p) {
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  const int Dm = p.dim;
  const float* x = p.x + (int64_t)row * Dm;
  const T* dl = p.delta ? reinterpret_cast<const T*>(p.delta) + (int64_t)row * Dm : nullptr;
  float v[kG][4];
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const int e = (g * 32 + lane) * 4;
    if (e < Dm) {
      load4<float>(x + e, v[g]);
      if (dl) {
        float d[4];
        load4<T>(dl + e, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[g][j] += d[j];
        if (p.delta_bias) {  // bias of the Linear that produced delta, folded in here
          load4<float>(p.delta_bias + e, d);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[g][j] += d[j];
        }
      }
      if (p.x_out) store4<float>(p.x_out + (int64_t)row * Dm + e, v[g]);
#pragma unroll
      for (int j = 0; j < 4; ++j) s += v[g][j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[g][j] = 0.f;
    }
  }
  // affine / FiLM operands do not depend on the row statistics: request them before the reductions so the
  // row costs two dependent memory round trips instead of three (decode: 64 rows, pure latency)
  const int bidx = row / p.rows_per_batch;
  const float* gam = p.film_gamma ? p.film_gamma + (int64_t)bidx * Dm : nullptr;
  const float* bet = p.film_beta ? p.film_beta + (int64_t)bidx * Dm : nullptr;
  // (rows wider than 512: 64 more registers per thread would halve the resident warps of a kernel that lives on
  // bytes in flight -- the operands are fetched where they are used instead, from L1)
  constexpr bool kPre = kG <= 4;
  float w[kPre ? kG : 1][4], b[kPre ? kG : 1][4];
#pragma unroll
  for (int g = 0; g < (kPre ? kG : 0); ++g) {
    const int e = (g * 32 + lane) * 4;
    if (e < Dm) {
      load4<float>(p.ln_weight + e, w[g]);
      load4<float>(p.ln_bias + e, b[g]);
      if (gam) {  // out = gamma (xhat w + b) + beta = xhat (gamma w) + (gamma b + beta)
        float gm[4], bt[4];
        load4<float>(gam + e, gm);
        load4<float>(bet + e, bt);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w[g][j] *= gm[j];
          b[g][j] = fmaf(gm[j], b[g][j], bt[j]);
        }
      }
    }
  }
  const float mean = warp_sum(s) / (float)Dm;
  float q = 0.f;
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const int e = (g * 32 + lane) * 4;
    if (e < Dm) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float c = v[g][j] - mean;
        q = fmaf(c, c, q);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)Dm + p.eps);
  if (lane == 0) {
    if (p.mean) p.mean[row] = mean;
    if (p.rstd) p.rstd[row] = rstd;
  }
  T* out = reinterpret_cast<T*>(p.out) + (int64_t)row * Dm;
#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const int e = (g * 32 + lane) * 4;
    if (e < Dm) {
      float o[4];
      if constexpr (kPre) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf((v[g][j] - mean) * rstd, w[g][j], b[g][j]);
      } else {
        float wv[4], bv[4];
        load4<float>(p.ln_weight + e, wv);
        load4<float>(p.ln_bias + e, bv);
        if (gam) {
          float gm[4], bt[4];
          load4<float>(gam + e, gm);
          load4<float>(bet + e, bt);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            wv[j] *= gm[j];
            bv[j] = fmaf(gm[j], bv[j], bt[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf((v[g][j] - mean) * rstd, wv[j], bv[j]);
      }
      store4<T>(out + e, o);
    }
  }
}

// Backward.  With xhat = (x_out - mean) rstd, y = xhat w + b, out = gamma y + beta:
//   g      = dout * gamma * w                       (d/d xhat)
//   dx     = dx_out + rstd (g - mean_e(g) - xhat mean_e(g xhat))
//   S1[b]  = sum_t dout xhat,  S2[b] = sum_t dout,  S3[b] = sum_t dx  (per batch element and column,
//            accumulated into `colsum` (batch, 3, dim)); S3 is the gradient of delta's bias; the host finishes dw = sum_b gamma_b S1_b, db = sum_b gamma_b S2_b,
//            dgamma_b = w S1_b + bias S2_b, dbeta_b = S2_b  on (batch, dim)-sized tensors.
// warps per CTA of the backward: the per-warp column statistics (3 x dim floats) must fit the static 48 KB
template <int kG>
__host__ __device__ constexpr int ln_bwd_warps() { return kG <= 4 ? 4 : 2; }

template <typename T, int kG>
__global__ void __launch_bounds__(ln_bwd_warps<kG>() * 32, kG <= 4 ? 5 : 7)
add_layernorm_bwd_kernel(const mtts_add_layernorm_bwd_params p, const int rows_per_warp) {
  constexpr int kLnWarps = ln_bwd_warps<kG>();     // (shadows the forward's constant inside this kernel)
  // Column statistics live in shared memory (one private row set per warp, each lane owns its columns, so
  // plain load/add/store): keeping them in registers cost 48 registers per thread and with them half the
  // resident warps -- the kernel is bound by bytes in flight, not by LSU slots.
  __shared__ __align__(16) float red[kLnWarps][3][kG * 128];
  // ln_weight * gamma_b lives in shared memory (frees 16-32 registers per thread)
  constexpr bool kWsm = true;
  __shared__ __align__(16) float wsm[kG * 128];
  float wreg[1][4];
  constexpr bool kLateUp = false;    // (fetching the upstream gradient where it is used costs a second round trip per row: 186 -> 284 us at dim 1024)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dm = p.dim;
  // a CTA never straddles two batch elements: grid.x tiles rows_per_batch, grid.y = batch
  const int bidx = blockIdx.y;
  const int r_begin = blockIdx.x * kLnWarps * rows_per_warp + warp * rows_per_warp;
  const int r_end = min(p.rows_per_batch, r_begin + rows_per_warp);
  const float* gam = p.film_gamma ? p.film_gamma + (int64_t)bidx * Dm : nullptr;

#pragma unroll
  for (int g = 0; g < kG; ++g) {
    const int e = (g * 32 + lane) * 4;
    if (warp == 0 || !kWsm) {
      float wv[4] = {0.f, 0.f, 0.f, 0.f};
      if (e < Dm) {
        load4<float>(p.ln_weight + e, wv);
        if (gam) {
          float gm[4];
          load4<float>(gam + e, gm);
#pragma unroll
          for (int j = 0; j < 4; ++j) wv[j] *= gm[j];
        }
      }
      if constexpr (kWsm) {
        *reinterpret_cast<float4*>(&wsm[e]) = make_float4(wv[0], wv[1], wv[2], wv[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) wreg[g][j] = wv[j];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      *reinterpret_cast<float4*>(&red[warp][k][e]) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  for (int r = r_begin; r < r_end; ++r) {
    const int64_t row = (int64_t)bidx * p.rows_per_batch + r;
    const float mean = p.mean[row], rstd = p.rstd[row];
    const float* xo = p.x_out + row * Dm;
    const T* go = reinterpret_cast<const T*>(p.dout) + row * Dm;
    float xh[kG][4], gg[kG][4], up[kG][4];
    float a = 0.f, bsum = 0.f;
    // the NEXT row of this warp is asked into L2 now (one 128-byte line per lane and tensor, no registers): its loads
    // then cost an L2 hit instead of an HBM round trip per row of the warp's dependent row-by-row chain
    if (r + 1 < r_end) {
      const int64_t nrow = row + 1;
      for (int off = lane * 128; off < Dm * 4; off += 32 * 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.x_out + nrow * Dm) + off));
        if (p.dx_out)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.dx_out + nrow * Dm) + off));
        if (off < Dm * (int)sizeof(T))
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(
              reinterpret_cast<const T*>(p.dout) + nrow * Dm) + off));
      }
    }
    // the upstream residual gradient is fetched together with x_out / dout (one round trip per row)
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const int e = (g * 32 + lane) * 4;
      if (!kLateUp && p.dx_out && e < Dm) {
        load4<float>(p.dx_out + row * Dm + e, up[g]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) up[g][j] = 0.f;
      }
    }
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const int e = (g * 32 + lane) * 4;
      if (e < Dm) {
        float xv[4], dv[4];
        load4<float>(xo + e, xv);
        load4<T>(go + e, dv);
        float4 c1 = *reinterpret_cast<const float4*>(&red[warp][0][e]);
        float4 c2 = *reinterpret_cast<const float4*>(&red[warp][1][e]);
        float* c1v = reinterpret_cast<float*>(&c1);
        float* c2v = reinterpret_cast<float*>(&c2);
        float wv[4];
        if constexpr (kWsm) {
          const float4 w4 = *reinterpret_cast<const float4*>(&wsm[e]);
          wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) wv[j] = wreg[g][j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xh[g][j] = (xv[j] - mean) * rstd;
          c1v[j] = fmaf(dv[j], xh[g][j], c1v[j]);
          c2v[j] += dv[j];
          gg[g][j] = dv[j] * wv[j];
          a += gg[g][j];
          bsum = fmaf(gg[g][j], xh[g][j], bsum);
        }
        *reinterpret_cast<float4*>(&red[warp][0][e]) = c1;
        *reinterpret_cast<float4*>(&red[warp][1][e]) = c2;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) xh[g][j] = gg[g][j] = 0.f;
      }
    }
    a = warp_sum(a) / (float)Dm;
    bsum = warp_sum(bsum) / (float)Dm;
#pragma unroll
    for (int g = 0; g < kG; ++g) {
      const int e = (g * 32 + lane) * 4;
      if (e < Dm) {
        float dx[4];
        if constexpr (kLateUp) {
          if (p.dx_out) load4<float>(p.dx_out + row * Dm + e, up[g]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) dx[j] = rstd * (gg[g][j] - a - xh[g][j] * bsum) + up[g][j];
        float4 c3 = *reinterpret_cast<const float4*>(&red[warp][2][e]);
        c3.x += dx[0]; c3.y += dx[1]; c3.z += dx[2]; c3.w += dx[3];
        *reinterpret_cast<float4*>(&red[warp][2][e]) = c3;
        store4<float>(p.dx + row * Dm + e, dx);
        if (p.ddelta) store4<T>(reinterpret_cast<T*>(p.ddelta) + row * Dm + e, dx);
      }
    }
  }

  // column sums: warps -> CTA -> one RED per column per CTA
  __syncthreads();
  // (16-byte vector REDs: a quarter of the atomic operations the L2 has to serialise per address group)
  const int nv = Dm / 4;
  for (int idx = threadIdx.x; idx < 3 * nv; idx += kLnWarps * 32) {
    const int which = idx / nv, e = (idx - which * nv) * 4;
    float4 t = *reinterpret_cast<const float4*>(&red[0][which][e]);
#pragma unroll
    for (int wv = 1; wv < kLnWarps; ++wv) {
      const float4 u = *reinterpret_cast<const float4*>(&red[wv][which][e]);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float* dst = p.colsum + ((int64_t)bidx * 3 + which) * Dm + e;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w)
                 : "memory");
  }
}

template <typename T>
static int dispatch_ln_fwd(const mtts_add_layernorm_fwd_params& p, cudaStream_t s) {
  const dim3 grid((p.rows + kLnWarps - 1) / kLnWarps);
  const int groups = (p.dim + 127) / 128;
  if (groups <= 2) add_layernorm_fwd_kernel<T, 2><<<grid, kLnWarps * 32, 0, s>>>(p);
  else if (groups <= 4) add_layernorm_fwd_kernel<T, 4><<<grid, kLnWarps * 32, 0, s>>>(p);
  else if (groups <= 8) add_layernorm_fwd_kernel<T, 8><<<grid, kLnWarps * 32, 0, s>>>(p);
  else add_layernorm_fwd_kernel<T, 16><<<grid, kLnWarps * 32, 0, s>>>(p);
  return launch_status();
}

template <typename T, int kG>
static int launch_ln_bwd(const mtts_add_layernorm_bwd_params& p, cudaStream_t s) {
  constexpr int W = ln_bwd_warps<kG>();
  const int batch = p.rows / p.rows_per_batch;
  // ONE wave of CTAs (a second, partly filled wave costs as much as a full one: 1.6 waves measured 66 us where
  // 0.9 waves take 5x us at (32768, 512)); every warp walks `rows_per_warp` consecutive rows of one batch element
  constexpr int kCtasPerSm = kG <= 4 ? 5 : 7;          // = the kernel's __launch_bounds__
  const int slots = kNumSMs * kCtasPerSm * W;
  int rows_per_warp = (p.rows + slots - 1) / slots;
  rows_per_warp = max(1, min(rows_per_warp, 64));
  // the grid tiles every batch element separately: shrink the tile until the whole grid fits the wave
  while (rows_per_warp < 64 &&
         (long long)((p.rows_per_batch + rows_per_warp * W - 1) / (rows_per_warp * W)) * batch > kNumSMs * kCtasPerSm)
    ++rows_per_warp;
  const int per_cta = rows_per_warp * W;
  const dim3 grid((p.rows_per_batch + per_cta - 1) / per_cta, batch);
  add_layernorm_bwd_kernel<T, kG><<<grid, W * 32, 0, s>>>(p, rows_per_warp);
  return launch_status();
}

template <typename T>
static int dispatch_ln_bwd(const mtts_add_layernorm_bwd_params& p, cudaStream_t s) {
  const int groups = (p.dim + 127) / 128;
  if (groups <= 2) return launch_ln_bwd<T, 2>(p, s);
  if (groups <= 4) return launch_ln_bwd<T, 4>(p, s);
  if (groups <= 8) return launch_ln_bwd<T, 8>(p, s);
  return MTTS_ERR_SHAPE;
}

}  // namespace mtts

namespace mtts {

// Parameter gradients of a FiLM'd LayerNorm from the backward kernel's per-batch column sums: one thread per column
// walks the batch (a (batch, dim)-sized problem: replaces two einsum, an addcmul, a mul and a sum launch).
__global__ void __launch_bounds__(128)
add_layernorm_film_finish_kernel(const mtts_add_layernorm_finish_params p) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= p.dim) return;
  const float w = p.ln_weight[d], bias = p.ln_bias[d];
  float dw = 0.f, db = 0.f, d3 = 0.f;
  for (int b = 0; b < p.batch; ++b) {
    const float* cs = p.colsum + (size_t)b * 3 * p.dim;
    const float s1 = cs[d], s2 = cs[p.dim + d];
    const float gm = p.film_gamma[(size_t)b * p.dim + d];
    dw = fmaf(gm, s1, dw);
    db = fmaf(gm, s2, db);
    d3 += cs[2 * p.dim + d];
    p.dgamma[(size_t)b * p.dim + d] = fmaf(w, s1, bias * s2);
    p.dbeta[(size_t)b * p.dim + d] = s2;
  }
  p.dweight[d] = dw;
  p.dbias[d] = db;
  if (p.ddelta_bias) p.ddelta_bias[d] = d3;
}

}  // namespace mtts

extern "C" int mtts_add_layernorm_bwd_finish(const mtts_add_layernorm_finish_params* p, mtts_stream_t stream) {
  if (!p || !p->colsum || !p->film_gamma || !p->ln_weight || !p->ln_bias || !p->dweight || !p->dbias || !p->dgamma ||
      !p->dbeta)
    return MTTS_ERR_NULL;
  if (p->batch < 1 || p->dim < 1) return MTTS_ERR_SHAPE;
  mtts::add_layernorm_film_finish_kernel<<<(p->dim + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return mtts::launch_status();
}

extern "C" int mtts_add_layernorm_fwd(const mtts_add_layernorm_fwd_params* p, mtts_stream_t stream) {
  if (!p || !p->x || !p->ln_weight || !p->ln_bias || !p->out) return MTTS_ERR_NULL;
  if ((p->film_gamma == nullptr) != (p->film_beta == nullptr)) return MTTS_ERR_NULL;
  if (p->rows < 0 || p->dim < 4 || p->dim % 4 != 0 || p->dim > 2048 || p->rows_per_batch < 1)
    return MTTS_ERR_SHAPE;
  if (p->rows == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::dispatch_ln_fwd<float>(*p, s);
    case MTTS_BF16: return mtts::dispatch_ln_fwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}

extern "C" int mtts_add_layernorm_bwd(const mtts_add_layernorm_bwd_params* p, mtts_stream_t stream) {
  if (!p || !p->x_out || !p->dout || !p->mean || !p->rstd || !p->ln_weight || !p->dx || !p->colsum)
    return MTTS_ERR_NULL;
  if (!mtts::aligned16(p->colsum)) return MTTS_ERR_ALIGN;   // 16-byte vector REDs
  if (p->rows < 0 || p->dim < 4 || p->dim % 4 != 0 || p->dim > 1024 || p->rows_per_batch < 1 ||
      p->rows % p->rows_per_batch != 0 || p->rows / p->rows_per_batch > 65535)
    return MTTS_ERR_SHAPE;
  if (p->rows == 0) return MTTS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: return mtts::dispatch_ln_bwd<float>(*p, s);
    case MTTS_BF16: return mtts::dispatch_ln_bwd<__nv_bfloat16>(*p, s);
    default: return MTTS_ERR_DTYPE;
  }
}
