// The caller's side of the decoder's training step, for sm_100a:
//   embed_sum_fwd / _bwd   x = token_embed[tok] + pos_embed[pos] + quant_embed[q]       (mamba_decoder.py:167-171,
//                          train.py:115-131 `embed_codec_tokens`) and its scatter-add backward
//   ce_loss                cross entropy with ignore_index (train.py:31-42 `codec_ce_loss`): one pass over the
//                          logits gives the loss AND d loss / d logits
//   grad_sumsq / adam_step `clip_grad_norm_(decoder.parameters(), 1.0)` + `torch.optim.Adam.step()`
//                          (train.py:152-159,233-234) over a device-resident table of tensors: the clip
//                          coefficient is applied inside the Adam update, the gradients are never rescaled in memory
// All HBM-bound streaming kernels.
#include "common.cuh"

namespace mtts {
namespace {

// ---- embedding sum -----------------------------------------------------------------------------------------
// grid (seq positions, batch chunks); a CTA = one position l, threads over the model dimension in float4s
__global__ void __launch_bounds__(128)
embed_sum_fwd_kernel(const mtts_embed_sum_params p) {
  const int l = blockIdx.x;
  const int64_t pos = p.pos_ids[l];
  const int64_t q = p.quant_ids ? p.quant_ids[l] : -1;
  const int nvec = p.dim / 4;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float4 base = reinterpret_cast<const float4*>(p.pos_embed + pos * p.dim)[v];
    if (q >= 0) {
      const float4 qe = reinterpret_cast<const float4*>(p.quant_embed + q * p.dim)[v];
      base.x += qe.x; base.y += qe.y; base.z += qe.z; base.w += qe.w;
    }
    for (int b = blockIdx.y; b < p.batch; b += gridDim.y) {
      const int64_t tok = p.tokens[(int64_t)b * p.seqlen + l];
      const float4 te = reinterpret_cast<const float4*>(p.token_embed + tok * p.dim)[v];
      // the reference adds tok + pos + quant in this order (mamba_decoder.py:171); fp32 addition is commutative
      // but not associative: (tok + pos) + quant
      float4 o;
      if (q >= 0) {
        const float4 pe = reinterpret_cast<const float4*>(p.pos_embed + pos * p.dim)[v];
        const float4 qe = reinterpret_cast<const float4*>(p.quant_embed + q * p.dim)[v];
        o = make_float4((te.x + pe.x) + qe.x, (te.y + pe.y) + qe.y, (te.z + pe.z) + qe.z, (te.w + pe.w) + qe.w);
      } else {
        o = make_float4(te.x + base.x, te.y + base.y, te.z + base.z, te.w + base.w);
      }
      reinterpret_cast<float4*>(p.x + ((int64_t)b * p.seqlen + l) * p.dim)[v] = o;
    }
  }
}

// d token_embed[tok] += dx (one RED per element), d pos_embed[pos[l]] / d quant_embed[q[l]] += sum_b dx[b, l]
// (summed over the batch in registers first: one RED per position and column)
__global__ void __launch_bounds__(128)
embed_sum_bwd_kernel(const mtts_embed_sum_params p) {
  const int l = blockIdx.x;
  const int64_t pos = p.pos_ids[l];
  const int64_t q = p.quant_ids ? p.quant_ids[l] : -1;
  const int nvec = p.dim / 4;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = blockIdx.y; b < p.batch; b += gridDim.y) {
      const float4 g = reinterpret_cast<const float4*>(p.x + ((int64_t)b * p.seqlen + l) * p.dim)[v];
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      if (p.token_embed) {
        const int64_t tok = p.tokens[(int64_t)b * p.seqlen + l];
        float* d = p.token_embed + tok * p.dim + 4 * v;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                     : "memory");
      }
    }
    if (p.pos_embed) {
      float* d = p.pos_embed + pos * p.dim + 4 * v;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(acc.x), "f"(acc.y), "f"(acc.z),
                   "f"(acc.w)
                   : "memory");
    }
    if (p.quant_embed && q >= 0) {
      float* d = p.quant_embed + q * p.dim + 4 * v;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(acc.x), "f"(acc.y), "f"(acc.z),
                   "f"(acc.w)
                   : "memory");
    }
  }
}

// ---- cross entropy -----------------------------------------------------------------------------------------
// warp = one row of logits; two passes over the row held in L1/L2 (max + sum of exp, then the gradient)
template <typename T>
__global__ void __launch_bounds__(256)
ce_loss_kernel(const mtts_ce_loss_params p) {
  constexpr int VE = Io<T>::kVecElems;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= p.rows) return;
  const T* x = reinterpret_cast<const T*>(p.logits) + row * p.ld;
  T* dx = p.dlogits ? reinterpret_cast<T*>(p.dlogits) + row * p.ld : nullptr;
  const int64_t tgt = p.targets[row];
  const bool ignored = tgt == p.ignore_index;
  const int nvec = p.vocab / VE;
  if (ignored) {
    if (dx) {
      for (int v = lane; v < nvec; v += 32) *reinterpret_cast<uint4*>(dx + v * VE) = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  float mx = -INFINITY;
  for (int v = lane; v < nvec; v += 32) {
    float f[VE];
    Io<T>::unpack(*reinterpret_cast<const uint4*>(x + v * VE), f);
#pragma unroll
    for (int i = 0; i < VE; ++i) mx = fmaxf(mx, f[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[VE];
    Io<T>::unpack(*reinterpret_cast<const uint4*>(x + v * VE), f);
#pragma unroll
    for (int i = 0; i < VE; ++i) sum += ex2f((f[i] - mx) * kLog2e);
  }
  sum = warp_sum(sum);
  const float lse = mx + lg2f(sum) * kLn2;
  const float xt = Io<T>::to_f(x[tgt]);
  const float inv_n = 1.f / fmaxf(*p.n_valid, 1.f);
  if (lane == 0) {
    atomicAdd(p.loss_sum, lse - xt);
    if (p.row_loss) p.row_loss[row] = lse - xt;
  }
  if (dx) {
    const float inv = inv_n * p.grad_scale / sum;
    for (int v = lane; v < nvec; v += 32) {
      float f[VE];
      Io<T>::unpack(*reinterpret_cast<const uint4*>(x + v * VE), f);
#pragma unroll
      for (int i = 0; i < VE; ++i) {
        float g = ex2f((f[i] - mx) * kLog2e) * inv;
        if (v * VE + i == tgt) g -= inv_n * p.grad_scale;
        f[i] = g;
      }
      *reinterpret_cast<uint4*>(dx + v * VE) = Io<T>::pack(f);
    }
  }
}

// ---- clip + Adam over a table of tensors -------------------------------------------------------------------
// chunk table: chunk c covers elements [start, start + count) of tensor t (kAdamChunk elements at most)
constexpr int kAdamChunk = 8192;

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const mtts_adam_tensor* __restrict__ tensors, const int2* __restrict__ chunks, float* __restrict__ out) {
  const int2 ck = chunks[blockIdx.x];
  const mtts_adam_tensor t = tensors[ck.x];
  const int64_t start = (int64_t)ck.y * kAdamChunk;
  const int64_t end = start + kAdamChunk < t.numel ? start + kAdamChunk : t.numel;
  float s = 0.f;
  for (int64_t i = start + threadIdx.x; i < end; i += blockDim.x) {
    const float g = t.grad[i];
    s = fmaf(g, g, s);
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    atomicAdd(out, tot);
  }
}

__global__ void __launch_bounds__(256)
adam_step_kernel(const mtts_adam_tensor* __restrict__ tensors, const int2* __restrict__ chunks,
                 const mtts_adam_params p) {
  const int2 ck = chunks[blockIdx.x];
  const mtts_adam_tensor t = tensors[ck.x];
  const int64_t start = (int64_t)ck.y * kAdamChunk;
  const int64_t end = start + kAdamChunk < t.numel ? start + kAdamChunk : t.numel;
  // clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6)); max_norm <= 0 disables clipping
  float coef = 1.f;
  if (p.max_norm > 0.f) coef = fminf(1.f, p.max_norm / (sqrtf(*p.grad_sumsq) + 1e-6f));
  for (int64_t i = start + threadIdx.x; i < end; i += blockDim.x) {
    const float g = t.grad[i] * coef;
    const float m = fmaf(p.beta1, t.exp_avg[i], (1.f - p.beta1) * g);          // torch: lerp(exp_avg, grad, 1 - beta1)
    const float v = fmaf(p.beta2, t.exp_avg_sq[i], (1.f - p.beta2) * g * g);
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
    const float denom = sqrtf(v) / p.bias_correction2_sqrt + p.eps;
    t.param[i] -= p.step_size * (m / denom);
  }
}

}  // namespace
}  // namespace mtts

extern "C" int mtts_embed_sum_fwd(const mtts_embed_sum_params* p, mtts_stream_t stream) {
  if (!p || !p->tokens || !p->pos_ids || !p->token_embed || !p->pos_embed || !p->x) return MTTS_ERR_NULL;
  if (p->quant_ids && !p->quant_embed) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->seqlen < 0 || p->dim < 4 || p->dim % 4) return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->seqlen == 0) return MTTS_OK;
  const dim3 grid(p->seqlen, p->batch < 4 ? p->batch : 4);
  mtts::embed_sum_fwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return mtts::launch_status();
}

extern "C" int mtts_embed_sum_bwd(const mtts_embed_sum_params* p, mtts_stream_t stream) {
  if (!p || !p->tokens || !p->pos_ids || !p->x) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->seqlen < 0 || p->dim < 4 || p->dim % 4) return MTTS_ERR_SHAPE;
  if (p->batch == 0 || p->seqlen == 0) return MTTS_OK;
  const dim3 grid(p->seqlen, p->batch < 4 ? p->batch : 4);
  mtts::embed_sum_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(*p);
  return mtts::launch_status();
}

extern "C" int mtts_ce_loss(const mtts_ce_loss_params* p, mtts_stream_t stream) {
  if (!p || !p->logits || !p->targets || !p->loss_sum || !p->n_valid) return MTTS_ERR_NULL;
  if (p->rows < 0 || p->vocab < 1) return MTTS_ERR_SHAPE;
  const int ve = p->io_dtype == MTTS_BF16 ? 8 : 4;
  if (p->vocab % ve || p->ld % ve || !mtts::aligned16(p->logits) || (p->dlogits && !mtts::aligned16(p->dlogits)))
    return MTTS_ERR_ALIGN;
  if (p->rows == 0) return MTTS_OK;
  const int grid = (int)((p->rows + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->io_dtype) {
    case MTTS_F32: mtts::ce_loss_kernel<float><<<grid, 256, 0, s>>>(*p); break;
    case MTTS_BF16: mtts::ce_loss_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(*p); break;
    default: return MTTS_ERR_DTYPE;
  }
  return mtts::launch_status();
}

extern "C" int mtts_grad_sumsq(const mtts_adam_params* p, mtts_stream_t stream) {
  if (!p || !p->tensors || !p->chunks || !p->grad_sumsq) return MTTS_ERR_NULL;
  if (p->num_chunks < 0) return MTTS_ERR_SHAPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(p->grad_sumsq, 0, sizeof(float), s);
  if (e != cudaSuccess) return -static_cast<int>(e);
  if (p->num_chunks == 0) return MTTS_OK;
  mtts::grad_sumsq_kernel<<<p->num_chunks, 256, 0, s>>>(p->tensors, reinterpret_cast<const int2*>(p->chunks),
                                                        p->grad_sumsq);
  return mtts::launch_status();
}

extern "C" int mtts_adam_step(const mtts_adam_params* p, mtts_stream_t stream) {
  if (!p || !p->tensors || !p->chunks) return MTTS_ERR_NULL;
  if (p->max_norm > 0.f && !p->grad_sumsq) return MTTS_ERR_NULL;
  if (p->num_chunks < 0) return MTTS_ERR_SHAPE;
  if (p->num_chunks == 0) return MTTS_OK;
  mtts::adam_step_kernel<<<p->num_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p->tensors, reinterpret_cast<const int2*>(p->chunks), *p);
  return mtts::launch_status();
}

extern "C" int mtts_adam_chunk_elems(void) { return mtts::kAdamChunk; }
