// decode_step's dense contractions for a handful of rows (batch <= 64): one launch computes
//     out = act( LNopt(x + delta) @ W^T + bias )
// i.e. the residual add, the LayerNorm (+FiLM) that mamba_decoder.py:59,67,81-86 applies before each
// projection, the projection itself and (for ff[0]) the exact-erf GELU of :39-43.  With M <= 64 the GEMM
// is weight-streaming bound: every CTA owns 16 output columns x one 512-wide K range (split-K over a
// thread-block cluster, partials summed through DSMEM), keeps the M rows of its K range in shared
// memory (bf16, cp.async) and streams its weight slice exactly once.  Tensor cores via mma.sync m16n8k16
// (a 128-row tcgen05 tile would be >= 50 % padding at these sizes).  bf16 only; fp32 decode keeps the
// library GEMM so that greedy ids stay bit-identical to the fp32 oracle path.
#include <cooperative_groups.h>

#include "common.cuh"

namespace mtts {

constexpr int kSkThreads = 256;  // 8 warps = 4 row tiles of 16 x 2 column tiles of 8
constexpr int kSkNT = 16;        // output columns per CTA
constexpr int kSkKC = 512;       // K chunk resident in shared memory
constexpr int kSkPad = 8;        // bf16 elements of row padding (bank-conflict-free fragment loads)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4],
                                               const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
}

__device__ __forceinline__ void sk_cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}

// Grid: (ceil(N/16) * ksplit) CTAs, clusters of `ksplit` CTAs along K (ksplit = 1 when ln_mode).
// CTA rank r of a cluster owns the K range [r*kc, (r+1)*kc); partial products are summed by rank 0
// through distributed shared memory.
// kG = 0: plain mode (A copied as is); kG > 0: LN mode with K == kG * 128.
template <int kG>
__global__ void __launch_bounds__(kSkThreads)
skinny_linear_kernel(const mtts_skinny_linear_params p, const int kc, const int ksplit) {
  using bf16 = __nv_bfloat16;
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char sk_smem[];
  const int M = p.m, N = p.n, K = p.k;
  const int lda = kc + kSkPad;
  bf16* As = reinterpret_cast<bf16*>(sk_smem);                     // [64][lda]
  bf16* Ws = As + 64 * lda;                                        // [kSkNT][lda]
  float* part = reinterpret_cast<float*>(Ws + kSkNT * lda);        // [64][kSkNT] split-K partial
  const int krank = (ksplit > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const int n0 = (blockIdx.x / ksplit) * kSkNT;
  const int k0 = krank * kc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int vec_per_row = kc / 8;

  // ---- W slice and (plain mode) A chunk: all copies in flight at once, no staging registers ----
  {
    const bf16* w = reinterpret_cast<const bf16*>(p.w);
    for (int idx = threadIdx.x; idx < kSkNT * vec_per_row; idx += kSkThreads) {
      const int r = idx / vec_per_row, v = idx - r * vec_per_row;
      bf16* dst = Ws + r * lda + v * 8;
      if (n0 + r < N) sk_cp_async16(dst, w + (int64_t)(n0 + r) * K + k0 + v * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    if constexpr (kG == 0) {
      const bf16* a = reinterpret_cast<const bf16*>(p.a);
      for (int idx = threadIdx.x; idx < 64 * vec_per_row; idx += kSkThreads) {
        const int r = idx / vec_per_row, v = idx - r * vec_per_row;
        bf16* dst = As + r * lda + v * 8;
        if (r < M) sk_cp_async16(dst, a + (int64_t)r * K + k0 + v * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  if constexpr (kG > 0) {
    // LN mode, K == kG * 128: one warp per row, the whole row in registers (one global read per
    // element), two rows in flight per warp.  Every load below is unconditional and batched -- guards
    // inside the unrolled loops would split them into dependent basic blocks and serialise the
    // round trips.
    constexpr int kWarpsSk = kSkThreads / 32;
    for (int rb = warp; rb < 64; rb += 2 * kWarpsSk) {
      float v[2][kG][4];
      bool live[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = rb + u * kWarpsSk;
        live[u] = r < M;
        const float* x = p.x + (int64_t)(live[u] ? r : 0) * K;
#pragma unroll
        for (int gI = 0; gI < kG; ++gI) load4<float>(x + (gI * 32 + lane) * 4, v[u][gI]);
      }
      if (p.delta) {
        float d[2][kG][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int r = rb + u * kWarpsSk;
          const bf16* dl = reinterpret_cast<const bf16*>(p.delta) + (int64_t)(live[u] ? r : 0) * K;
#pragma unroll
          for (int gI = 0; gI < kG; ++gI) load4<bf16>(dl + (gI * 32 + lane) * 4, d[u][gI]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int gI = 0; gI < kG; ++gI)
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) v[u][gI][j2] += d[u][gI][j2];
      }
      float w[kG][4], bb[kG][4];
#pragma unroll
      for (int gI = 0; gI < kG; ++gI) {
        load4<float>(p.ln_weight + (gI * 32 + lane) * 4, w[gI]);
        load4<float>(p.ln_bias + (gI * 32 + lane) * 4, bb[gI]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = rb + u * kWarpsSk;
        bf16* arow = As + r * lda;
        float sum = 0.f;
#pragma unroll
        for (int gI = 0; gI < kG; ++gI)
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) sum += v[u][gI][j2];
        const float mean = warp_sum(sum) / (float)K;
        float q = 0.f;
#pragma unroll
        for (int gI = 0; gI < kG; ++gI)
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) q = fmaf(v[u][gI][j2] - mean, v[u][gI][j2] - mean, q);
        const float rstd = rsqrtf(warp_sum(q) / (float)K + p.eps);
        if (live[u] && p.x_out && blockIdx.x == 0) {
#pragma unroll
          for (int gI = 0; gI < kG; ++gI)
            store4<float>(p.x_out + (int64_t)r * K + (gI * 32 + lane) * 4, v[u][gI]);
        }
        float o[kG][4];
#pragma unroll
        for (int gI = 0; gI < kG; ++gI)
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2)
            o[gI][j2] = live[u] ? fmaf((v[u][gI][j2] - mean) * rstd, w[gI][j2], bb[gI][j2]) : 0.f;
        if (p.film_gamma && live[u]) {
          float gm[kG][4], bt[kG][4];
#pragma unroll
          for (int gI = 0; gI < kG; ++gI) {
            load4<float>(p.film_gamma + (int64_t)r * K + (gI * 32 + lane) * 4, gm[gI]);
            load4<float>(p.film_beta + (int64_t)r * K + (gI * 32 + lane) * 4, bt[gI]);
          }
#pragma unroll
          for (int gI = 0; gI < kG; ++gI)
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) o[gI][j2] = fmaf(gm[gI][j2], o[gI][j2], bt[gI][j2]);
        }
#pragma unroll
        for (int gI = 0; gI < kG; ++gI) store4<bf16>(arow + (gI * 32 + lane) * 4, o[gI]);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- tensor-core product: warp = 16 rows x 8 columns over this CTA's K range ----
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int mt = warp & 3, nt = warp >> 2;
  if (mt * 16 < M) {
    const bf16* Ar0 = As + (mt * 16 + g) * lda + t4 * 2;
    const bf16* Ar1 = Ar0 + 8 * lda;
    const bf16* Wr0 = Ws + (nt * 8 + g) * lda + t4 * 2;   // column n0 + nt*8 + g
#pragma unroll 8
    for (int ks = 0; ks < kc; ks += 16) {
      uint32_t af[4], b0[2];
      af[0] = *reinterpret_cast<const uint32_t*>(Ar0 + ks);
      af[1] = *reinterpret_cast<const uint32_t*>(Ar1 + ks);
      af[2] = *reinterpret_cast<const uint32_t*>(Ar0 + ks + 8);
      af[3] = *reinterpret_cast<const uint32_t*>(Ar1 + ks + 8);
      b0[0] = *reinterpret_cast<const uint32_t*>(Wr0 + ks);
      b0[1] = *reinterpret_cast<const uint32_t*>(Wr0 + ks + 8);
      mma_bf16_16816(acc, af, b0);
    }
  }

  // ---- split-K: ranks > 0 publish their partial tile, rank 0 sums through DSMEM ----
  if (ksplit > 1) {
    cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = mt * 16 + g + half * 8, col = nt * 8 + t4 * 2;
      part[row * kSkNT + col] = acc[half * 2];
      part[row * kSkNT + col + 1] = acc[half * 2 + 1];
    }
    cluster.sync();
    if (krank == 0) {
      for (int r = 1; r < ksplit; ++r) {
        const float* peer = cluster.map_shared_rank(part, r);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int row = mt * 16 + g + half * 8, col = nt * 8 + t4 * 2;
          acc[half * 2] += peer[row * kSkNT + col];
          acc[half * 2 + 1] += peer[row * kSkNT + col + 1];
        }
      }
    }
    cluster.sync();  // peers stay resident until rank 0 has read their partials
    if (krank != 0) return;
  }

  // ---- epilogue: bias, GELU, store (accumulator element j: row g + 8*(j/2), column 2*t4 + (j&1)) ----
  if (mt * 16 < M) {
    const bf16* bias = reinterpret_cast<const bf16*>(p.bias);
    bf16* out = reinterpret_cast<bf16*>(p.out);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = mt * 16 + g + half * 8;
      const int col = n0 + nt * 8 + t4 * 2;
      if (row < M && col < N) {
        float v0 = acc[half * 2], v1 = acc[half * 2 + 1];
        if (bias) {
          v0 += __bfloat162float(bias[col]);
          if (col + 1 < N) v1 += __bfloat162float(bias[col + 1]);
        }
        if (p.gelu) {
          v0 = gelu_erf(v0);
          v1 = gelu_erf(v1);
        }
        if (col + 1 < N && (N & 1) == 0) {
          const __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
          *reinterpret_cast<__nv_bfloat162*>(out + (int64_t)row * N + col) = pk;
        } else {
          out[(int64_t)row * N + col] = __float2bfloat16_rn(v0);
          if (col + 1 < N) out[(int64_t)row * N + col + 1] = __float2bfloat16_rn(v1);
        }
      }
    }
  }
}

}  // namespace mtts

extern "C" int mtts_skinny_linear(const mtts_skinny_linear_params* p, mtts_stream_t stream) {
  if (!p || !p->w || !p->out) return MTTS_ERR_NULL;
  if (p->ln_mode ? (!p->x || !p->ln_weight || !p->ln_bias) : !p->a) return MTTS_ERR_NULL;
  if ((p->film_gamma == nullptr) != (p->film_beta == nullptr)) return MTTS_ERR_NULL;
  if (p->io_dtype != MTTS_BF16) return MTTS_ERR_DTYPE;
  if (p->m < 0 || p->m > 64 || p->n < 1 || p->k < 16 || p->k % 16 != 0) return MTTS_ERR_SHAPE;
  // LN mode keeps whole rows in registers / shared memory: K in {128, 256, 512, 1024}
  if (p->ln_mode && !(p->k == 128 || p->k == 256 || p->k == 512 || p->k == 1024)) return MTTS_ERR_SHAPE;
  if (!p->ln_mode && p->k > mtts::kSkKC && p->k % mtts::kSkKC != 0) return MTTS_ERR_SHAPE;
  if (p->ln_mode && p->x_out == p->x && p->delta) return MTTS_ERR_UNSUPPORTED;  // CTAs race on x
  if (p->m == 0) return MTTS_OK;
  if (!mtts::aligned16(p->w) || (p->a && !mtts::aligned16(p->a))) return MTTS_ERR_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // K split over a cluster (plain mode): 512-wide chunks, at most 8 CTAs per cluster
  int ksplit = 1, kc = p->k;
  if (!p->ln_mode && p->k > mtts::kSkKC) {
    ksplit = p->k / mtts::kSkKC;
    kc = mtts::kSkKC;
    if (ksplit > 8) return MTTS_ERR_SHAPE;
  }
  const size_t smem = sizeof(__nv_bfloat16) * (size_t)(64 + mtts::kSkNT) * (kc + mtts::kSkPad) +
                      sizeof(float) * 64 * mtts::kSkNT;
  void (*kern)(const mtts_skinny_linear_params, const int, const int) = mtts::skinny_linear_kernel<0>;
  if (p->ln_mode) {
    switch (p->k) {
      case 128: kern = mtts::skinny_linear_kernel<1>; break;
      case 256: kern = mtts::skinny_linear_kernel<2>; break;
      case 512: kern = mtts::skinny_linear_kernel<4>; break;
      default: kern = mtts::skinny_linear_kernel<8>; break;
    }
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((p->n + mtts::kSkNT - 1) / mtts::kSkNT) * ksplit);
  cfg.blockDim = dim3(mtts::kSkThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ksplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, *p, kc, ksplit);
  if (e != cudaSuccess) return -static_cast<int>(e);
  return mtts::launch_status();
}
