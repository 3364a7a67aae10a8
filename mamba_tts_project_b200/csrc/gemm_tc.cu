// Dense contraction of the decoder's FFN on the 5th-generation tensor cores (tcgen05 + TMEM + TMA):
//     out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] )          (nn.Linear layout: both K-major)
// i.e. mamba_decoder.py:39-43,88 `ff[0]` + `nn.GELU()` (gelu = 1) or `ff[2]` (gelu = 0), bf16 operands,
// fp32 accumulation in tensor memory, exact-erf GELU in the epilogue.
//
// One 128 x 128 output tile per CTA, K walked in 64-element (128-byte, swizzle-128B) blocks:
//   warp 4  (1 lane)  TMA producer: cp.async.bulk.tensor A/B tiles -> 4-stage shared-memory ring
//   warp 5  (1 lane)  MMA issuer:   tcgen05.mma cta_group::1 kind::f16, M128 N128 K16, D in TMEM;
//                                   tcgen05.commit frees the stage / signals the epilogue
//   warps 0-3         epilogue:     tcgen05.ld 32x32b (warp w owns TMEM lanes 32w..32w+31 = rows),
//                                   + bias, GELU, bf16, 16-byte global stores
// Two CTAs fit per SM (TMEM 2 x 128 columns, smem 2 x ~97 KB with 3 stages), so one CTA's epilogue
// overlaps the other's main loop.
#include <cuda.h>

#include "common.cuh"

namespace mtts {

constexpr int kGM = 128, kGN = 128, kGK = 64;   // CTA tile; kGK * 2 B = one 128-byte swizzle row
constexpr int kGStages = 3;
constexpr int kGThreads = 192;
constexpr uint32_t kGTileBytes = kGM * kGK * 2;  // 16 KiB per operand tile
constexpr uint32_t kTmemCols = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// K-major, swizzle-128B shared-memory matrix descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kGN >> 3) << 17) |
                            ((uint32_t)(kGM >> 4) << 24);

__device__ __forceinline__ float gelu_erf_tc(float v) {
  return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
}

__global__ void __launch_bounds__(kGThreads)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const mtts_gemm_bf16_params p) {
  extern __shared__ unsigned char gsm_raw[];
  // swizzle-128B tiles need 1024-byte alignment
  unsigned char* gsm = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(gsm_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sA = gsm;
  unsigned char* sB = gsm + kGStages * kGTileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gsm + 2 * kGStages * kGTileBytes);
  uint64_t* empty_bar = full_bar + kGStages;
  uint64_t* tmem_full = empty_bar + kGStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kGN, m0 = blockIdx.y * kGM;
  const int num_kb = (p.k + kGK - 1) / kGK;

  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < kGStages; ++s) {
        mbar_init(full_bar + s, 1);
        mbar_init(empty_bar + s, 1);
      }
      mbar_init(tmem_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kGStages;
        const uint32_t ph = (kb / kGStages) & 1;
        mbar_wait(empty_bar + s, ph ^ 1);  // fresh barrier: the "previous" phase counts as complete
        mbar_expect_tx(full_bar + s, 2 * kGTileBytes);
        tma_load_2d(sA + s * kGTileBytes, &map_a, full_bar + s, kb * kGK, m0);
        tma_load_2d(sB + s * kGTileBytes, &map_w, full_bar + s, kb * kGK, n0);
      }
    }
  } else if (warp == 5) {
    // ===== MMA issuer (one thread drives the tensor core for the whole CTA) =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kGStages;
        const uint32_t ph = (kb / kGStages) & 1;
        mbar_wait(full_bar + s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = umma_desc_sw128(smem_u32(sA + s * kGTileBytes));
        const uint64_t db = umma_desc_sw128(smem_u32(sB + s * kGTileBytes));
#pragma unroll
        for (int k = 0; k < kGK / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the swizzled row: +2 in the (>>4) start-address field
          const uint32_t acc = (kb | k) ? 1u : 0u;
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
              "}\n" ::"r"(tmem_base),
              "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(kIdesc), "r"(acc)
              : "memory");
        }
        // frees the smem stage once the MMAs that read it have retired
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(empty_bar + s))
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(tmem_full))
                   : "memory");
    }
  } else {
    // ===== epilogue: TMEM -> registers -> bias / GELU -> bf16 -> global =====
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + warp * 32 + lane;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)row * p.ldo;
    __nv_bfloat16* pre = p.pre_out ? reinterpret_cast<__nv_bfloat16*>(p.pre_out) + (int64_t)row * p.ldo
                                   : nullptr;
#pragma unroll 1
    for (int cchunk = 0; cchunk < kGN / 32; ++cchunk) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cchunk * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
            "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
            "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
            "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int col0 = n0 + cchunk * 32;
      if (row < p.m) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float v[8], pv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float x = __uint_as_float(r[j + q]);
            if (p.bias && col0 + j + q < p.n) x += p.bias[col0 + j + q];
            pv[q] = x;
            v[q] = p.gelu ? gelu_erf_tc(x) : x;
          }
          if (col0 + j + 8 <= p.n) {
            *reinterpret_cast<uint4*>(out + col0 + j) = Io<__nv_bfloat16>::pack(v);
            if (pre) *reinterpret_cast<uint4*>(pre + col0 + j) = Io<__nv_bfloat16>::pack(pv);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (col0 + j + q < p.n) {
                out[col0 + j + q] = __float2bfloat16_rn(v[q]);
                if (pre) pre[col0 + j + q] = __float2bfloat16_rn(pv[q]);
              }
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 2-D bf16 row-major (rows, cols) with leading dimension ld -> tiles of (box_rows, 64) swizzle-128B
static bool make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld,
                     int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kGK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace mtts

extern "C" int mtts_gemm_bf16(const mtts_gemm_bf16_params* p, mtts_stream_t stream) {
  if (!p || !p->a || !p->w || !p->out) return MTTS_ERR_NULL;
  if (p->m < 0 || p->n < 1 || p->k < 8 || p->k % 8 != 0 || p->n % 8 != 0) return MTTS_ERR_SHAPE;
  if (p->lda % 8 != 0 || p->ldw % 8 != 0 || p->ldo % 8 != 0 || p->lda < p->k || p->ldw < p->k ||
      p->ldo < p->n)
    return MTTS_ERR_ALIGN;
  if (!mtts::aligned16(p->a) || !mtts::aligned16(p->w) || !mtts::aligned16(p->out) ||
      (p->pre_out && !mtts::aligned16(p->pre_out)))
    return MTTS_ERR_ALIGN;
  if (p->m == 0) return MTTS_OK;
  CUtensorMap map_a, map_w;
  if (!mtts::make_map(&map_a, p->a, p->m, p->k, p->lda, mtts::kGM) ||
      !mtts::make_map(&map_w, p->w, p->n, p->k, p->ldw, mtts::kGN))
    return MTTS_ERR_UNSUPPORTED;
  const size_t smem = 2 * mtts::kGStages * mtts::kGTileBytes + (2 * mtts::kGStages + 1) * 8 + 16 + 1024;
  cudaError_t e = cudaFuncSetAttribute(mtts::gemm_bf16_tc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return -static_cast<int>(e);
  const dim3 grid((p->n + mtts::kGN - 1) / mtts::kGN, (p->m + mtts::kGM - 1) / mtts::kGM);
  mtts::gemm_bf16_tc_kernel<<<grid, mtts::kGThreads, smem, static_cast<cudaStream_t>(stream)>>>(map_a, map_w, *p);
  return mtts::launch_status();
}
