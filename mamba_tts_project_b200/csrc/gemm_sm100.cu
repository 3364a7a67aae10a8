// Dense contractions of the decoder on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//     C[b][m, n] = epilogue( sum_k A[b][m, k] * B[b][n, k] )        bf16 operands, fp32 accumulation in TMEM
//
// One kernel family serves every GEMM of the teacher-forced path (mamba_decoder.py:29,39-43,61,72-77,88,118):
// the Mamba projections in their channel-major layouts, q / kv / o projections, QK^T with a fused masked
// softmax, PV, the FFN with bias + GELU fused, the head -- forward, data gradients and weight gradients.
// Either operand may be K-major (contraction index contiguous: nn.Linear weights, token-major activations)
// or MN-major (row index contiguous: channel-major activations, transposed weights, every weight gradient):
// the layout goes into the TMA box and the UMMA descriptors, nothing is ever transposed in memory.
//
// Persistent, warp-specialised CTA (one per SM), tile 128 x BN (BN = 256 / 128 / 64), K in 64-element blocks:
//   warp 0   (1 lane)  TMA producer: cp.async.bulk.tensor.4d (matrix x inner batch x outer batch) into a
//                      ring of kStages shared-memory stages (swizzle-128B), mbarrier complete_tx
//   warp 1   (1 lane)  MMA issuer: tcgen05.mma cta_group::1 kind::f16 M128 x N BN x K16, accumulators in
//                      TMEM, DOUBLE BUFFERED (2 x BN columns): tile i+1 is accumulated while tile i drains;
//                      tcgen05.commit releases smem stages / publishes an accumulator
//   warps 2-9          epilogue: tcgen05.ld 32x32b (two warps per 32-lane group, splitting the columns)
//                      -> bias / GELU / GELU' / masked softmax / softmax backward -> bf16 or fp32 -> global
// Work items (batch, k-slice, m-block, n-block) are dealt round-robin with the n-block fastest, so CTAs running
// at the same time share A panels through L2.  Weight gradients contract over batch x time: the k-loop walks
// all batches (k_batches) and is split across CTAs (split_k) with fp32 vector REDs into the zeroed output.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mtts {
namespace g100 {

constexpr int BM = 128, BK = 64;
enum : int { EPI_STORE = 0, EPI_GELU = 1, EPI_GELU_BWD = 2, EPI_SOFTMAX = 3, EPI_DSOFTMAX = 4, EPI_MUL_AUX = 5 };
// Epilogue warps: EPW / 4 per 32-lane group of TMEM, each draining its share of the tile's columns.  The epilogue is
// a chain of dependent fixed-latency steps (tcgen05.ld -> math -> staging -> store), so what hides its latency is
// warps: the row-wise epilogues (one k-block of MMA per tile, so the epilogue IS the kernel) and the two-output
// GELU epilogue run 16 (softmax 82 -> 56 us, dsoftmax 103 -> 86 us at C2), the others 8 -- there the extra 32 KB
// of store staging would cost a pipeline stage (-5 % on the K >= 512 GEMMs).
template <int EPI>
__host__ __device__ constexpr int epi_warps() { return (EPI == EPI_SOFTMAX || EPI == EPI_DSOFTMAX || EPI == EPI_GELU) ? 16 : 8; }
constexpr uint32_t kABytes = BM * BK * 2;       // 16 KiB
constexpr float kInvSqrt2 = 0.70710678118654752f;
constexpr float kInvSqrt2Pi = 0.39894228040143268f;

// CG = CTAs per MMA (tcgen05 cta_group): with CG = 2 a pair of CTAs on neighbouring SMs computes one 256 x BN tile,
// each CTA staging its own 128 rows of A and HALF of B -- L2 -> SM traffic per flop drops by a third against two
// independent 128 x BN tiles and the shared-memory read rate of the MMA by the same amount.
template <int BN, int CG, int EPW>
struct Tile {
  static constexpr uint32_t kBBytes = BN / CG * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr size_t kStagingBytes = (size_t)EPW * 4096;
  static constexpr size_t kMiscBytes = 8 * (2 * 8 + 4) + 16;   // barriers of up to 8 stages + 4 + the TMEM slot
  // as many stages as fit beside the store staging (227 KB per CTA, minus the alignment slack), at most 8
  static constexpr int kFit = (int)((232448 - 1024 - kStagingBytes - kMiscBytes) / kStageBytes);
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr uint32_t kTmemCols = 2 * BN;
  // stages + per-epilogue-warp store staging (32 rows x 128 B; its first 256 B double as the warp's slot of the
  // row-statistics exchange) + barriers (full, empty, tmem full/empty) + tmem slot; the dynamic shared window is
  // declared 1024-byte aligned (swizzle-128B atoms)
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kStagingBytes + kMiscBytes;
};

struct Args {
  int m, n;
  int nb_m, nb_n, split_k, kb_total, kb_per_split, kpb, k_batches;
  int batch_inner, batches;
  int a_bi, a_bo, b_bi, b_bo;      // 1 when the operand has that batch dimension, 0 when it is broadcast
  int out_f32, accumulate, atomic, vec_ok, debug, aux_gelu_grad;
  void* out;
  long long ldc, c_bo, c_bi;
  const float* bias_n;
  const float* bias_m;
  void* aux;
  long long ld_aux, aux_bo, aux_bi;
  const unsigned char* mask;
  long long mask_bo;
  float scale;
  float* row_stat;
};

// ---- epilogue math ----------------------------------------------------------------------------------------
// GELU on the bf16 tensor-core path.  The epilogue has one MUFU slot per element before it, not the main loop,
// bounds a K = 512 tile (128 x 256 elements against 4096 tensor cycles), so Phi(x) = 0.5 (1 + erf(x / sqrt 2)) is
// evaluated as 0.5 (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) with MUFU.TANH: |gelu - x Phi(x)| <= 5e-4 absolute, below
// the rounding of the bf16 value it is stored as for every |gelu| >= 0.13.  (The fp32 path keeps exact-erf GELU:
// ffn_glue.cu.)  The backward differentiates the same form.
constexpr float kGeluC0 = 0.7978845608028654f;    // sqrt(2 / pi)
constexpr float kGeluC1 = 0.035677408136300125f;  // 0.044715 sqrt(2 / pi)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tc(float x) {
  const float t = tanh_approx(x * fmaf(kGeluC1, x * x, kGeluC0));
  return fmaf(0.5f * x, t, 0.5f * x);
}
__device__ __forceinline__ float gelu_grad_tc(float x) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(kGeluC1, x2, kGeluC0));
  const float dy = fmaf(3.f * kGeluC1, x2, kGeluC0);          // d/dx of the tanh argument
  return fmaf(0.5f * x * dy, fmaf(-t, t, 1.f), fmaf(0.5f, t, 0.5f));
}

// value and derivative from ONE tanh (the forward can leave gelu'(pre) for the backward instead of pre: the
// data-gradient GEMM's epilogue is then a single multiply per element instead of a second tanh evaluation)
__device__ __forceinline__ void gelu_both_tc(float x, float& y, float& dy) {
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(kGeluC1, x2, kGeluC0));
  const float h = fmaf(0.5f, t, 0.5f);
  y = x * h;
  dy = fmaf(0.5f * x * fmaf(3.f * kGeluC1, x2, kGeluC0), fmaf(-t, t, 1.f), h);
}

// Store 32 consecutive columns v[0..32) of one row starting at column `col` (multiple of 32) of a row of n
// columns.  vec: the row base is 16-byte aligned and n % 8 == 0.
template <bool kF32>
__device__ __forceinline__ void store_row32(void* row_base, int col, int n, const float* v, bool vec, bool accumulate,
                                            bool atomic) {
  if constexpr (kF32) {
    float* o = reinterpret_cast<float*>(row_base) + col;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (vec && col + j + 4 <= n) {
        if (atomic) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(v[j]), "f"(v[j + 1]),
                       "f"(v[j + 2]), "f"(v[j + 3])
                       : "memory");
        } else {
          float4 w = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(o + j);
            w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
          }
          *reinterpret_cast<float4*>(o + j) = w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (col + j + q < n) {
            if (atomic) atomicAdd(o + j + q, v[j + q]);
            else o[j + q] = accumulate ? o[j + q] + v[j + q] : v[j + q];
          }
        }
      }
    }
  } else {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(row_base) + col;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (vec && col + j + 8 <= n) {
        float w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = v[j + q];
        if (accumulate) {
          float old[8];
          unpack_bf16x8(*reinterpret_cast<const uint4*>(o + j), old);
#pragma unroll
          for (int q = 0; q < 8; ++q) w[q] += old[q];
        }
        *reinterpret_cast<uint4*>(o + j) =
            make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (col + j + q < n) {
            const float x = accumulate ? v[j + q] + __bfloat162float(o[j + q]) : v[j + q];
            o[j + q] = __float2bfloat16_rn(x);
          }
        }
      }
    }
  }
}

// Warp-collective store of a 32-row x 32-column block (lane = row, v = its 32 columns starting at column `col`, a
// multiple of 32): the rows go through a swizzled shared-memory staging tile and leave as 16-byte vectors with the
// lanes of a quarter-warp on consecutive addresses of ONE row (full 32-byte sectors per request; per-lane row
// strides would touch 32 half-written sectors per instruction).  Needs 16-byte aligned rows and n % 8 == 0.
template <bool kF32>
__device__ __forceinline__ void warp_store32(unsigned char* stage, int lane, unsigned char* warp_base,
                                             long long ld_bytes, int rows_valid, int col, int n, const float* v,
                                             bool accumulate, bool atomic, const uint4* old_bf16 = nullptr) {
  constexpr int SLOTS = kF32 ? 8 : 4;      // 16-byte slots per row
  constexpr int EL = kF32 ? 4 : 8;         // elements per slot
  constexpr int PITCH = SLOTS * 16;
  {
    const int key = kF32 ? (lane & 7) : ((lane >> 1) & 3);
    unsigned char* srow = stage + lane * PITCH;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      uint4 w;
      if constexpr (kF32) {
        w = make_uint4(__float_as_uint(v[4 * sl]), __float_as_uint(v[4 * sl + 1]), __float_as_uint(v[4 * sl + 2]),
                       __float_as_uint(v[4 * sl + 3]));
      } else {
        w = make_uint4(pack_bf16(v[8 * sl], v[8 * sl + 1]), pack_bf16(v[8 * sl + 2], v[8 * sl + 3]),
                       pack_bf16(v[8 * sl + 4], v[8 * sl + 5]), pack_bf16(v[8 * sl + 6], v[8 * sl + 7]));
      }
      *reinterpret_cast<uint4*>(srow + ((sl ^ key) << 4)) = w;
    }
  }
  __syncwarp();
#pragma unroll
  for (int pass = 0; pass < SLOTS; ++pass) {
    const int idx = pass * 32 + lane;
    const int r = idx / SLOTS, sl = idx % SLOTS;
    const int key = kF32 ? (r & 7) : ((r >> 1) & 3);
    uint4 w = *reinterpret_cast<const uint4*>(stage + r * PITCH + ((sl ^ key) << 4));
    const int c = col + sl * EL;
    if (r < rows_valid && c + EL <= n) {
      unsigned char* dst = warp_base + (long long)r * ld_bytes + (size_t)c * (kF32 ? 4 : 2);
      if constexpr (kF32) {
        if (atomic) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(w.x)),
                       "f"(__uint_as_float(w.y)), "f"(__uint_as_float(w.z)), "f"(__uint_as_float(w.w))
                       : "memory");
        } else {
          if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(dst);
            w.x = __float_as_uint(__uint_as_float(w.x) + old.x);
            w.y = __float_as_uint(__uint_as_float(w.y) + old.y);
            w.z = __float_as_uint(__uint_as_float(w.z) + old.z);
            w.w = __float_as_uint(__uint_as_float(w.w) + old.w);
          }
          *reinterpret_cast<uint4*>(dst) = w;
        }
      } else {
        if (accumulate) {
          float a[8], b[8];
          unpack_bf16x8(w, a);
          unpack_bf16x8(old_bf16 ? old_bf16[pass] : *reinterpret_cast<const uint4*>(dst), b);
          w = make_uint4(pack_bf16(a[0] + b[0], a[1] + b[1]), pack_bf16(a[2] + b[2], a[3] + b[3]),
                         pack_bf16(a[4] + b[4], a[5] + b[5]), pack_bf16(a[6] + b[6], a[7] + b[7]));
        }
        *reinterpret_cast<uint4*>(dst) = w;
      }
    }
  }
  __syncwarp();   // the staging tile is rewritten by the next block
}

// 64 bf16 columns at a time (lane = row, v = its 64 columns starting at `col`, a multiple of 64): the rows leave as
// whole 128-byte lines, eight lanes per row.  v is scaled by `mul` on the way (softmax normalisation).
__device__ __forceinline__ void warp_store64_bf16(unsigned char* stage, int lane, unsigned char* warp_base,
                                                  long long ld_bytes, int rows_valid, int col, int n, const float* v,
                                                  float mul) {
  {
    const int key = lane & 7;
    unsigned char* srow = stage + lane * 128;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
      const uint4 w = make_uint4(pack_bf16(v[8 * sl] * mul, v[8 * sl + 1] * mul),
                                 pack_bf16(v[8 * sl + 2] * mul, v[8 * sl + 3] * mul),
                                 pack_bf16(v[8 * sl + 4] * mul, v[8 * sl + 5] * mul),
                                 pack_bf16(v[8 * sl + 6] * mul, v[8 * sl + 7] * mul));
      *reinterpret_cast<uint4*>(srow + ((sl ^ key) << 4)) = w;
    }
  }
  __syncwarp();
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const int idx = pass * 32 + lane;
    const int r = idx >> 3, sl = idx & 7;
    const uint4 w = *reinterpret_cast<const uint4*>(stage + r * 128 + ((sl ^ (r & 7)) << 4));
    const int c = col + sl * 8;
    if (r < rows_valid && c + 8 <= n)
      *reinterpret_cast<uint4*>(warp_base + (long long)r * ld_bytes + (size_t)c * 2) = w;
  }
  __syncwarp();
}

// Prefetch of the bf16 values a warp_store32<false>(.., accumulate) will add to, in that function's own
// (row, slot) assignment, so that the loads are in flight before the accumulator is waited for.
__device__ __forceinline__ void warp_ldg_old32(const unsigned char* warp_base, long long ld_bytes, int rows_valid,
                                               int col, int n, int lane, uint4 (&old)[4]) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int idx = pass * 32 + lane;
    const int r = idx >> 2, c = col + (idx & 3) * 8;
    old[pass] = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid && c + 8 <= n)
      old[pass] = *reinterpret_cast<const uint4*>(warp_base + (long long)r * ld_bytes + (size_t)c * 2);
  }
}

// The reverse for a bf16 operand of the epilogue (aux), in two steps so that the global loads can be issued long
// before their data is needed (before the accumulator is even complete):
//   warp_ldg32      the warp's 32-row x 32-column block, lanes of a quarter-warp on consecutive 16-byte vectors of
//                   two rows (zeros outside the matrix) -> 4 registers of raw data per lane, any row
//   warp_own_rows   parks them in the staging tile and hands every lane its OWN row:
//                   raw[sl] = columns [col + 8 sl, col + 8 sl + 8) of row `lane`
__device__ __forceinline__ void warp_ldg32(const unsigned char* warp_base, long long ld_bytes, int rows_valid, int col,
                                           int n, int lane, uint4 (&w)[4]) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int idx = pass * 32 + lane;
    const int r = idx >> 2, c = col + (idx & 3) * 8;
    w[pass] = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid && c + 8 <= n)
      w[pass] = __ldg(reinterpret_cast<const uint4*>(warp_base + (long long)r * ld_bytes + (size_t)c * 2));
  }
}
__device__ __forceinline__ void warp_own_rows(unsigned char* stage, int lane, const uint4 (&w)[4], uint4 (&raw)[4]) {
  // (w and raw may be the same array: w is consumed before raw is written)
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int idx = pass * 32 + lane;
    const int r = idx >> 2, sl = idx & 3;
    *reinterpret_cast<uint4*>(stage + r * 64 + ((sl ^ ((r >> 1) & 3)) << 4)) = w[pass];
  }
  __syncwarp();
#pragma unroll
  for (int sl = 0; sl < 4; ++sl)
    raw[sl] = *reinterpret_cast<const uint4*>(stage + lane * 64 + ((sl ^ ((lane >> 1) & 3)) << 4));
  __syncwarp();
}
__device__ __forceinline__ void unpack32(const uint4 (&raw)[4], float* v) {
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) unpack_bf16x8(raw[sl], v + 8 * sl);
}

// 32 bf16 of a row starting at column col -> fp32 (zeros past n)
__device__ __forceinline__ void load_row32_bf16(const void* row_base, int col, int n, float* v, bool vec) {
  const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(row_base) + col;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    if (vec && col + j + 8 <= n) {
      unpack_bf16x8(*reinterpret_cast<const uint4*>(s + j), v + j);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[j + q] = (col + j + q < n) ? __bfloat162float(s[j + q]) : 0.f;
    }
  }
}

// group barrier: the PARTS epilogue warps that share a TMEM lane group (ids 1..4; 0 is __syncthreads)
template <int PARTS>
__device__ __forceinline__ void group_sync(int lg) {
  asm volatile("bar.sync %0, %1;" ::"r"(lg + 1), "n"(32 * PARTS) : "memory");
}

template <int BN, int AMAJ, int BMAJ, int EPI, int CG>
__global__ void __launch_bounds__(64 + 32 * epi_warps<EPI>(), 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Args g) {
  constexpr int kEpiWarps = epi_warps<EPI>();
  constexpr int kEpiParts = kEpiWarps / 4;
  using T = Tile<BN, CG, kEpiWarps>;
  constexpr int BNH = BN / CG;                       // rows of B staged by one CTA
  constexpr int kStages = T::kStages;
  constexpr uint32_t kStageBytes = T::kStageBytes;
  extern __shared__ __align__(1024) unsigned char gsm_raw[];
  unsigned char* sm = gsm_raw;                                   // swizzle-128B atoms need 1024-byte alignment
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  unsigned char* staging = sm + (size_t)kStages * kStageBytes;   // [kEpiWarps][32 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + T::kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  unsigned long long gt_entry;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;       // 0 = leader: issues the MMAs, owns the barriers
  const int cta = blockIdx.x / CG, ncta = gridDim.x / CG;      // work is dealt to CTA pairs

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar + s, 1);
        mbar_init(empty_bar + s, 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar + a, 1);
        mbar_init(tempty_bar + a, kEpiWarps * CG);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(T::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {   // both CTAs of the pair issue the allocation, from the same warp
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(T::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 1) __syncthreads();
  else cluster_sync_all();          // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  unsigned long long gt_ready;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_ready));

  const int tiles_per_slice = g.nb_n * g.nb_m;
  const int items = tiles_per_slice * g.split_k * g.batches;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      long long dbg_empty = 0, dbg_p0 = clock64();
      for (int item = cta; item < items; item += ncta) {
        const int n_blk = item % g.nb_n;
        int r = item / g.nb_n;
        const int m_blk = r % g.nb_m;
        r /= g.nb_m;
        const int ks = r % g.split_k;
        const int bb = r / g.split_k;
        const int bo = bb / g.batch_inner, bi = bb - bo * g.batch_inner;
        const int kb0 = ks * g.kb_per_split;
        const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
        // this CTA's 128 rows of A and its BN / CG rows of B
        const int m0 = m_blk * (BM * CG) + rank * BM, n0 = n_blk * BN + rank * BNH;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int kbatch = kb / g.kpb;
          const int k0 = (kb - kbatch * g.kpb) * BK;
          const int bo_eff = g.k_batches > 1 ? kbatch : bo;
          long long te0 = clock64();
          mbar_wait(empty_bar + s, ph ^ 1);
          dbg_empty += clock64() - te0;
          unsigned char* sa = sm + (size_t)s * kStageBytes;
          unsigned char* sb = sa + kABytes;
          if constexpr (CG == 1) {
            mbar_expect_tx(full_bar + s, kStageBytes);
            if constexpr (AMAJ == 0) {
              tma_load_4d(sa, &map_a, full_bar + s, k0, m0, bi * g.a_bi, bo_eff * g.a_bo);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_4d(sa + j * 8192, &map_a, full_bar + s, m0 + 64 * j, k0, bi * g.a_bi, bo_eff * g.a_bo);
            }
            if constexpr (BMAJ == 0) {
              tma_load_4d(sb, &map_b, full_bar + s, k0, n0, bi * g.b_bi, bo_eff * g.b_bo);
            } else {
#pragma unroll
              for (int j = 0; j < BNH / 64; ++j)
                tma_load_4d(sb + j * 8192, &map_b, full_bar + s, n0 + 64 * j, k0, bi * g.b_bi, bo_eff * g.b_bo);
            }
          } else {
            // both CTAs' bytes are counted on the LEADER's barrier, which its MMA thread waits on
            if (rank == 0) mbar_expect_tx(full_bar + s, kStageBytes * 2);
            const uint32_t lbar = mapa_rank(smem_u32(full_bar + s), 0);
            if constexpr (AMAJ == 0) {
              tma_load_4d_pair(sa, &map_a, lbar, k0, m0, bi * g.a_bi, bo_eff * g.a_bo);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_4d_pair(sa + j * 8192, &map_a, lbar, m0 + 64 * j, k0, bi * g.a_bi, bo_eff * g.a_bo);
            }
            if constexpr (BMAJ == 0) {
              tma_load_4d_pair(sb, &map_b, lbar, k0, n0, bi * g.b_bi, bo_eff * g.b_bo);
            } else {
#pragma unroll
              for (int j = 0; j < BNH / 64; ++j)
                tma_load_4d_pair(sb + j * 8192, &map_b, lbar, n0 + 64 * j, k0, bi * g.b_bi, bo_eff * g.b_bo);
            }
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
      if ((g.debug & 16) && blockIdx.x == 0) {
        long long* d = reinterpret_cast<long long*>(g.aux);
        d[3] = clock64() - dbg_p0; d[4] = dbg_empty;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = instr_desc<BM * CG, BN, AMAJ, BMAJ>();
      int s = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      long long dbg_tempty = 0, dbg_full = 0, dbg_t0 = clock64();
      for (int item = cta; item < items; item += ncta) {
        const int ks = (item / tiles_per_slice) % g.split_k;
        const int kb0 = ks * g.kb_per_split;
        const int nkb = min(g.kb_total, kb0 + g.kb_per_split) - kb0;
        long long tw0 = clock64();
        mbar_wait(tempty_bar + acc, aph ^ 1);   // the epilogue has drained this accumulator
        dbg_tempty += clock64() - tw0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int i = 0; i < nkb; ++i) {
          long long tf0 = clock64();
          mbar_wait(full_bar + s, ph);
          dbg_full += clock64() - tf0;
          tc_fence_after();
          const uint32_t sa = smem_u32(sm + (size_t)s * kStageBytes);
          const uint64_t da = umma_desc<AMAJ>(sa);
          const uint64_t db = umma_desc<BMAJ>(sa + kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // 16 k: K-major +32 B inside the swizzled row; MN-major +16 k-rows of 128 B
            const uint64_t ka = (uint64_t)((AMAJ == 0 ? 32 : 2048) * k >> 4);
            const uint64_t kb = (uint64_t)((BMAJ == 0 ? 32 : 2048) * k >> 4);
            if constexpr (CG == 1) tc_mma(tmem_d, da + ka, db + kb, idesc, (i | k) ? 1u : 0u);
            else tc_mma_pair(tmem_d, da + ka, db + kb, idesc, (i | k) ? 1u : 0u);
          }
          // frees the stage (in both CTAs of a pair) once the MMAs that read it have retired
          if constexpr (CG == 1) tc_commit(empty_bar + s);
          else tc_commit_pair(empty_bar + s);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        if constexpr (CG == 1) tc_commit(tfull_bar + acc);
        else tc_commit_pair(tfull_bar + acc);
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
      if ((g.debug & 16) && blockIdx.x == 0) {
        long long* d = reinterpret_cast<long long*>(g.aux);
        d[0] = clock64() - dbg_t0; d[1] = dbg_tempty; d[2] = dbg_full;
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        d[5] = (long long)gt_entry; d[6] = (long long)gt_ready; d[7] = (long long)gt;
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fused elementwise / row-wise work -> global =====
    const int ew = warp - 2;
    const int lg = warp & 3;             // TMEM lane group this warp may touch
    const int part = ew >> 2;            // which share of the tile's columns
    constexpr int CW = (BN / kEpiParts) < 32 ? 32 : (BN / kEpiParts);   // columns per warp
    const bool active = part * CW < BN;  // (narrow tiles have fewer 32-column blocks than warps)
    int acc = 0;
    uint32_t aph = 0;
    for (int item = cta; item < items; item += ncta) {
      const int n_blk = item % g.nb_n;
      int r = item / g.nb_n;
      const int m_blk = r % g.nb_m;
      r /= g.nb_m;
      const int bb = r / g.split_k;
      const int bo = bb / g.batch_inner, bi = bb - bo * g.batch_inner;
      const int row0 = m_blk * (BM * CG) + rank * BM + lg * 32;      // first row of this warp
      const int row = row0 + lane;
      const bool row_ok = row < g.m;
      const int rows_valid = min(32, g.m - row0);
      const int n0 = n_blk * BN + part * CW;     // first global column of this warp
      const size_t esz = g.out_f32 ? 4 : 2;
      unsigned char* out_w = reinterpret_cast<unsigned char*>(g.out) +
                             esz * (size_t)(bo * g.c_bo + bi * g.c_bi + (long long)row0 * g.ldc);
      unsigned char* aux_w = reinterpret_cast<unsigned char*>(g.aux) +
                             2 * (size_t)(bo * g.aux_bo + bi * g.aux_bi + (long long)row0 * g.ld_aux);
      unsigned char* out_row = out_w + esz * (size_t)lane * g.ldc;
      const unsigned char* aux_row = aux_w + 2 * (size_t)lane * g.ld_aux;
      unsigned char* stg = staging + ew * 4096;
      // one 32-column block of this warp's rows -> global (coalesced through the staging tile when aligned)
      const bool p_vec = g.vec_ok, p_f32 = g.out_f32, p_acc = g.accumulate, p_atomic = g.atomic, p_skip = g.debug & 1;
      const long long p_ldc = g.ldc, p_ldaux = g.ld_aux;
      const int p_n = g.n;
      auto put = [&](int col, const float* v, bool to_aux, const uint4* old = nullptr) {
        if (p_skip) return;
        if (to_aux) {
          if (p_vec) warp_store32<false>(stg, lane, aux_w, 2 * p_ldaux, rows_valid, col, p_n, v, false, false);
          else if (row_ok) store_row32<false>(const_cast<unsigned char*>(aux_row), col, p_n, v, false, false, false);
        } else if (p_f32) {
          if (p_vec) warp_store32<true>(stg, lane, out_w, 4 * p_ldc, rows_valid, col, p_n, v, p_acc, p_atomic);
          else if (row_ok) store_row32<true>(out_row, col, p_n, v, false, p_acc, p_atomic);
        } else {
          if (p_vec) warp_store32<false>(stg, lane, out_w, 2 * p_ldc, rows_valid, col, p_n, v, p_acc, false, old);
          else if (row_ok) store_row32<false>(out_row, col, p_n, v, false, p_acc, false);
        }
      };
      // operands of the epilogue that do not depend on the accumulator are requested before waiting for it
      [[maybe_unused]] uint4 wq[4];                      // GELU': the next 32-column block of the pre-activation;
                                                         // accumulating bf16 store: the next block of old values
      [[maybe_unused]] uint4 wall[CW / 32][4];           // dsoftmax: the warp's whole block of P
      const bool pre_old = EPI == EPI_STORE && p_acc && !p_f32 && p_vec;
      if (active) {
        if constexpr (EPI == EPI_GELU_BWD || EPI == EPI_MUL_AUX) {
          if (p_vec && n0 < p_n) warp_ldg32(aux_w, 2 * p_ldaux, rows_valid, n0, p_n, lane, wq);
        } else if constexpr (EPI == EPI_DSOFTMAX) {
          if (p_vec) {
#pragma unroll
            for (int c = 0; c < CW; c += 32)
              if (n0 + c < p_n) warp_ldg32(aux_w, 2 * p_ldaux, rows_valid, n0 + c, p_n, lane, wall[c / 32]);
          }
        } else if constexpr (EPI == EPI_STORE) {
          if (pre_old && n0 < p_n) warp_ldg_old32(out_w, 2 * p_ldc, rows_valid, n0, p_n, lane, wq);
        }
      }
      mbar_wait_warp(tfull_bar + acc, aph, lane);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * BN + part * CW);

      // kernel parameters used per element are hoisted into registers: tested where they are used they cost a
      // constant-bank load and a uniform-register move per ELEMENT (measured: 2x on the whole kernel)
      const int ncols = g.n;
      if constexpr (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_GELU_BWD || EPI == EPI_MUL_AUX) {
        if (active) {
        const float* bias_n = g.bias_n;
        const bool has_bm = g.bias_m != nullptr;
        const float bm = (has_bm && row_ok) ? g.bias_m[row] : 0.f;
        // bias of the next 32 columns: one coalesced load per lane, a block ahead of its use, broadcast to the rows
        // through the staging tile
        bool done = false;
        if constexpr (EPI == EPI_STORE && CW % 64 == 0) {
          // plain bf16 store: 64 columns per step, so that every row segment that leaves is a whole 128-byte line
          if (!p_f32 && p_vec && !p_acc && !p_skip) {
            done = true;
#pragma unroll 1
            for (int c = 0; c < CW; c += 64) {
              if (n0 + c >= ncols) break;
              uint32_t rr[64];
              tmem_ld32_nowait(taddr + c, rr);
              tmem_ld32_nowait(taddr + c + 32, rr + 32);
              float bl0 = 0.f, bl1 = 0.f;
              if (bias_n != nullptr) {
                if (n0 + c + lane < ncols) bl0 = __ldg(bias_n + n0 + c + lane);
                if (n0 + c + 32 + lane < ncols) bl1 = __ldg(bias_n + n0 + c + 32 + lane);
              }
              tmem_wait_ld_tied(rr);
              reg_tie32(rr + 32);
              float* v = reinterpret_cast<float*>(rr);
              if (has_bm) {
#pragma unroll
                for (int j = 0; j < 64; ++j) v[j] += bm;
              }
              if (bias_n != nullptr) {
                reinterpret_cast<float*>(stg)[lane] = bl0;
                reinterpret_cast<float*>(stg)[32 + lane] = bl1;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 64; j += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(stg + 4 * j);
                  v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                }
                __syncwarp();
              }
              warp_store64_bf16(stg, lane, out_w, 2 * p_ldc, rows_valid, n0 + c, ncols, v, 1.f);
            }
          }
        }
        float bnext = (!done && bias_n != nullptr && n0 + lane < ncols) ? __ldg(bias_n + n0 + lane) : 0.f;
#pragma unroll 1
        for (int c = 0; c < CW; c += 32) {
          if (done || n0 + c >= ncols) break;    // warp-uniform: nothing left in this row block
          uint32_t rr[32];
          tmem_ld32(taddr + c, rr);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
          if (has_bm) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bm;
          }
          if (bias_n != nullptr) {
            reinterpret_cast<float*>(stg)[lane] = bnext;
            const int cn = n0 + c + 32 + lane;
            bnext = (c + 32 < CW && cn < ncols) ? __ldg(bias_n + cn) : 0.f;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(stg + 4 * j);
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
            __syncwarp();
          }
          if constexpr (EPI == EPI_GELU) {
            if (g.aux && g.aux_gelu_grad) {
              float gp[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) gelu_both_tc(v[j], v[j], gp[j]);
              put(n0 + c, gp, true);
            } else {
              if (g.aux) put(n0 + c, v, true);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_tc(v[j]);
            }
          } else if constexpr (EPI == EPI_GELU_BWD || EPI == EPI_MUL_AUX) {
            float pre[32];
            if (p_vec) {
              uint4 raw[4];
              warp_own_rows(stg, lane, wq, raw);
              if (c + 32 < CW && n0 + c + 32 < ncols)     // the block after this one, a whole block ahead of its use
                warp_ldg32(aux_w, 2 * p_ldaux, rows_valid, n0 + c + 32, ncols, lane, wq);
              unpack32(raw, pre);
            } else if (row_ok) {
              load_row32_bf16(aux_row, n0 + c, ncols, pre, false);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) pre[j] = 0.f;
            }
            if constexpr (EPI == EPI_GELU_BWD) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= gelu_grad_tc(pre[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= pre[j];
            }
          }
          if constexpr (EPI == EPI_STORE) {
            if (pre_old) {
              uint4 old[4] = {wq[0], wq[1], wq[2], wq[3]};
              if (c + 32 < CW && n0 + c + 32 < ncols)
                warp_ldg_old32(out_w, 2 * p_ldc, rows_valid, n0 + c + 32, ncols, lane, wq);
              put(n0 + c, v, false, old);
            } else {
              put(n0 + c, v, false);
            }
          } else {
            put(n0 + c, v, false);
          }
        }
        }
      } else if constexpr (EPI == EPI_SOFTMAX) {
        // P = softmax_n(scale * S + key mask): the whole row lives in this tile (n <= BN); the kEpiParts warps of a
        // lane group own a share of the columns each and exchange (max, sum) through shared memory.  The warp's
        // columns are read from TMEM ONCE and stay in registers (one MUFU.EX2 per element); the reductions run on
        // four independent accumulators (the chains, not the issue slots, bound the two-warps-per-scheduler epilogue).
        static_assert(BN == 256 && CW % 32 == 0, "row-wise epilogues use the 256-column tile");
        const float sc = g.scale * kLog2e;       // > 0 (checked by the host side)
        const unsigned char* mk = g.mask ? g.mask + (size_t)bo * g.mask_bo : nullptr;
        uint32_t keep[CW / 32];                  // bit j of keep[c / 32]: column n0 + c + j takes part
#pragma unroll
        for (int c = 0; c < CW; c += 32) {
          const int col = n0 + c + lane;
          const bool on = col < ncols && (mk == nullptr || mk[col] != 0);
          keep[c / 32] = __ballot_sync(0xffffffffu, on);
        }
        uint32_t sr[CW];
#pragma unroll
        for (int c = 0; c < CW; c += 32) tmem_ld32_nowait(taddr + c, sr + c);
        tmem_wait_ld_tied(sr);
#pragma unroll
        for (int c = 32; c < CW; c += 32) reg_tie32(sr + c);
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < CW; c += 32) {
          if (keep[c / 32] != 0xffffffffu) {     // warp-uniform: some column of the block is masked or out of range
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (!((keep[c / 32] >> j) & 1u)) sr[c + j] = 0xff800000u;   // -inf
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(sr[c + j]));
        }
        float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        // exchange slots: the head of every warp's own staging tile; the warps of this lane group are ew & 3 (mod 4)
        float* slot = reinterpret_cast<float*>(stg);
        auto other_slot = [&](int q) { return reinterpret_cast<float*>(staging + (size_t)(q * 4 + (ew & 3)) * 4096); };
        slot[2 * lane] = mx;
        group_sync<kEpiParts>(lg);
#pragma unroll
        for (int q = 0; q < kEpiParts; ++q) mx = fmaxf(mx, other_slot(q)[2 * lane]);
        const float base = mx == -INFINITY ? 0.f : mx * sc;   // fully masked row: zeros (the reference gives NaN)
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        float* e = reinterpret_cast<float*>(sr);
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          e[j] = ex2f(fmaf(__uint_as_float(sr[j]), sc, -base));      // ex2(-inf) = 0 for the masked columns
          s4[j & 3] += e[j];
        }
        float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        slot[2 * lane + 1] = sum;
        group_sync<kEpiParts>(lg);
        sum = 0.f;
#pragma unroll
        for (int q = 0; q < kEpiParts; ++q) sum += other_slot(q)[2 * lane + 1];
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        if (g.row_stat != nullptr && part == 0 && row_ok)      // base-2 log-sum-exp of the row (fused backward)
          g.row_stat[(size_t)bb * g.m + row] = sum > 0.f ? base + lg2f(sum) : INFINITY;
        group_sync<kEpiParts>(lg);    // every warp has read the slots: the staging tiles are free for the stores
        if (!p_skip) {
          if (p_vec && CW % 64 == 0) {
#pragma unroll
            for (int c = 0; c < CW; c += 64)
              if (n0 + c < ncols) warp_store64_bf16(stg, lane, out_w, 2 * p_ldc, rows_valid, n0 + c, ncols, e + c, inv);
          } else {
#pragma unroll
            for (int c = 0; c < CW; c += 32) {
              if (n0 + c < ncols) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = e[c + j] * inv;
                put(n0 + c, v, false);
              }
            }
          }
        }
        group_sync<kEpiParts>(lg);    // the staging tiles double as exchange slots of the next item
      } else {
        // dS = scale * P o (dP - delta),  delta = sum_n P o dP  (softmax backward; P read from aux, once: its
        // columns stay packed in registers between the two sweeps over dP in TMEM; 8 columns are unpacked at a time
        // so that the sweeps fit the 96 registers of the 16-warp epilogue)
        static_assert(BN == 256 && CW % 32 == 0, "row-wise epilogues use the 256-column tile");
        const float sc = g.scale;
        uint4 (&praw)[CW / 32][4] = wall;      // the prefetched block, turned into own rows in place
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < CW; c += 32) {
          if (n0 + c < ncols) {                  // warp-uniform
            uint32_t rr[32];
            tmem_ld32_nowait(taddr + c, rr);
            if (p_vec) {
              warp_own_rows(stg, lane, wall[c / 32], praw[c / 32]);
            } else {
              float pv[32];
              if (row_ok) {
                load_row32_bf16(aux_row, n0 + c, ncols, pv, false);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) pv[j] = 0.f;
              }
#pragma unroll
              for (int sl = 0; sl < 4; ++sl)
                praw[c / 32][sl] = make_uint4(pack_bf16(pv[8 * sl], pv[8 * sl + 1]), pack_bf16(pv[8 * sl + 2], pv[8 * sl + 3]),
                                              pack_bf16(pv[8 * sl + 4], pv[8 * sl + 5]), pack_bf16(pv[8 * sl + 6], pv[8 * sl + 7]));
            }
            tmem_wait_ld_tied(rr);
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              float pv[8];
              unpack_bf16x8(praw[c / 32][sl], pv);
#pragma unroll
              for (int q = 0; q < 8; ++q) d4[q & 3] = fmaf(pv[q], __uint_as_float(rr[8 * sl + q]), d4[q & 3]);
            }
          }
        }
        float delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        reinterpret_cast<float*>(stg)[lane] = delta;       // exchange slot: the head of the warp's own staging tile
        group_sync<kEpiParts>(lg);
        delta = 0.f;
#pragma unroll
        for (int q = 0; q < kEpiParts; ++q)
          delta += reinterpret_cast<float*>(staging + (size_t)(q * 4 + (ew & 3)) * 4096)[lane];
        group_sync<kEpiParts>(lg);      // all slots read: the staging tiles are free for the stores
#pragma unroll
        for (int c = 0; c < CW; c += 32) {
          if (n0 + c < ncols) {
            uint32_t rr[32];
            tmem_ld32(taddr + c, rr);
            float v[32];
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              float pv[8];
              unpack_bf16x8(praw[c / 32][sl], pv);
#pragma unroll
              for (int q = 0; q < 8; ++q) v[8 * sl + q] = sc * pv[q] * (__uint_as_float(rr[8 * sl + q]) - delta);
            }
            put(n0 + c, v, false);
          }
        }
        group_sync<kEpiParts>(lg);
      }

      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(tempty_bar + acc);
        else mbar_arrive_cluster(mapa_rank(smem_u32(tempty_bar + acc), 0));   // the leader's MMA thread waits on it
      }
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }

  if ((g.debug & 16) && blockIdx.x == 0 && threadIdx.x == 64) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    reinterpret_cast<long long*>(g.aux)[8] = (long long)gt;
  }
  tc_fence_before();
  if constexpr (CG == 1) __syncthreads();
  else cluster_sync_all();          // neither CTA may retire while the pair's MMAs / signals can still touch it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T::kTmemCols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T::kTmemCols)
                   : "memory");
    if ((g.debug & 16) && lane == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      if (blockIdx.x == 0) reinterpret_cast<long long*>(g.aux)[9] = (long long)gt;
      reinterpret_cast<long long*>(g.aux)[16 + 2 * blockIdx.x] = (long long)gt_entry;
      reinterpret_cast<long long*>(g.aux)[17 + 2 * blockIdx.x] = (long long)gt;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
template <int BN, int AMAJ, int BMAJ, int EPI, int CG>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Args& a, int grid, cudaStream_t stream) {
  auto kern = gemm_kernel<BN, AMAJ, BMAJ, EPI, CG>;
  const size_t smem = Tile<BN, CG, epi_warps<EPI>()>::kSmemBytes;
  static bool configured = false;   // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -static_cast<int>(e);
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid * CG);          // `grid` CTAs or CTA pairs
  cfg.blockDim = dim3(64 + 32 * epi_warps<EPI>());
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, a);
  if (e != cudaSuccess) return -static_cast<int>(e);
  return launch_status();
}

template <int BN, int EPI, int CG>
static int launch_major(int am, int bm, const CUtensorMap& ma, const CUtensorMap& mb, const Args& a, int grid,
                        cudaStream_t s) {
  if (am == 0 && bm == 0) return launch<BN, 0, 0, EPI, CG>(ma, mb, a, grid, s);
  if constexpr (EPI == EPI_STORE) {
    if (am == 0 && bm == 1) return launch<BN, 0, 1, EPI, CG>(ma, mb, a, grid, s);
    if (am == 1 && bm == 0) return launch<BN, 1, 0, EPI, CG>(ma, mb, a, grid, s);
    return launch<BN, 1, 1, EPI, CG>(ma, mb, a, grid, s);
  } else if constexpr (EPI == EPI_GELU_BWD || EPI == EPI_MUL_AUX) {
    if (am == 0 && bm == 1) return launch<BN, 0, 1, EPI, CG>(ma, mb, a, grid, s);
  }
  return MTTS_ERR_UNSUPPORTED;
}

}  // namespace g100
}  // namespace mtts

extern "C" int mtts_gemm(const mtts_gemm_params* p, mtts_stream_t stream) {
  using namespace mtts::g100;
  if (!p || !p->a || !p->b || !p->out) return MTTS_ERR_NULL;
  if (p->m < 0 || p->n < 1 || p->k < 1 || p->batch_outer < 0 || p->batch_inner < 1) return MTTS_ERR_SHAPE;
  if (p->a_major < 0 || p->a_major > 1 || p->b_major < 0 || p->b_major > 1) return MTTS_ERR_UNSUPPORTED;
  if (p->epilogue < EPI_STORE || p->epilogue > EPI_MUL_AUX) return MTTS_ERR_UNSUPPORTED;
  if (p->out_dtype != MTTS_F32 && p->out_dtype != MTTS_BF16) return MTTS_ERR_DTYPE;
  if (p->m == 0 || p->batch_outer == 0) return MTTS_OK;
  const int k_batches = p->k_batches > 1 ? p->k_batches : 1;
  if (k_batches > 1 && p->batch_outer != 1) return MTTS_ERR_SHAPE;
  // TMA: 16-byte aligned bases and strides
  if (!mtts::aligned16(p->a) || !mtts::aligned16(p->b) || p->lda % 8 || p->ldb % 8 || p->a_bo_stride % 8 ||
      p->a_bi_stride % 8 || p->b_bo_stride % 8 || p->b_bi_stride % 8)
    return MTTS_ERR_ALIGN;
  const bool row_epi = p->epilogue == EPI_SOFTMAX || p->epilogue == EPI_DSOFTMAX;
  if (row_epi && (p->n > 256 || p->out_dtype != MTTS_BF16)) return MTTS_ERR_SHAPE;
  if (p->epilogue == EPI_SOFTMAX && !(p->scale > 0.f)) return MTTS_ERR_UNSUPPORTED;   // the row max is taken before scaling
  if ((p->epilogue == EPI_GELU_BWD || p->epilogue == EPI_DSOFTMAX || p->epilogue == EPI_MUL_AUX) && !p->aux)
    return MTTS_ERR_NULL;
  if (p->epilogue != EPI_STORE && p->out_dtype != MTTS_BF16) return MTTS_ERR_DTYPE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  const int BN = (row_epi || p->epilogue == EPI_GELU || p->epilogue == EPI_GELU_BWD || p->epilogue == EPI_MUL_AUX ||
                  p->n > 128) ? 256
                 : (p->n > 64 ? 128 : 64);
  Args a{};
  a.m = p->m;
  a.n = p->n;
  // CTA pairs (tcgen05 cta_group::2, 256-row tiles) whenever the problem has more than one 128-row block
  const int CG = (p->m > BM && BN >= 128 && !(p->flags & MTTS_GEMM_SINGLE_CTA)) ? 2 : 1;   // 64-row B halves at least
  a.nb_m = (p->m + BM * CG - 1) / (BM * CG);
  a.nb_n = (p->n + BN - 1) / BN;
  a.kpb = (p->k + BK - 1) / BK;
  a.k_batches = k_batches;
  a.kb_total = a.kpb * k_batches;
  a.batch_inner = p->batch_inner;
  a.batches = p->batch_outer * p->batch_inner;
  const long long base_items = (long long)a.nb_m * a.nb_n * a.batches;
  int split = p->split_k;
  if (split < 0) {   // auto: about two waves of work items, at least 4 k-blocks each
    split = 1;
    const long long units = mtts::kNumSMs / CG;      // CTAs or CTA pairs that run at a time
    if (base_items < units) {
      split = (int)((2 * units) / base_items);         // at most two full rounds of work items
      if (split > a.kb_total / 4) split = a.kb_total / 4;
      if (split < 1) split = 1;
    }
  }
  if (split < 1) split = 1;
  if (split > 1 && (p->out_dtype != MTTS_F32 || p->epilogue != EPI_STORE || a.batches != 1 || p->bias_n || p->bias_m))
    return MTTS_ERR_UNSUPPORTED;
  a.kb_per_split = (a.kb_total + split - 1) / split;
  a.split_k = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;   // every slice gets at least one k-block
  a.a_bi = p->a_bi_stride != 0;
  a.a_bo = p->a_bo_stride != 0;
  a.b_bi = p->b_bi_stride != 0;
  a.b_bo = p->b_bo_stride != 0;
  a.out_f32 = p->out_dtype == MTTS_F32;
  a.accumulate = p->accumulate != 0;
  a.atomic = a.split_k > 1;
  a.out = p->out;
  a.ldc = p->ldc;
  a.c_bo = p->c_bo_stride;
  a.c_bi = p->c_bi_stride;
  a.bias_n = p->bias_n;
  a.bias_m = p->bias_m;
  a.aux = p->aux;
  a.ld_aux = p->ld_aux;
  a.aux_bo = p->aux_bo_stride;
  a.aux_bi = p->aux_bi_stride;
  a.mask = p->mask;
  a.mask_bo = p->mask_bo_stride;
  a.scale = p->scale;
  a.row_stat = p->row_stat;
  a.debug = p->flags >> 8;
  a.aux_gelu_grad = (p->flags & MTTS_GEMM_AUX_GELU_GRAD) != 0;
  const int ov = a.out_f32 ? 4 : 8;   // elements per 16-byte vector of the output
  a.vec_ok = mtts::aligned16(p->out) && p->ldc % ov == 0 && p->c_bo_stride % ov == 0 && p->c_bi_stride % ov == 0 &&
             p->n % 8 == 0 &&
             (!p->aux || (mtts::aligned16(p->aux) && p->ld_aux % 8 == 0 && p->aux_bo_stride % 8 == 0 &&
                          p->aux_bi_stride % 8 == 0));
  if (a.atomic && !a.accumulate) {
    // the split-k slices add into the output: start from zero
    cudaError_t e = cudaMemset2DAsync(p->out, (size_t)p->ldc * 4, 0, (size_t)p->n * 4, (size_t)p->m, s);
    if (e != cudaSuccess) return -static_cast<int>(e);
  }

  // operand maps.  K-major: d0 = k, d1 = rows, box {64, tile rows};  MN-major: d0 = rows, d1 = k, box {64, 64}.
  // With k_batches the outer-batch dimension of the map is the reduction batch.
  const int64_t nbo = k_batches > 1 ? k_batches : p->batch_outer;
  CUtensorMap ma, mb;
  const bool oka = p->a_major == 0
                       ? make_map(&ma, p->a, p->k, p->m, p->lda, p->batch_inner, p->a_bi_stride, nbo, p->a_bo_stride, BK, BM)
                       : make_map(&ma, p->a, p->m, p->k, p->lda, p->batch_inner, p->a_bi_stride, nbo, p->a_bo_stride, 64, BK);
  const bool okb = p->b_major == 0
                       ? make_map(&mb, p->b, p->k, p->n, p->ldb, p->batch_inner, p->b_bi_stride, nbo, p->b_bo_stride, BK, BN / CG)
                       : make_map(&mb, p->b, p->n, p->k, p->ldb, p->batch_inner, p->b_bi_stride, nbo, p->b_bo_stride, 64, BK);
  if (!oka || !okb) return MTTS_ERR_UNSUPPORTED;

  const long long items = base_items * a.split_k;
  const long long units = mtts::kNumSMs / CG;
  const int grid = (int)(items < units ? items : units);
#define MTTS_GEMM_LAUNCH(BN_, EPI_)                                                             \
  (CG == 2 ? launch_major<BN_, EPI_, 2>(p->a_major, p->b_major, ma, mb, a, grid, s)              \
           : launch_major<BN_, EPI_, 1>(p->a_major, p->b_major, ma, mb, a, grid, s))
  switch (p->epilogue) {
    case EPI_STORE:
      if (BN == 256) return MTTS_GEMM_LAUNCH(256, EPI_STORE);
      if (BN == 128) return MTTS_GEMM_LAUNCH(128, EPI_STORE);
      return MTTS_GEMM_LAUNCH(64, EPI_STORE);
    case EPI_GELU: return MTTS_GEMM_LAUNCH(256, EPI_GELU);
    case EPI_GELU_BWD: return MTTS_GEMM_LAUNCH(256, EPI_GELU_BWD);
    case EPI_SOFTMAX: return MTTS_GEMM_LAUNCH(256, EPI_SOFTMAX);
    case EPI_MUL_AUX: return MTTS_GEMM_LAUNCH(256, EPI_MUL_AUX);
    default: return MTTS_GEMM_LAUNCH(256, EPI_DSOFTMAX);
  }
#undef MTTS_GEMM_LAUNCH
}

// The round-1 entry point (FFN-shaped: out = act(a @ w^T + bias), optional pre-activation output), now a view
// onto the general kernel.
extern "C" int mtts_gemm_bf16(const mtts_gemm_bf16_params* p, mtts_stream_t stream) {
  if (!p || !p->a || !p->w || !p->out) return MTTS_ERR_NULL;
  if (p->m < 0 || p->n < 1 || p->k < 8 || p->k % 8 != 0 || p->n % 8 != 0) return MTTS_ERR_SHAPE;
  if (p->lda % 8 != 0 || p->ldw % 8 != 0 || p->ldo % 8 != 0 || p->lda < p->k || p->ldw < p->k || p->ldo < p->n)
    return MTTS_ERR_ALIGN;
  if (!mtts::aligned16(p->a) || !mtts::aligned16(p->w) || !mtts::aligned16(p->out) ||
      (p->pre_out && !mtts::aligned16(p->pre_out)))
    return MTTS_ERR_ALIGN;
  mtts_gemm_params g{};
  g.m = p->m; g.n = p->n; g.k = p->k;
  g.batch_outer = 1; g.batch_inner = 1; g.k_batches = 1;
  g.out_dtype = MTTS_BF16;
  g.epilogue = p->gelu ? MTTS_EPI_GELU : MTTS_EPI_STORE;
  g.a = p->a; g.lda = p->lda;
  g.b = p->w; g.ldb = p->ldw;
  g.out = p->out; g.ldc = p->ldo;
  g.bias_n = p->bias;
  if (p->gelu) {
    g.aux = p->pre_out; g.ld_aux = p->ldo;
  }
  const int rc = mtts_gemm(&g, stream);
  if (rc == MTTS_OK && !p->gelu && p->pre_out && p->m > 0) {   // no activation: the pre-activation IS the output
    cudaError_t e = cudaMemcpy2DAsync(p->pre_out, (size_t)p->ldo * 2, p->out, (size_t)p->ldo * 2, (size_t)p->n * 2,
                                      (size_t)p->m, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return -static_cast<int>(e);
  }
  return rc;
}
