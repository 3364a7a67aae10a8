// Pieces shared by the selective-scan forward and backward kernels.
//
// Work decomposition (both directions):
//   CTA   = one batch element x a group of G channels, walking the sequence tile by tile;
//   warp  = one channel at a time; its 32 lanes split a tile of 32*kItems timesteps, each lane owning
//           kItems CONSECUTIVE timesteps (recurrence in registers), lanes are stitched together with a
//           warp-shuffle scan of the affine pairs (decay, state); two dstate rows are processed per
//           instruction with packed fp32x2 arithmetic;
//   smem  = the B/C tile of the batch element (dstate row pairs x tile timesteps, fp32), staged once
//           and shared by every channel of the group -- upstream re-reads it from L2 per channel.
#pragma once

#include <cstdlib>

#include "common.cuh"

namespace mtts {

constexpr int kScanNChunk = 16;  // dstate rows resident in shared memory at a time

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) -----
// A make_float2(s, s) operand folds into a scalar-broadcast register operand (R.F32) in SASS.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  return d;
}
__device__ __forceinline__ float2 dup2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 ex2f2(float2 x) { return make_float2(ex2f(x.x), ex2f(x.y)); }
__device__ __forceinline__ float2 shfl_up2(float2 v, int d) {
  return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ float2 shfl_down2(float2 v, int d) {
  return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}

// ---- cp.async (LDGSTS) staging into thread-private shared slots ----------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- pair-interleaved tile: two dstate rows (2p, 2p+1) share one smem row ---------------------------
// element (n, t) lives at  (n/2)*kRow + (t/kItems)*kSeg + (t%kItems)*2 + (n&1):  a lane reads its
// kItems timesteps of BOTH rows as kItems float2 with LDS.128; kSeg = 2*kItems + 4 keeps the eight
// lanes of a quarter-warp on distinct 16-byte bank groups.
template <int kItems>
struct PairTile {
  static constexpr int kSeg = 2 * kItems + 4;
  static constexpr int kRow = 32 * kSeg;
  static constexpr int kLen = 32 * kItems;
  static constexpr int kPairs = kScanNChunk / 2;  // dstate row pairs resident at a time
  static_assert(kLen % MTTS_SCAN_CHUNK == 0, "tile vs chunk");
};

// Stage dstate rows [n0, n0+ncnt) (ncnt <= kScanNChunk) x timesteps [t0, t0+kLen) interleaved by
// pairs, as fp32; missing rows / timesteps >= len are zero (the scan's identity).
template <typename T, int kItems, bool kVec, int kThreads>
__device__ __forceinline__ void stage_pairs(const T* __restrict__ src, int64_t row_stride, int n0,
                                            int ncnt, int t0, int len, float* __restrict__ dst) {
  using Tile = PairTile<kItems>;
  const int npairs = (ncnt + 1) >> 1;
  if constexpr (kVec) {
    constexpr int VE = Io<T>::kVecElems;
    constexpr int kVecPerRow = Tile::kLen / VE;
    const int total = npairs * kVecPerRow;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
      const int pr = idx / kVecPerRow;
      const int tt = (idx - pr * kVecPerRow) * VE;
      float lo[VE], hi[VE];
      const bool in_t = t0 + tt < len;
      if (in_t) {
        Io<T>::unpack(ldg16(src + (int64_t)(n0 + 2 * pr) * row_stride + t0 + tt), lo);
      } else {
#pragma unroll
        for (int j = 0; j < VE; ++j) lo[j] = 0.f;
      }
      if (in_t && 2 * pr + 1 < ncnt) {
        Io<T>::unpack(ldg16(src + (int64_t)(n0 + 2 * pr + 1) * row_stride + t0 + tt), hi);
      } else {
#pragma unroll
        for (int j = 0; j < VE; ++j) hi[j] = 0.f;
      }
      float* d = dst + pr * Tile::kRow + (tt / kItems) * Tile::kSeg + (tt % kItems) * 2;
#pragma unroll
      for (int j = 0; j < VE; j += 2)
        *reinterpret_cast<float4*>(d + 2 * j) = make_float4(lo[j], hi[j], lo[j + 1], hi[j + 1]);
    }
  } else {
    const int total = npairs * 2 * Tile::kLen;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
      const int r = idx / Tile::kLen;  // row within the chunk (may be one past ncnt for odd ncnt)
      const int tt = idx - r * Tile::kLen;
      float v = 0.f;
      if (r < ncnt && t0 + tt < len) v = Io<T>::to_f(src[(int64_t)(n0 + r) * row_stride + t0 + tt]);
      dst[(r >> 1) * Tile::kRow + (tt / kItems) * Tile::kSeg + (tt % kItems) * 2 + (r & 1)] = v;
    }
  }
}

// This lane's kItems float2 (row 2p, row 2p+1) values of one staged pair row.
template <int kItems>
__device__ __forceinline__ void lane_pairs(const float* __restrict__ p, float2* out) {
#pragma unroll
  for (int j = 0; j < kItems; j += 2) {
    const float4 v = *reinterpret_cast<const float4*>(p + 2 * j);
    out[j] = make_float2(v.x, v.y);
    out[j + 1] = make_float2(v.z, v.w);
  }
}

// Inclusive warp scans of affine maps, two independent rows at once.
__device__ __forceinline__ void warp_scan_affine_up2(float2& P, float2& h, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float2 Pp = shfl_up2(P, off);
    const float2 hp = shfl_up2(h, off);
    if (lane >= off) {
      h = ffma2(P, hp, h);
      P = fmul2(P, Pp);
    }
  }
}
__device__ __forceinline__ void warp_scan_affine_down2(float2& P, float2& g, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float2 Pn = shfl_down2(P, off);
    const float2 gn = shfl_down2(g, off);
    if (lane + off < 32) {
      g = ffma2(P, gn, g);
      P = fmul2(P, Pn);
    }
  }
}

// True when (ptr, strides, seqlen) allow 16-byte vector access along the sequence.
template <typename T>
inline bool vec_ok(const void* p, int64_t s0, int64_t s1, int seqlen) {
  constexpr int VE = Io<T>::kVecElems;
  return p == nullptr || (aligned16(p) && s0 % VE == 0 && s1 % VE == 0 && seqlen % VE == 0);
}

// Sum 4 per-timestep scalars over the NG slice lanes of a channel (lane bits [0, log2 NG)) and store
// them to row[0..3].  The last two lane bits are folded with a transposing butterfly.
template <int NG>
__device__ __forceinline__ void slice_reduce_store(float (&v)[4], int g, float* row) {
#pragma unroll
  for (int o = NG / 2; o >= 4; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
  }
  if constexpr (NG == 1) {
    *reinterpret_cast<float4*>(row) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (NG == 2) {
    const bool hi = g & 1;
    const float s0 = hi ? v[0] : v[2], s1 = hi ? v[1] : v[3];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1);
    const float r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    const float k0 = (hi ? v[2] : v[0]) + r0, k1 = (hi ? v[3] : v[1]) + r1;
    *reinterpret_cast<float2*>(row + (hi ? 2 : 0)) = make_float2(k0, k1);
  } else {
    const bool b1 = g & 2, b0 = g & 1;
    const float s0 = b1 ? v[0] : v[2], s1 = b1 ? v[1] : v[3];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2);
    const float r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    const float k0 = (b1 ? v[2] : v[0]) + r0, k1 = (b1 ? v[3] : v[1]) + r1;
    const float s = b0 ? k0 : k1;
    const float r = __shfl_xor_sync(0xffffffffu, s, 1);
    const float tot = (b0 ? k1 : k0) + r;
    if (g < 4) row[g & 3] = tot;  // NG > 4: every group of 4 lanes holds the totals, the first writes
  }
}

// Which kernel family runs a scan.  The time-sequential kernels (scan_fwd.cu / scan_bwd.cu: one thread per
// channel and 4-state slice) need batch x dim channels to fill the machine; long sequences over few
// channels go to the time-parallel kernels (scan_*_wide.cu: lanes = timesteps).  mtts_set_scan_impl() overrides
// the choice (tests exercise both families on the same inputs); no environment is read on the launch path.
int scan_impl_override();   // api.cu: 0 = automatic, 1 = time-sequential, 2 = time-parallel ("wide")
inline bool scan_use_wide(int batch, int dim, int seqlen) {
  const int o = scan_impl_override();
  if (o == 1) return false;
  if (o == 2) return true;
  const int64_t channels = (int64_t)batch * dim;
  return channels < 8192 && channels * seqlen >= (int64_t(1) << 24);
}

}  // namespace mtts
