// Pieces shared by the selective-scan forward and backward kernels.
//
// Work decomposition (both directions):
//   CTA   = one batch element x a group of G channels, walking the sequence tile by tile;
//   warp  = one channel at a time; its 32 lanes split a tile of 32*kItems timesteps, each lane owning
//           kItems CONSECUTIVE timesteps (recurrence in registers), lanes are stitched together with a
//           warp-shuffle scan of the affine pairs (decay, state);
//   smem  = the B/C tile of the batch element (dstate rows x tile timesteps, fp32), staged once and
//           shared by every channel of the group -- upstream re-reads it from L2 per channel.
#pragma once

#include "common.cuh"

namespace mtts {

constexpr int kScanNChunk = 16;  // dstate rows resident in shared memory at a time

template <int kItems>
struct ScanTile {
  // A lane reads its kItems consecutive floats with LDS.128; padding every lane segment by 4 words
  // makes the 8 lanes of each quarter-warp phase hit 8 distinct 16-byte bank groups.
  static constexpr int kSeg = kItems + 4;
  static constexpr int kRow = 32 * kSeg;   // words per dstate row
  static constexpr int kLen = 32 * kItems; // timesteps per tile
  static_assert(MTTS_SCAN_CHUNK % kItems == 0 && kLen % MTTS_SCAN_CHUNK == 0,
                "tile must be a whole number of checkpoint chunks");
};

// Stage rows [n0, n0+ncnt) x timesteps [t0, t0+kLen) of a (dstate, seqlen) matrix into `dst` as fp32
// in the padded layout above; timesteps >= len are zero-filled (0 is the identity of the scan).
template <typename T, int kItems, bool kVec, int kThreads>
__device__ __forceinline__ void stage_rows(const T* __restrict__ src, int64_t row_stride, int n0,
                                           int ncnt, int t0, int len, float* __restrict__ dst) {
  using Tile = ScanTile<kItems>;
  if constexpr (kVec) {
    constexpr int VE = Io<T>::kVecElems;
    constexpr int kVecPerRow = Tile::kLen / VE;
    const int total = ncnt * kVecPerRow;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
      const int r = idx / kVecPerRow;
      const int tt = (idx - r * kVecPerRow) * VE;
      float v[VE];
      if (t0 + tt < len) {
        const uint4 raw = ldg16(src + (int64_t)(n0 + r) * row_stride + t0 + tt);
        Io<T>::unpack(raw, v);
      } else {
#pragma unroll
        for (int j = 0; j < VE; ++j) v[j] = 0.f;
      }
      float* d = dst + r * Tile::kRow + (tt / kItems) * Tile::kSeg + (tt % kItems);
#pragma unroll
      for (int j = 0; j < VE; j += 4)
        *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  } else {
    const int total = ncnt * Tile::kLen;
    for (int idx = threadIdx.x; idx < total; idx += kThreads) {
      const int r = idx / Tile::kLen;
      const int tt = idx - r * Tile::kLen;
      float v = 0.f;
      if (t0 + tt < len) v = Io<T>::to_f(src[(int64_t)(n0 + r) * row_stride + t0 + tt]);
      dst[r * Tile::kRow + (tt / kItems) * Tile::kSeg + (tt % kItems)] = v;
    }
  }
}

// Read this lane's kItems staged values of one dstate row.
template <int kItems>
__device__ __forceinline__ void lane_row(const float* __restrict__ row_lane, float* out) {
#pragma unroll
  for (int j = 0; j < kItems; j += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row_lane + j);
    out[j] = v.x;
    out[j + 1] = v.y;
    out[j + 2] = v.z;
    out[j + 3] = v.w;
  }
}

// Inclusive warp scan (lane 0 first) of affine maps s -> P*s + h.
__device__ __forceinline__ void warp_scan_affine_up(float& P, float& h, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float Pp = __shfl_up_sync(0xffffffffu, P, off);
    const float hp = __shfl_up_sync(0xffffffffu, h, off);
    if (lane >= off) {
      h = fmaf(P, hp, h);
      P *= Pp;
    }
  }
}
// Same, running from lane 31 down (for the reverse-time recurrence of the backward pass).
__device__ __forceinline__ void warp_scan_affine_down(float& P, float& g, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float Pn = __shfl_down_sync(0xffffffffu, P, off);
    const float gn = __shfl_down_sync(0xffffffffu, g, off);
    if (lane + off < 32) {
      g = fmaf(P, gn, g);
      P *= Pn;
    }
  }
}

// True when (ptr, strides, seqlen) allow 16-byte vector access along the sequence.
template <typename T>
inline bool vec_ok(const void* p, int64_t s0, int64_t s1, int seqlen) {
  constexpr int VE = Io<T>::kVecElems;
  return p == nullptr || (aligned16(p) && s0 % VE == 0 && s1 % VE == 0 && seqlen % VE == 0);
}

}  // namespace mtts
