// Fused backward of the training cross-attention core (mamba_decoder.py:72-77 under loss.backward(), train.py:231)
// on the 5th-generation tensor cores: S, P, dP and dS never leave the SM.
//
//   given   q (B, T, E), k | v (B, Tk, 2E), o and d_o (B, T, E) bf16 (head h = columns [64 h, 64 h + 64)),
//           lse2 (B, H, T) = base-2 log-sum-exp of every score row (the forward's softmax GEMM writes it)
//   P  = exp2(scale log2(e) q k^T - lse2)   (masked / padded keys: 0)        delta = rowsum(d_o o)
//   dS = P o (d_o v^T - delta)        dV = P^T d_o        dK = scale dS^T q        dQ = scale dS k
//
// One CTA per (batch, head).  Everything is computed TRANSPOSED (keys on the TMEM lanes, queries along the columns):
// with the row statistics of the forward at hand the softmax backward is elementwise, and with keys on the lanes the
// statistics are per-COLUMN broadcasts (LDS.128 from a 1 KB table) instead of cross-lane reductions.  The 256 keys
// are taken in two halves of 128 (one M = 128 MMA); per half the CTA walks the query tiles of 128:
//   MMA   S^T = K_h Q_i^T,  dP^T = V_h dO_i^T                       -> TMEM columns [0, 128), [128, 256)
//   warps P^T, dS^T (bf16) -> shared memory, in the K-major swizzle-128B operand layout
//   MMA   dV_h += P^T dO_i,  dK_h += dS^T Q_i   (accumulated in TMEM across the query tiles, columns [256, 384)),
//         dQ_i(h) = dS K_h                       (columns [384, 448); the SAME shared-memory bytes of dS^T serve as
//                                                 the MN-major A operand, Q_i / dO_i / K_h as MN-major B operands)
//   warps dQ_i: half 0 stores it, half 1 adds its own on top;  after the last tile: dK_h, dV_h -> global.
// warp 0 = TMA producer (K_h / V_h once per half, Q_i / dO_i / O_i double buffered), warp 1 = MMA issuer, warps 2-9 =
// elementwise + epilogue (thread = one key row of S^T / dP^T, or one query row of dQ).  HBM traffic per layer: q, o, d_o
// read twice, k / v once, dq written (twice with two halves), dk / dv once -- against five passes over the
// (B, H, T, Tk) probability / score-gradient tensors of the GEMM-by-GEMM path.
#include "tc_ptx.cuh"

namespace mtts {
namespace g100 {

namespace attn {
constexpr int DH = 64;            // head dimension
constexpr int TQ = 128;           // query tile
constexpr int TKH = 128;          // keys per half
constexpr int kEpiWarps = 16;     // 4 per TMEM lane group: each thread owns 32 query columns of S^T / dP^T (16 of dQ / dK / dV)
constexpr int kParts = kEpiWarps / 4;
constexpr int kSC = TQ / kParts;  // S^T / dP^T columns per thread
constexpr int kOC = DH / kParts;  // dQ / dK / dV columns per thread
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kTileBytes = 128 * 128;                 // a 128-row x 64-element bf16 operand tile
constexpr uint32_t kStageBytes = 3 * kTileBytes;           // Q_i, dO_i, O_i
// shared memory: K_h, V_h | 2 stages | P^T, dS^T (2 k-blocks each) | statistics [2][2][128] | barriers
constexpr size_t kSmemBytes = 2 * kTileBytes + 2 * kStageBytes + 4 * kTileBytes + 2 * 2 * TQ * 4 + 24 * 8 + 16;
constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 320, kColDQ = 384;   // dQ: two buffers of 64 columns

struct Args {
  int B, H, T, Tk, E;
  const float* lse2;            // (B, H, T)
  const unsigned char* mask;    // (B, Tk) or nullptr
  __nv_bfloat16* dq;            // (B, T, E)
  __nv_bfloat16* dkv;           // (B, Tk, 2E)
  float scale;
};
}  // namespace attn

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// release arrive: orders this warp's shared-memory stores (made visible to the async proxy by the fence above it)
// before the MMA thread that waits on the barrier
__device__ __forceinline__ void mbar_arrive_release(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_sync() {   // the epilogue warps
  asm volatile("bar.sync 1, %0;" ::"n"(32 * attn::kEpiWarps) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns, no wait
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld_tied16(uint32_t* a, uint32_t* b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                 "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}

__global__ void __launch_bounds__(attn::kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const attn::Args g) {
  using namespace attn;
  extern __shared__ __align__(1024) unsigned char asm_raw[];
  unsigned char* sm = asm_raw;
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  unsigned char* sK = sm;
  unsigned char* sV = sK + kTileBytes;
  unsigned char* sStage = sV + kTileBytes;                       // [2][Q | dO | O]
  unsigned char* sPT = sStage + 2 * kStageBytes;                 // [2 k-blocks of 64 q][128 key rows][128 B]
  unsigned char* sDST = sPT + 2 * kTileBytes;
  float* sStat = reinterpret_cast<float*>(sDST + 2 * kTileBytes);   // [stage][lse2 | delta][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * 2 * TQ);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* q_full = bars + 2;      // [2]
  uint64_t* q_empty = bars + 4;     // [2]
  uint64_t* s_ready = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* dq_ready = bars + 12;   // [2]: dQ is double buffered -- the warps read tile it - 1 after the elementwise
                                    // work of tile it, so they never wait for the gradient MMAs
  uint64_t* sdp_free = bars + 9;    // the warps have read S^T / dP^T of the tile (they may be overwritten)
  uint64_t* dq_free = bars + 14;    // [2]
  uint64_t* kv_done = bars + 10;
  uint64_t* acc_free = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / g.H, head = blockIdx.x % g.H;
  const int nq = (g.T + TQ - 1) / TQ;
  const int nh = g.Tk > TKH ? 2 : 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_do)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_o)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_v)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(kv_full, 1);
      mbar_init(kv_empty, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(q_full + s, 1);
        mbar_init(q_empty + s, 1);
      }
      mbar_init(s_ready, 1);
      mbar_init(p_ready, kEpiWarps);
      mbar_init(dq_ready, 1);
      mbar_init(dq_ready + 1, 1);
      mbar_init(sdp_free, kEpiWarps);
      mbar_init(dq_free, kEpiWarps);
      mbar_init(dq_free + 1, kEpiWarps);
      mbar_init(kv_done, 1);
      mbar_init(acc_free, kEpiWarps);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int h = 0; h < nh; ++h) {
        mbar_wait(kv_empty, (uint32_t)((h & 1) ^ 1));            // every MMA of the previous half has retired
        mbar_expect_tx(kv_full, 2 * kTileBytes);
        tma_load_4d(sK, &map_k, kv_full, 0, h * TKH, head, b);
        tma_load_4d(sV, &map_v, kv_full, 0, h * TKH, head, b);
        for (int i = 0; i < nq; ++i, ++it) {
          const int s = it & 1;
          mbar_wait(q_empty + s, (uint32_t)(((it >> 1) & 1) ^ 1));
          unsigned char* st = sStage + (size_t)s * kStageBytes;
          mbar_expect_tx(q_full + s, kStageBytes);
          tma_load_4d(st, &map_q, q_full + s, 0, i * TQ, head, b);
          tma_load_4d(st + kTileBytes, &map_do, q_full + s, 0, i * TQ, head, b);
          tma_load_4d(st + 2 * kTileBytes, &map_o, q_full + s, 0, i * TQ, head, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t id_kk = instr_desc<128, 128, 0, 0>();   // S^T, dP^T: A K-major, B K-major, N = 128
      constexpr uint32_t id_kn = instr_desc<128, 64, 0, 1>();    // dV, dK: A K-major (P^T / dS^T), B MN-major, N = 64
      constexpr uint32_t id_nn = instr_desc<128, 64, 1, 1>();    // dQ: A MN-major (dS), B MN-major (K_h)
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPT = smem_u32(sPT), aDST = smem_u32(sDST);
      // Software pipelined over the (half, tile) iterations: the score MMAs of iteration it + 1 are issued right behind
      // the gradient MMAs of iteration it, so they run while the warps drain dQ(it) -- the warps never wait for S^T.
      const int total = nh * nq;
      auto issue_scores = [&](int it) {                          // S^T = K_h Q_i^T, dP^T = V_h dO_i^T
        const int s = it & 1;
        const uint32_t aQ = smem_u32(sStage + (size_t)s * kStageBytes), aDO = aQ + kTileBytes;
        if (it % nq == 0) {                                      // first tile of a half: its K_h / V_h
          const int h = it / nq;
          mbar_wait(kv_full, (uint32_t)(h & 1));
        }
        mbar_wait(q_full + s, (uint32_t)((it >> 1) & 1));
        mbar_wait(sdp_free, (uint32_t)((it & 1) ^ 1));           // S^T / dP^T of the previous iteration were read
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)                        // K-major operands: +32 B per 16 k
          tc_mma(tmem_base + kColS, umma_desc<0>(aK + 32 * k), umma_desc<0>(aQ + 32 * k), id_kk, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          tc_mma(tmem_base + kColDP, umma_desc<0>(aV + 32 * k), umma_desc<0>(aDO + 32 * k), id_kk, k ? 1u : 0u);
        tc_commit(s_ready);
      };
      issue_scores(0);
      for (int it = 0; it < total; ++it) {
        const int h = it / nq, i = it - h * nq;
        const int s = it & 1;
        const uint32_t aQ = smem_u32(sStage + (size_t)s * kStageBytes), aDO = aQ + kTileBytes;
        if (i == 0) mbar_wait(acc_free, (uint32_t)((h & 1) ^ 1));   // the warps have read the previous half's dK / dV
        mbar_wait(p_ready, (uint32_t)(it & 1));
        mbar_wait(dq_free + (it & 1), (uint32_t)(((it >> 1) & 1) ^ 1));   // this dQ buffer (tile it - 2) was read
        tc_fence_after();
        // contraction over the 128 queries of the tile: A k-block kb (64 q) at +16 KB, +32 B per 16 q inside it;
        // B (dO_i / Q_i read MN-major): +16 rows of 128 B per 16 q
#pragma unroll
        for (int k = 0; k < TQ / 16; ++k) {
          const uint32_t ao = (uint32_t)(k >> 2) * kTileBytes + (uint32_t)(k & 3) * 32;
          tc_mma(tmem_base + kColDV, umma_desc<0>(aPT + ao), umma_desc<1>(aDO + 2048 * k), id_kn, (i | k) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < TQ / 16; ++k) {
          const uint32_t ao = (uint32_t)(k >> 2) * kTileBytes + (uint32_t)(k & 3) * 32;
          tc_mma(tmem_base + kColDK, umma_desc<0>(aDST + ao), umma_desc<1>(aQ + 2048 * k), id_kn, (i | k) ? 1u : 0u);
        }
        // dQ_i = dS K_h: A = the dS^T tile read MN-major (M = 128 q = two 64-q chunks 16 KB apart, 16 key rows of
        // 128 B per step), B = K_h read MN-major
#pragma unroll
        for (int k = 0; k < TKH / 16; ++k) {
          tc_mma(tmem_base + kColDQ + 64 * (it & 1), umma_desc_lbo<1>(aDST + 2048 * k, kTileBytes), umma_desc<1>(aK + 2048 * k), id_nn,
                 k ? 1u : 0u);
        }
        tc_commit(dq_ready + (it & 1));
        tc_commit(q_empty + s);
        if (i == nq - 1) {                                       // last tile of the half
          tc_commit(kv_done);
          tc_commit(kv_empty);
        }
        if (it + 1 < total) issue_scores(it + 1);
      }
    }
  } else {
    // ===== elementwise + epilogue warps =====
    const int ew = warp - 2;
    const int lg = warp & 3;               // TMEM lane group
    const int part = ew >> 2;              // which kSC query columns of S^T / dP^T, which kOC columns of dQ / dK / dV
    const int etid = ew * 32 + lane;       // 0 .. 32 kEpiWarps
    const int row = lg * 32 + lane;        // key row (S^T, dK, dV) or query row (dQ) of the tile
    const float sc2 = g.scale * kLog2e, scale = g.scale;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    // dQ of iteration jt (half jt / nq, tile jt % nq): query row `row`, kOC of the 64 head columns; the second half adds
    // onto the first half's result, requested before the wait for the tensor core
    auto dq_out = [&](int jt) {
      const int hh = jt / nq, qrow = (jt - hh * nq) * TQ + row;
      __nv_bfloat16* dst = g.dq + ((size_t)b * g.T + (qrow < g.T ? qrow : 0)) * g.E + head * DH + part * kOC;
      uint4 oldv[kOC / 8];
      if (hh > 0 && qrow < g.T) {
#pragma unroll
        for (int j = 0; j < kOC / 8; ++j) oldv[j] = *reinterpret_cast<const uint4*>(dst + 8 * j);
      }
      mbar_wait_warp(dq_ready + (jt & 1), (uint32_t)((jt >> 1) & 1), lane);
      tc_fence_after();
      uint32_t qv[16], dummy[16];
      static_assert(kOC == 16, "dQ / dK / dV: 16 columns per thread");
      tmem_ld16_nowait(lane_addr + kColDQ + 64 * (jt & 1) + part * kOC, qv);
      tmem_ld16_nowait(lane_addr + kColDQ + 64 * (jt & 1) + part * kOC, dummy);
      tmem_wait_ld_tied16(qv, dummy);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free + (jt & 1));
      if (qrow < g.T) {
#pragma unroll
        for (int j = 0; j < kOC / 8; ++j) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(qv[8 * j + q]) * scale;
          if (hh > 0) {
            float old[8];
            unpack_bf16x8(oldv[j], old);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] += old[q];
          }
          *reinterpret_cast<uint4*>(dst + 8 * j) =
              make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
      }
    };
    float lse_next = (etid < TQ && etid < g.T) ? g.lse2[((size_t)b * g.H + head) * g.T + etid] : 0.f;
    int it = 0;
    for (int h = 0; h < nh; ++h) {
      const int key = h * TKH + row;
      const bool keep = key < g.Tk && (g.mask == nullptr || g.mask[(size_t)b * g.Tk + key] != 0);
      for (int i = 0; i < nq; ++i, ++it) {
        const int s = it & 1;
        const int q0 = i * TQ;
        float* stat = sStat + s * 2 * TQ;
        const unsigned char* st = sStage + (size_t)s * kStageBytes;
        // (a) lse2 of the tile's 128 queries -> shared memory (one value per thread of the first four warps); the
        // value was requested an iteration ago (a global load here would sit on the critical path of every tile)
        if (etid < TQ) {
          stat[etid] = lse_next;
          const int qn = ((i + 1 < nq) ? (i + 1) : 0) * TQ + etid;       // next tile (or tile 0 of the next half)
          lse_next = (qn < g.T) ? g.lse2[((size_t)b * g.H + head) * g.T + qn] : 0.f;
        }
        // (b) delta = rowsum(dO o O) from the staged tiles: kParts threads per query row, 16-byte chunks dealt round robin
        mbar_wait_warp(q_full + s, (uint32_t)((it >> 1) & 1), lane);
        {
          constexpr int TPR = 32 * kEpiWarps / TQ;                    // threads per row (4)
          const int r = etid / TPR, sub = etid % TPR;
          const unsigned char* rdo = st + kTileBytes + r * 128;
          const unsigned char* ro = st + 2 * kTileBytes + r * 128;
          float acc = 0.f;
#pragma unroll
          for (int c = sub; c < 8; c += TPR) {
            const int chunk = (c ^ (r & 7)) << 4;                     // swizzle-128B: 16-byte chunk index ^ (row & 7)
            float a[8], o8[8];
            unpack_bf16x8(*reinterpret_cast<const uint4*>(rdo + chunk), a);
            unpack_bf16x8(*reinterpret_cast<const uint4*>(ro + chunk), o8);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(a[j], o8[j], acc);
          }
#pragma unroll
          for (int o = 1; o < TPR; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (sub == 0) stat[TQ + r] = acc;
        }
        epi_sync();
        // (c) P^T and dS^T of this thread's key row, kSC query columns, straight into the operand tiles
        mbar_wait_warp(s_ready, (uint32_t)(it & 1), lane);
        tc_fence_after();
        {
          const int cbase = part * kSC;                               // first query column of this thread
          unsigned char* prow = sPT + (size_t)(cbase >> 6) * kTileBytes + row * 128;
          unsigned char* drow = sDST + (size_t)(cbase >> 6) * kTileBytes + row * 128;
#pragma unroll
          for (int c = 0; c < kSC; c += 16) {
            uint32_t sv[16], dv[16];
            tmem_ld16_nowait(lane_addr + kColS + cbase + c, sv);
            tmem_ld16_nowait(lane_addr + kColDP + cbase + c, dv);
            tmem_wait_ld_tied16(sv, dv);
#pragma unroll
            for (int j8 = 0; j8 < 16; j8 += 8) {
              float p[8], d[8];
#pragma unroll
              for (int j4 = 0; j4 < 8; j4 += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(stat + cbase + c + j8 + j4);
                const float4 e4 = *reinterpret_cast<const float4*>(stat + TQ + cbase + c + j8 + j4);
                const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float pj = keep ? ex2f(fmaf(__uint_as_float(sv[j8 + j4 + j]), sc2, -lv[j])) : 0.f;
                  p[j4 + j] = pj;
                  d[j4 + j] = pj * (__uint_as_float(dv[j8 + j4 + j]) - ev[j]);   // (x scale: applied to dQ / dK)
                }
              }
              const int chunk = ((((cbase & 63) + c + j8) >> 3) ^ (row & 7)) << 4;
              *reinterpret_cast<uint4*>(prow + chunk) =
                  make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
              *reinterpret_cast<uint4*>(drow + chunk) =
                  make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
            }
          }
        }
        fence_proxy_async();               // the generic-proxy stores above are read by the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_release(p_ready);
          mbar_arrive(sdp_free);
        }
        // (d) dQ of the PREVIOUS tile (its MMAs were issued when this tile's S^T was being computed: long done)
        if (i > 0) dq_out(it - 1);
      }
      // dK_h, dV_h of this thread's key row (and the dQ of the half's last tile)
      mbar_wait_warp(kv_done, (uint32_t)(h & 1), lane);
      tc_fence_after();
      dq_out(it - 1);
      {
        uint32_t kv[16], vv[16];
        tmem_ld16_nowait(lane_addr + kColDK + part * kOC, kv);
        tmem_ld16_nowait(lane_addr + kColDV + part * kOC, vv);
        tmem_wait_ld_tied16(kv, vv);
        if (key < g.Tk) {
          __nv_bfloat16* dk = g.dkv + ((size_t)b * g.Tk + key) * (2 * g.E) + head * DH + part * kOC;
          __nv_bfloat16* dvp = dk + g.E;
#pragma unroll
          for (int j = 0; j < kOC; j += 8) {
            *reinterpret_cast<uint4*>(dk + j) = make_uint4(
                pack_bf16(__uint_as_float(kv[j]) * scale, __uint_as_float(kv[j + 1]) * scale),
                pack_bf16(__uint_as_float(kv[j + 2]) * scale, __uint_as_float(kv[j + 3]) * scale),
                pack_bf16(__uint_as_float(kv[j + 4]) * scale, __uint_as_float(kv[j + 5]) * scale),
                pack_bf16(__uint_as_float(kv[j + 6]) * scale, __uint_as_float(kv[j + 7]) * scale));
            *reinterpret_cast<uint4*>(dvp + j) = make_uint4(
                pack_bf16(__uint_as_float(vv[j]), __uint_as_float(vv[j + 1])),
                pack_bf16(__uint_as_float(vv[j + 2]), __uint_as_float(vv[j + 3])),
                pack_bf16(__uint_as_float(vv[j + 4]), __uint_as_float(vv[j + 5])),
                pack_bf16(__uint_as_float(vv[j + 6]), __uint_as_float(vv[j + 7])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Forward of the same core: o = softmax(scale q k^T + mask) v and the per-row base-2 log-sum-exp, scores and
// probabilities never leaving the SM.  One CTA per (batch, head): K (K-major B operand of S = Q K^T) and V (the same
// kind of tile, read MN-major as the B operand of O = P V) are loaded once, the CTA walks the query tiles of 128:
//   MMA   S = Q_i K^T (128 x 256)                                  -> TMEM columns [0, 256)
//   warps thread = query row x 64 keys: row max (masked) and, in a second sweep over TMEM, e = exp2(.) -> bf16 into
//         the K-major operand tile of shared memory (UNnormalised: the row sum divides O at the end); max and sum are
//         exchanged between the four warps of a TMEM lane group through shared memory
//   MMA   O = e V (128 x 64, K = 256)                              -> TMEM columns [256, 320)
//   warps o = O / sum -> global, lse2 = base + log2(sum)
// The next tile's score MMA is issued right behind this tile's O MMA.
namespace attnf {
constexpr int DH = 64, TQ = 128, TK = 256;
constexpr int kEpiWarps = 16, kParts = 4, kSC = TK / kParts, kOC = DH / kParts;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kTileBytes = 128 * 128;
// K, V (256 rows each) | 2 Q stages | P (4 k-blocks) | exchange [4 lane groups][4 parts][32][2] | barriers
constexpr size_t kSmemBytes = 4 * kTileBytes + 2 * kTileBytes + 4 * kTileBytes + 4 * 4 * 32 * 2 * 4 + 16 * 8 + 16;
constexpr uint32_t kColS = 0, kColO = 256;
struct Args {
  int B, H, T, Tk, E;
  const unsigned char* mask;
  __nv_bfloat16* o;
  float* lse2;
  float scale;
};
}  // namespace attnf

__global__ void __launch_bounds__(attnf::kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const attnf::Args g) {
  using namespace attnf;
  extern __shared__ __align__(1024) unsigned char asm_raw[];
  unsigned char* sm = asm_raw;
  if ((smem_u32(sm) & 1023u) != 0) __trap();
  unsigned char* sK = sm;
  unsigned char* sV = sK + 2 * kTileBytes;
  unsigned char* sQ = sV + 2 * kTileBytes;                      // [2]
  unsigned char* sP = sQ + 2 * kTileBytes;                      // [4 k-blocks of 64 keys][128 q rows][128 B]
  float* sX = reinterpret_cast<float*>(sP + 4 * kTileBytes);    // [lane group][part][lane][max | sum]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 4 * 4 * 32 * 2);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;      // [2]
  uint64_t* q_empty = bars + 3;     // [2]
  uint64_t* s_ready = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_ready = bars + 7;
  uint64_t* o_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / g.H, head = blockIdx.x % g.H;
  const int nq = (g.T + TQ - 1) / TQ;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_v)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(kv_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(q_full + s, 1);
        mbar_init(q_empty + s, 1);
      }
      mbar_init(s_ready, 1);
      mbar_init(p_ready, kEpiWarps);
      mbar_init(o_ready, 1);
      mbar_init(o_free, kEpiWarps);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 4 * kTileBytes);
      tma_load_4d(sK, &map_k, kv_full, 0, 0, head, b);
      tma_load_4d(sV, &map_v, kv_full, 0, 0, head, b);
      for (int it = 0; it < nq; ++it) {
        const int s = it & 1;
        mbar_wait(q_empty + s, (uint32_t)(((it >> 1) & 1) ^ 1));
        mbar_expect_tx(q_full + s, kTileBytes);
        tma_load_4d(sQ + (size_t)s * kTileBytes, &map_q, q_full + s, 0, it * TQ, head, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_s = instr_desc<128, 256, 0, 0>();    // S: A = Q K-major, B = K K-major, N = 256
      constexpr uint32_t id_o = instr_desc<128, 64, 0, 1>();     // O: A = e K-major, B = V MN-major, N = 64
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      auto issue_s = [&](int it) {
        const int s = it & 1;
        const uint32_t aQ = smem_u32(sQ + (size_t)s * kTileBytes);
        mbar_wait(q_full + s, (uint32_t)((it >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          tc_mma(tmem_base + kColS, umma_desc<0>(aQ + 32 * k), umma_desc<0>(aK + 32 * k), id_s, k ? 1u : 0u);
        tc_commit(s_ready);
        tc_commit(q_empty + s);
      };
      mbar_wait(kv_full, 0);
      issue_s(0);
      for (int it = 0; it < nq; ++it) {
        mbar_wait(p_ready, (uint32_t)(it & 1));                  // e(it) is in shared memory, S(it) was read
        mbar_wait(o_free, (uint32_t)((it & 1) ^ 1));             // O of the previous tile was read
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < TK / 16; ++k) {
          const uint32_t ao = (uint32_t)(k >> 2) * kTileBytes + (uint32_t)(k & 3) * 32;
          tc_mma(tmem_base + kColO, umma_desc<0>(aP + ao), umma_desc<1>(aV + 2048 * k), id_o, k ? 1u : 0u);
        }
        tc_commit(o_ready);
        if (it + 1 < nq) issue_s(it + 1);
      }
    }
  } else {
    const int ew = warp - 2;
    const int lg = warp & 3, part = ew >> 2;
    const int row = lg * 32 + lane;
    const float sc2 = g.scale * kLog2e;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16);
    // keys of this warp: [64 part, 64 part + 64); bit j of keep[c / 32]: key 64 part + c + j takes part
    uint32_t keep[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int key = part * kSC + c * 32 + lane;
      const bool on = key < g.Tk && (g.mask == nullptr || g.mask[(size_t)b * g.Tk + key] != 0);
      keep[c] = __ballot_sync(0xffffffffu, on);
    }
    float* xmine = sX + ((lg * 4 + part) * 32 + lane) * 2;
    auto xother = [&](int q) { return sX + ((lg * 4 + q) * 32 + lane) * 2; };
    auto lg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(lg + 1) : "memory"); };
    unsigned char* prow = sP + (size_t)part * kTileBytes + row * 128;
    for (int it = 0; it < nq; ++it) {
      const int q0 = it * TQ;
      mbar_wait_warp(s_ready, (uint32_t)(it & 1), lane);
      tc_fence_after();
      // sweep 1: row maximum over this warp's keys
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < kSC; c += 32) {
        uint32_t a[16], bq[16];
        tmem_ld16_nowait(lane_addr + kColS + part * kSC + c, a);
        tmem_ld16_nowait(lane_addr + kColS + part * kSC + c + 16, bq);
        tmem_wait_ld_tied16(a, bq);
        const uint32_t kb = keep[c >> 5];
        if (kb == 0xffffffffu) {               // warp-uniform: no masked / padded key among these 32
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            m0 = fmaxf(m0, __uint_as_float(a[j]));
            m1 = fmaxf(m1, __uint_as_float(bq[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if ((kb >> j) & 1u) m0 = fmaxf(m0, __uint_as_float(a[j]));
            if ((kb >> (16 + j)) & 1u) m1 = fmaxf(m1, __uint_as_float(bq[j]));
          }
        }
      }
      float mx = fmaxf(m0, m1);
      xmine[0] = mx;
      lg_sync();
#pragma unroll
      for (int q = 0; q < kParts; ++q) mx = fmaxf(mx, xother(q)[0]);
      const float base = mx == -INFINITY ? 0.f : mx * sc2;       // fully masked row: zeros (the reference gives NaN)
      // sweep 2: e = exp2(scale log2(e) s - base) -> bf16 operand tile, row sum
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < kSC; c += 32) {
        uint32_t a[16], bq[16];
        tmem_ld16_nowait(lane_addr + kColS + part * kSC + c, a);
        tmem_ld16_nowait(lane_addr + kColS + part * kSC + c + 16, bq);
        tmem_wait_ld_tied16(a, bq);
        const uint32_t kb = keep[c >> 5];
        float e[32];
        if (kb == 0xffffffffu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            e[j] = ex2f(fmaf(__uint_as_float(a[j]), sc2, -base));
            e[16 + j] = ex2f(fmaf(__uint_as_float(bq[j]), sc2, -base));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            e[j] = ((kb >> j) & 1u) ? ex2f(fmaf(__uint_as_float(a[j]), sc2, -base)) : 0.f;
            e[16 + j] = ((kb >> (16 + j)) & 1u) ? ex2f(fmaf(__uint_as_float(bq[j]), sc2, -base)) : 0.f;
          }
        }
#pragma unroll
        for (int j8 = 0; j8 < 32; j8 += 8) {
          // the sum is taken over the ROUNDED values: numerator (tensor core) and denominator agree
          uint32_t w[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(e[j8 + 2 * q], e[j8 + 2 * q + 1]);
            w[q] = *reinterpret_cast<const uint32_t*>(&hh);
            s0 += __low2float(hh);
            s1 += __high2float(hh);
          }
          const int chunk = (((c + j8) >> 3) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + chunk) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_release(p_ready);
      float sum = s0 + s1;
      xmine[1] = sum;
      lg_sync();
      sum = 0.f;
#pragma unroll
      for (int q = 0; q < kParts; ++q) sum += xother(q)[1];
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      const int qrow = q0 + row;
      if (part == 0 && qrow < g.T)
        g.lse2[((size_t)b * g.H + head) * g.T + qrow] = sum > 0.f ? base + lg2f(sum) : INFINITY;
      // O of this tile: 16 of the 64 head columns per thread
      mbar_wait_warp(o_ready, (uint32_t)(it & 1), lane);
      tc_fence_after();
      {
        uint32_t ov[16], dummy[16];
        tmem_ld16_nowait(lane_addr + kColO + part * kOC, ov);
        tmem_ld16_nowait(lane_addr + kColO + part * kOC, dummy);
        tmem_wait_ld_tied16(ov, dummy);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        if (qrow < g.T) {
          __nv_bfloat16* dst = g.o + ((size_t)b * g.T + qrow) * g.E + head * DH + part * kOC;
#pragma unroll
          for (int j = 0; j < kOC; j += 8)
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(
                pack_bf16(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv),
                pack_bf16(__uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv),
                pack_bf16(__uint_as_float(ov[j + 4]) * inv, __uint_as_float(ov[j + 5]) * inv),
                pack_bf16(__uint_as_float(ov[j + 6]) * inv, __uint_as_float(ov[j + 7]) * inv));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace g100
}  // namespace mtts

extern "C" int mtts_attn_core_bwd(const mtts_attn_core_bwd_params* p, mtts_stream_t stream) {
  using namespace mtts::g100;
  if (!p || !p->q || !p->kv || !p->o || !p->d_o || !p->lse2 || !p->dq || !p->dkv) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->heads < 1 || p->t_q < 0 || p->t_kv < 1 || p->t_kv > 256 || p->d_model != p->heads * attn::DH)
    return MTTS_ERR_SHAPE;
  if (!(p->scale > 0.f)) return MTTS_ERR_UNSUPPORTED;
  if (p->batch == 0 || p->t_q == 0) return MTTS_OK;
  if (!mtts::aligned16(p->q) || !mtts::aligned16(p->kv) || !mtts::aligned16(p->o) || !mtts::aligned16(p->d_o) ||
      !mtts::aligned16(p->dq) || !mtts::aligned16(p->dkv))
    return MTTS_ERR_ALIGN;
  const int64_t B = p->batch, H = p->heads, T = p->t_q, Tk = p->t_kv, E = p->d_model;
  CUtensorMap mq, mdo, mo, mk, mv;
  // (d0 = head columns, d1 = rows, inner batch = heads, outer batch = batch); boxes of 64 columns x 128 rows
  const bool ok =
      make_map(&mq, p->q, attn::DH, T, E, H, attn::DH, B, T * E, 64, 128) &&
      make_map(&mdo, p->d_o, attn::DH, T, E, H, attn::DH, B, T * E, 64, 128) &&
      make_map(&mo, p->o, attn::DH, T, E, H, attn::DH, B, T * E, 64, 128) &&
      make_map(&mk, p->kv, attn::DH, Tk, 2 * E, H, attn::DH, B, Tk * 2 * E, 64, 128) &&
      make_map(&mv, reinterpret_cast<const unsigned char*>(p->kv) + 2 * E, attn::DH, Tk, 2 * E, H, attn::DH, B,
               Tk * 2 * E, 64, 128);
  if (!ok) return MTTS_ERR_UNSUPPORTED;
  attn::Args a{};
  a.B = (int)B; a.H = (int)H; a.T = (int)T; a.Tk = (int)Tk; a.E = (int)E;
  a.lse2 = p->lse2;
  a.mask = p->mask;
  a.dq = reinterpret_cast<__nv_bfloat16*>(p->dq);
  a.dkv = reinterpret_cast<__nv_bfloat16*>(p->dkv);
  a.scale = p->scale;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)attn::kSmemBytes);
    if (e != cudaSuccess) return -static_cast<int>(e);
    configured = true;
  }
  attn_bwd_kernel<<<(unsigned)(B * H), attn::kThreads, attn::kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      mq, mdo, mo, mk, mv, a);
  return mtts::launch_status();
}

extern "C" int mtts_attn_core_fwd(const mtts_attn_core_fwd_params* p, mtts_stream_t stream) {
  using namespace mtts::g100;
  if (!p || !p->q || !p->kv || !p->o || !p->lse2) return MTTS_ERR_NULL;
  if (p->batch < 0 || p->heads < 1 || p->t_q < 0 || p->t_kv < 1 || p->t_kv > 256 || p->d_model != p->heads * attnf::DH)
    return MTTS_ERR_SHAPE;
  if (!(p->scale > 0.f)) return MTTS_ERR_UNSUPPORTED;
  if (p->batch == 0 || p->t_q == 0) return MTTS_OK;
  if (!mtts::aligned16(p->q) || !mtts::aligned16(p->kv) || !mtts::aligned16(p->o)) return MTTS_ERR_ALIGN;
  const int64_t B = p->batch, H = p->heads, T = p->t_q, Tk = p->t_kv, E = p->d_model;
  CUtensorMap mq, mk, mv;
  const bool ok =
      make_map(&mq, p->q, attnf::DH, T, E, H, attnf::DH, B, T * E, 64, 128) &&
      make_map(&mk, p->kv, attnf::DH, Tk, 2 * E, H, attnf::DH, B, Tk * 2 * E, 64, 256) &&
      make_map(&mv, reinterpret_cast<const unsigned char*>(p->kv) + 2 * E, attnf::DH, Tk, 2 * E, H, attnf::DH, B,
               Tk * 2 * E, 64, 256);
  if (!ok) return MTTS_ERR_UNSUPPORTED;
  attnf::Args a{};
  a.B = (int)B; a.H = (int)H; a.T = (int)T; a.Tk = (int)Tk; a.E = (int)E;
  a.mask = p->mask;
  a.o = reinterpret_cast<__nv_bfloat16*>(p->o);
  a.lse2 = p->lse2;
  a.scale = p->scale;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)attnf::kSmemBytes);
    if (e != cudaSuccess) return -static_cast<int>(e);
    configured = true;
  }
  attn_fwd_kernel<<<(unsigned)(B * H), attnf::kThreads, attnf::kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      mq, mk, mv, a);
  return mtts::launch_status();
}
