// Forward kernel of the Transformer-XL rel-pos attention (attn_rel.cu instantiates kRel = true; the plain MHSA forward has its own
// kernel in attn_fwd_plain.cuh and the kRel = false instantiation of this template is kept as its shared-memory-P variant), head_dim 64, bf16.
//
// CTA = one (plain: two) 128-query tile(s) of one (clip, head); it walks the 128-key tiles.  Per query tile:
//   warps 0-7  softmax.  Warp w owns TMEM lane quarter wq = w & 3 (query rows 32 wq ..) and the column half g = w >> 2 of every
//              score tile.  The two halves of a row are INDEPENDENT online-softmax streams: each keeps its own running maximum
//              m_g and sum l_g and its own output accumulator O_g (TMEM), so the two warps of a row never talk inside the loop;
//              the halves are merged once, in the epilogue: O = (O_a 2^(m_a - M) + O_b 2^(m_b - M)) / (l_a 2^(m_a - M) + ...).
//              A thread reads its 64 scores from TMEM exactly once, keeps them in registers (max, exponentials, row sum with
//              packed f32x2 math) and writes P (bf16) into the SWIZZLE_128B operand tile.
//   +1 warp    TMA producer (Q once, K / V rings [, position window]).
//   +1 warp    tcgen05.mma issuer: S_{j+1} = Q K_{j+1}^T is issued as soon as the softmax warps hold S_j in registers, so it runs
//              under tile j's exponentials; O_a += P[:, :64] V[:64], O_b += P[:, 64:] V[64:] accumulate IN TMEM.
//   Rescaling is lazy: a row's reference maximum only moves when the new maximum exceeds it by more than 2^8 (exactness is not
//   affected: the reference cancels in O / l); the O_g rows of a warp are then rescaled in place (tcgen05.ld / .st).  With the
//   accumulator in TMEM there is no per-tile fold of O through registers.
//   kRel: next to AC = (q+u) K^T the tensor core forms BD = (q+v) Pw^T against the 256-row position window of the tile; rel_shift
//   is a per-row skew BD_shifted[r, c] = BD[r, 127 - r + c]: the multiples of 8 of the per-row offset are resolved by register selects,
//   the rest by parking 40 window columns of the row in a private, bank-conflict-free shared-memory row and reading 32 of them back at
//   the lane's offset (twice per tile).
//   plain: the CTA carries TWO query tiles (two groups of 8 softmax warps, K / V tiles shared) and each group's P tile is double
//   buffered, so a group never waits for its own P V product: while P_j V_j runs, the group is already in tile j+1's exponentials,
//   and the other group keeps the MUFU pipe busy during this group's TMEM loads (r2a ncu: with one P tile per group the softmax
//   warps spent 30 % of their time outside the exponential phase, in lock step, MUFU 48 % busy).  512 TMEM columns, 1 CTA / SM.
//   kRel: one query tile per CTA (the 256-column BD product fills the other half of TMEM), one P tile.
#pragma once
#include "attn_common.cuh"

namespace t4s {
namespace attn {
namespace fwd2 {

constexpr int kScrPitch = 44;                       // floats per parked row (40 used)
constexpr int kWarpScratch = (32 * kScrPitch + 24) * 4 + 32;   // 5760 B: 32 rows + the group shifts, padded to 64 bytes
constexpr float kRescaleThreshold = 8.0f;           // log2 units

template <bool kRel>
struct Layout {
  static constexpr int kGroups = kRel ? 1 : 2;                   // query tiles per CTA (8 softmax warps each)
  static constexpr bool kPT = !kRel;                             // P lives in TMEM (over the S columns it was computed from)
  static constexpr int kThreads = (8 * kGroups + 2) * 32;
  static constexpr int oQ = 0;                                   // plain: Q tile of group 0, 1;  rel: QU, QV
  static constexpr int oK = oQ + 2 * kTileBytes;
  static constexpr int kVStages = kRel ? 1 : 2;
  static constexpr int oV = oK + 2 * kTileBytes;
  static constexpr int oPw = oV + kVStages * kTileBytes;         // rel only: 256 x 64 bf16 window
  static constexpr int oP = kRel ? oPw + 32768 : oPw;            // rel: the P operand tile;  plain: [group][half][row] (m, l) exchange
  static constexpr int oScr = oP + (kRel ? kPBytes : kGroups * 2048);   // rel only
  static constexpr int oBar = kRel ? oScr + 8 * kWarpScratch : oScr;
  static constexpr int kSmem = oBar + 256;
  static constexpr int kTmemCols = 512;   // per group g: S [256 g, +128)  O_a [256 g + 128, +64)  O_b [256 g + 192, +64);  rel: BD [256, 512)
};
static_assert(Layout<false>::kSmem <= 232448 && Layout<true>::kSmem <= 232448, "attention forward: shared memory");
enum { bQFull = 0, bKFull = 1, bKEmpty = 3, bVFull = 5, bVEmpty = 7, bPwFull = 9, bPwEmpty = 10, bSFull = 11, bSFree = 13, bPFull = 15 /* [group] */,
       bOFull = 17 /* rel: [tile parity];  plain: [group], committed once, after the last tile */, bCount = 19 };

constexpr uint32_t kIdescBD = ptx::umma_idesc(1, 128, 256, 0, 0);

struct Maps {
  CUtensorMap q, qv, k, v, pos;   // plain: q, k, v
};

template <bool kRel>
__global__ void __launch_bounds__(Layout<kRel>::kThreads, 1)
attn_fwd2_kernel(const __grid_constant__ Maps tm, const Args a) {
  using L = Layout<kRel>;
  constexpr int kGroups = L::kGroups;
  constexpr bool kPT = L::kPT;
  constexpr int kProducer = 8 * kGroups, kIssuer = kProducer + 1;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (kGroups * kTile), h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = a.n_tiles;
  const int groups = (kGroups == 2 && q0 + kTile < a.N) ? 2 : 1;   // the second query tile may lie entirely beyond N

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("t4s attn_fwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::mbar_init(&bars[bQFull], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars[bKFull + i], 1);
      ptx::mbar_init(&bars[bKEmpty + i], 1);
      ptx::mbar_init(&bars[bVFull + i], 1);
      ptx::mbar_init(&bars[bVEmpty + i], 1);
      ptx::mbar_init(&bars[bSFull + i], 1);
      ptx::mbar_init(&bars[bSFree + i], 8);
      ptx::mbar_init(&bars[bPFull + i], 8);
      ptx::mbar_init(&bars[bOFull + i], 1);
    }
    ptx::mbar_init(&bars[bPwFull], 1);
    ptx::mbar_init(&bars[bPwEmpty], 1);
    ptx::fence_barrier_init();
  }
  if (warp == kProducer && ptx::elect_one()) {
    ptx::prefetch_tmap(&tm.q);
    ptx::prefetch_tmap(&tm.k);
    ptx::prefetch_tmap(&tm.v);
    if (kRel) {
      ptx::prefetch_tmap(&tm.qv);
      ptx::prefetch_tmap(&tm.pos);
    }
  }
  if (warp == kIssuer) {
    ptx::tmem_alloc(tmem_slot, L::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  T4S_TRACE_AT(kProducer, 15, warp == 0 ? 0 : 7);

  if (warp == kProducer) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      if (kRel) {
        ptx::mbar_arrive_expect_tx(&bars[bQFull], 2 * kTileBytes);
        ptx::tma_load_4d(smem + L::oQ, &tm.q, &bars[bQFull], 0, q0, h, b);
        ptx::tma_load_4d(smem + L::oQ + kTileBytes, &tm.qv, &bars[bQFull], 0, q0, h, b);
      } else {
        ptx::mbar_arrive_expect_tx(&bars[bQFull], groups * kTileBytes);
        for (int g = 0; g < groups; ++g) ptx::tma_load_4d(smem + L::oQ + g * kTileBytes, &tm.q, &bars[bQFull], 0, q0 + g * kTile, h, b);
      }
      auto load_k = [&](int j) {
        const int s = j & 1;
        ptx::mbar_wait(&bars[bKEmpty + s], ((j >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars[bKFull + s], kTileBytes);
        ptx::tma_load_4d(smem + L::oK + s * kTileBytes, &tm.k, &bars[bKFull + s], 0, j * kTile, h, b);
        if (kRel) {
          ptx::mbar_wait(&bars[bPwEmpty], (j & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&bars[bPwFull], 32768);
          ptx::tma_load_4d(smem + L::oPw, &tm.pos, &bars[bPwFull], 0, a.N - kTile - q0 + j * kTile, h, 0);
        }
      };
      load_k(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) load_k(j + 1);
        const int s = kRel ? 0 : (j & 1);
        ptx::mbar_wait(&bars[bVEmpty + s], (kRel ? (j & 1) : ((j >> 1) & 1)) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars[bVFull + s], kTileBytes);
        ptx::tma_load_4d(smem + L::oV + s * kTileBytes, &tm.v, &bars[bVFull + s], 0, j * kTile, h, b);
      }
    }
  } else if (warp == kIssuer) {
    // ---------------- MMA issuer ----------------
    const uint32_t sQ = ptx::smem_u32(smem + L::oQ), sK = ptx::smem_u32(smem + L::oK), sV = ptx::smem_u32(smem + L::oV),
                   sPw = ptx::smem_u32(smem + L::oPw), sP = ptx::smem_u32(smem + L::oP);
    // S_j of group g (rel: and BD_j) -> TMEM.  `last`: this is the last reader of the K slot / window.
    auto issue_s = [&](int j, int g, bool last) {
      const int s = j & 1;
      if (ptx::elect_one()) {
        mma_k64(tmem + 256 * g, sQ + (kRel ? 0 : g * kTileBytes), sK + s * kTileBytes, kIdescS, false);
        if (kRel) {
          mma_k64(tmem + 256, sQ + kTileBytes, sPw, kIdescBD, false);
          ptx::tc_commit(&bars[bPwEmpty]);
        }
        if (last) ptx::tc_commit(&bars[bKEmpty + s]);
        ptx::tc_commit(&bars[bSFull + g]);
      }
      __syncwarp();
    };
    ptx::mbar_wait(&bars[bQFull], 0);
    ptx::mbar_wait(&bars[bKFull], 0);
    if (kRel) ptx::mbar_wait(&bars[bPwFull], 0);
    ptx::tc_fence_after();
    for (int g = 0; g < groups; ++g) issue_s(0, g, g == groups - 1);
    for (int j = 0; j < n_tiles; ++j) {
      const int s = kRel ? 0 : (j & 1);
      if (!kPT && j + 1 < n_tiles) {
        // P is a shared-memory tile: S_{j+1} is issued as soon as the softmax warps hold S_j in registers
        ptx::mbar_wait(&bars[bKFull + ((j + 1) & 1)], ((j + 1) >> 1) & 1);
        if (kRel) ptx::mbar_wait(&bars[bPwFull], (j + 1) & 1);
        ptx::mbar_wait(&bars[bSFree], j & 1);
        ptx::tc_fence_after();
        issue_s(j + 1, 0, true);
      }
      ptx::mbar_wait(&bars[bVFull + s], kRel ? (j & 1) : ((j >> 1) & 1));
      for (int g = 0; g < groups; ++g) {
        T4S_TRACE_AT(kIssuer, j, 4 * g);
        ptx::mbar_wait(&bars[bPFull + g], j & 1);
        T4S_TRACE_AT(kIssuer, j, 4 * g + 1);
        if (kPT && j + 1 < n_tiles && g == 0) ptx::mbar_wait(&bars[bKFull + ((j + 1) & 1)], ((j + 1) >> 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // O_a (+)= P[:, 0:64] V[0:64, :],  O_b (+)= P[:, 64:128] V[64:128, :]   (V consumed MN-major)
          const uint64_t bdesc = ptx::umma_desc_sw128(sV + s * kTileBytes, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t d = tmem + 256 * g + 128 + 64 * (k >> 2);
            const uint32_t acc = (j > 0 || (k & 3) > 0) ? 1u : 0u;
            if (kPT) {
              // A = P from TMEM: half h = k >> 2 wrote its 64 keys as 32 packed columns over its own S columns [64 h, 64 h + 32)
              ptx::mma_f16_ts(d, tmem + 256 * g + 64 * (k >> 2) + 8 * (k & 3), bdesc + 128 * k, kIdescPV, acc);
            } else {
              const uint64_t adesc = ptx::umma_desc_sw128(sP + (k >> 2) * kTileBytes, 16, 1024) + 2 * (k & 3);
              ptx::mma_f16(d, adesc, bdesc + 128 * k, kIdescPV, acc);
            }
          }
          if (g == groups - 1) ptx::tc_commit(&bars[bVEmpty + s]);
          if (!kPT) ptx::tc_commit(&bars[bOFull + (j & 1)]);
          if (kPT && j + 1 == n_tiles) ptx::tc_commit(&bars[bOFull + g]);
        }
        __syncwarp();
        // P in TMEM: the tensor pipe runs MMAs in issue order, so S_{j+1} may follow the product that reads P_j out of the same columns
        if (kPT && j + 1 < n_tiles) issue_s(j + 1, g, g == groups - 1);
        T4S_TRACE_AT(kIssuer, j, 4 * g + 2);
      }
    }
  } else if ((warp >> 3) < groups) {
    // ---------------- softmax warps: thread = (query row, column half) of one group ----------------
    const int gq = warp >> 3, wq = warp & 3, g = (warp >> 2) & 1;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem + 256 * gq + ((uint32_t)(wq * 32) << 16);
    const uint32_t t_o = t_lane + 128 + 64 * g;
    uint64_t* o_full = &bars[bOFull];                      // rel: [tile parity]
    unsigned char* p_tile = smem + L::oP;                  // rel: the P operand tile
    const float sl2 = a.sl2;
    const uint64_t sl2_2 = ptx::pack2(sl2, sl2);
    float m = -INFINITY, l = 0.f;
    // scratch row of this lane: 40 floats at pitch 44, lane group l >> 3 shifted by 8 words each (16-byte stores and skewed 4-byte reads
    // are bank conflict free)
    float* scr = kRel ? reinterpret_cast<float*>(smem + L::oScr + (warp & 7) * kWarpScratch) + lane * kScrPitch + 8 * (lane >> 3) : nullptr;
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;

    for (int j = 0; j < n_tiles; ++j) {
      const int nvalid = a.N - j * kTile - 64 * g;   // columns of this half that exist (may be <= 0 or >= 64)
      T4S_TRACE_AT(warp, j, 0);
      ptx::mbar_wait(&bars[bSFull + gq], j & 1);
      ptx::tc_fence_after();
      T4S_TRACE_AT(warp, j, 1);
      float s[64];
      {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32(t_lane + 64 * g, v0);
        ptx::tmem_ld_32x32(t_lane + 64 * g + 32, v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = __uint_as_float(v0[i]);
          s[32 + i] = __uint_as_float(v1[i]);
        }
      }
      T4S_TRACE_AT(warp, j, 2);
      if (kRel) {
        // s[c] += BD[r][127 - r + 64 g + c]: two sub-chunks of 32 columns, each from a 64-column window of this warp's rows
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int wb = 96 - 32 * wq + 64 * g + 32 * q;
          uint32_t x0[32], x1[32];
          ptx::tmem_ld_32x32(t_lane + 256 + wb, x0);
          ptx::tmem_ld_32x32(t_lane + 256 + wb + 32, x1);
          ptx::tmem_ld_wait();
          // window offset of lane l = 31 - l = 8 (3 - (l >> 3)) + 7 - (l & 7): the multiples of 8 are resolved by register selects on lane
          // bits 4 and 3, so only 40 of the 64 window columns go through shared memory (10 STS.128 + 32 LDS instead of 16 + 32)
          uint32_t a16[48], y[40];
#pragma unroll
          for (int k = 0; k < 48; ++k) a16[k] = b4 ? (k < 32 ? x0[k] : x1[k - 32]) : (k < 16 ? x0[k + 16] : x1[k - 16]);
#pragma unroll
          for (int k = 0; k < 40; ++k) y[k] = b3 ? a16[k] : a16[k + 8];
#pragma unroll
          for (int k = 0; k < 10; ++k) *reinterpret_cast<uint4*>(scr + 4 * k) = make_uint4(y[4 * k], y[4 * k + 1], y[4 * k + 2], y[4 * k + 3]);
          const volatile float* rd = scr + (7 - (lane & 7));
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) s[32 * q + cc] += rd[cc];
        }
      }
      if (!kPT) {
        // S_j (and BD_j) are in registers: the tensor core may overwrite them with tile j+1
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars[bSFree + gq]);
      }
      if (nvalid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= nvalid) s[i] = -INFINITY;
      }
      float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
      for (int i = 4; i < 64; i += 4) {
        mx0 = ptx::max3(mx0, s[i], s[i + 1]);
        mx1 = ptx::max3(mx1, s[i + 2], s[i + 3]);
      }
      const float mx = fmaxf(mx0, mx1);
      if (j > 0) {
        // lazy rescale: only when a row's maximum grew by more than 2^8 (m = -inf: a half that was fully masked so far)
        const bool need = (mx - m) * sl2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // P_{j-1} V_{j-1} must have retired before O_g is touched.  P in TMEM: S_j was issued behind it, so it has.
          if (!kPT) {
            ptx::mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
            ptx::tc_fence_after();
          }
          const float alpha = need ? ex2((m - mx) * sl2) : 1.f;
          if (need) m = mx;
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            uint32_t o[8];
            ptx::tmem_ld_32x8(t_o + 8 * c, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st_32x8(t_o + 8 * c, o);
          }
          ptx::tmem_st_wait();
        }
      } else {
        m = mx;
      }
      // rel: the P tile was last read by the P V product of tile j - 1
      if (!kPT && j >= 1) ptx::mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
      T4S_TRACE_AT(warp, j, 3);
      const float mneg = (m == -INFINITY) ? 0.f : -m * sl2;
      const uint64_t mneg2 = ptx::pack2(mneg, mneg);
      uint64_t rs2 = ptx::pack2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float t0, t1;
        ptx::unpack2(ptx::fma2(ptx::pack2(s[2 * i], s[2 * i + 1]), sl2_2, mneg2), t0, t1);
        const float p0 = ex2(t0), p1 = ex2(t1);
        rs2 = ptx::add2(rs2, ptx::pack2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
      }
      float r0, r1;
      ptx::unpack2(rs2, r0, r1);
      l += r0 + r1;
      T4S_TRACE_AT(warp, j, 4);
      if (kPT) {
        // P_j (bf16 pairs) over this thread's own S columns: 64 keys -> 32 columns at [64 g, 64 g + 32)
        ptx::tmem_st_32x32(t_lane + 64 * g, pk);
        ptx::tmem_st_wait();
      } else {
        uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[0]);
        uint32_t (&hi)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[16]);
        store_row_chunk(p_tile, r, 64 * g, lo);
        store_row_chunk(p_tile, r, 64 * g + 32, hi);
        ptx::fence_proxy_async();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bPFull + gq]);
      T4S_TRACE_AT(warp, j, 5);
    }

    T4S_TRACE_AT(warp, 15, 0);
    // ---- epilogue: merge the two halves of every row ----
    if (kPT) ptx::mbar_wait(&bars[bOFull + gq], 0);
    else ptx::mbar_wait(&o_full[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
    ptx::tc_fence_after();
    float2* stat = reinterpret_cast<float2*>(kPT ? smem + L::oP + gq * 2048 : p_tile);   // rel: the P tile is free now.  [half][row] (m, l)
    stat[g * kTile + r] = make_float2(m, l);
    ptx::bar_sync(1 + 4 * gq + wq, 64);
    const float2 other = stat[(g ^ 1) * kTile + r];
    const float M = fmaxf(m, other.x);
    const float w_me = ex2((m - M) * sl2), w_ot = ex2((other.x - M) * sl2);
    const float Lsum = fmaf(l, w_me, other.y * w_ot);
    const float inv = 1.f / Lsum;
    const float wa = (g == 0 ? w_me : w_ot) * inv, wb = (g == 0 ? w_ot : w_me) * inv;
    uint32_t oa[32], ob[32];
    ptx::tmem_ld_32x32(t_lane + 128 + 32 * g, oa);        // O_a columns [32 g, 32 g + 32)
    ptx::tmem_ld_32x32(t_lane + 192 + 32 * g, ob);        // O_b same columns
    ptx::tmem_ld_wait();
    const int row = q0 + gq * kTile + r;
    if (g == 0) a.lse[((long long)b * a.H + h) * a.Nl + row] = fmaf(M, sl2, log2f(Lsum));
    if (row < a.N) {
      float out[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) out[i] = fmaf(__uint_as_float(oa[i]), wa, __uint_as_float(ob[i]) * wb);
      __nv_bfloat16* dst = a.o + (long long)b * a.o_bs + (long long)row * a.o_ld + h * kHd + 32 * g;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16(out[8 * q], out[8 * q + 1]);
        u.y = pack_bf16(out[8 * q + 2], out[8 * q + 3]);
        u.z = pack_bf16(out[8 * q + 4], out[8 * q + 5]);
        u.w = pack_bf16(out[8 * q + 6], out[8 * q + 7]);
        reinterpret_cast<uint4*>(dst)[q] = u;
      }
      if (a.o32) {
        float* d32 = a.o32 + ((long long)b * a.N + row) * ((long long)a.H * kHd) + h * kHd + 32 * g;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<float4*>(d32)[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
      }
    }
  }

  T4S_TRACE_AT(warp, 15, 1);
  ptx::tc_fence_before();
  __syncthreads();
  T4S_TRACE_AT(warp, 15, 2);
  if (warp == kIssuer) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, L::kTmemCols);
  }
}

}  // namespace fwd2
}  // namespace attn
}  // namespace t4s
