// K5: fused Transformer-XL relative-position attention (transformerXL.py:299-593, rel_shift :254-297), forward and backward.
//
//   score[i, j] = ((q_i + u) . k_j + (q_i + v) . p[T-1-i+j]) * scale          p = linear_pos(pos_emb)  [2T-1, H*64]
//
// The [T, 2T-1] position-score matrix and its rel_shift never exist in memory.  Per 128 x 128 score tile (queries i0.., keys j0..)
// the tensor core computes, next to AC = QU K^T, the product BD = QV Pw^T against the 256-row window Pw = p[T-128-i0+j0 ...] that
// holds every position row the tile can reference (TMA zero-fills the rows that fall outside the table); the shift is then a
// per-row skew, BD_shifted[r, c] = BD[r, 127 - r + c], applied by the softmax warps on the way out of TMEM: each thread (= query
// row) parks 64 accumulator columns in a private, bank-conflict-free shared-memory row and reads them back at its own offset.
//
//   forward    attn_fwd.cuh (kRel): 8 softmax warps, thread = (query row, column half); the two halves of a row are independent
//              online-softmax streams with their own O accumulators in TMEM, merged in the epilogue; lazy rescaling; the skewed
//              position scores are added to the AC scores in registers.  One CTA per SM (512 TMEM columns).
//   backward   two kernels, both with 8 softmax warps and thread = (query row, column half): dQ (CTA = query tile, loops over
//              key tiles) accumulates d(q+u) = scale dS K in TMEM and streams dS, un-shifted back to position coordinates, into
//              dBD: every warp stages its 32 x 64 dS block in shared memory and writes it out with coalesced 4-byte stores, a
//              funnel shift absorbing the odd element offsets (TMA tile stores cannot: they need 16-byte aligned inner
//              coordinates).  The [T, 2T-1] gradient is consumed by two plain GEMMs: d(q+v) = dBD p, dp = sum_b dBD^T (q+v);
//              dK/dV (CTA = key tile, loops over query tiles) consumes the P / dS tiles transposed in place as MN-major A
//              operands.  P is recomputed from lse (no max pass); dP reuses the S columns of TMEM once P is in registers.
//              (Keeping dBD out of HBM needs a 64 KB un-skewed dS operand tile next to 112 KB of operands, 32 KB of dS and the
//              skew scratch: it does not fit the 227 KB of one SM with 128-wide tiles; DESIGN.md §3 has the budget.)
#include "attn_common.cuh"
#include "attn_fwd.cuh"

namespace t4s {
namespace attn {
namespace rel {

constexpr uint32_t kIdescAmn = ptx::umma_idesc(1, 128, 64, 1, 1);  // A and B MN-major

struct RelArgs {
  Args a;
  __nv_bfloat16* dqu; long long dqu_ld, dqu_bs;
  __nv_bfloat16* dbd; long long dbd_ld;
};

// D[128 x 64] (+)= A^T . B with A = a [128 (K) x 128 (M)] K-major-written tile consumed MN-major (two 64-wide M blocks 16 KB
// apart) and B = [128 rows (K) x 64] tile consumed MN-major.
__device__ __forceinline__ void mma_k128_amn(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool accumulate) {
  const uint64_t adesc = ptx::umma_desc_sw128(a_addr, kTileBytes, 1024), bdesc = ptx::umma_desc_sw128(b_addr, 8192, 1024);
#pragma unroll
  for (int k = 0; k < 8; ++k) ptx::mma_f16(d_tmem, adesc + 128 * k, bdesc + 128 * k, kIdescAmn, (accumulate || k > 0) ? 1u : 0u);
}

// ======================================================================================================
// backward
// ======================================================================================================
namespace bwd {
// 10 warps: 0-7 softmax (thread = query row 32 (w & 3) + lane, column half g = w >> 2 of the 128-key tile), 8 TMA producer, 9 MMA issuer.
// Operand region, nine 16 KB tiles:
//   dQ  kernel: resident [QU][QV][dO] | ring [K][K'] | [V]  |        position chunks [Pw0][Pw1][Pw2]
//   dKV kernel: resident [K][V]       | ring [QU][QU'] | [dO] | [QV] | position chunks [Pw0][Pw1][Pw2]
// The ring operand is needed from the first to the last MMA of a step, so it is double-buffered; the others are re-filled inside the
// step, as soon as the one MMA that reads them has retired.  The 256-row position window of step t is two 128-row chunks; consecutive
// windows share one, so one chunk is loaded per step into a ring of three (BD = two N = 128 MMAs).
// Pipeline of step t:   [S(t), BD(t) were issued during step t-1]  phase A: P = exp2(AC + shift(BD) - lse)  ->  dP(t), S(t+1)  ->
// phase B: dS = P (dP - delta)  [BD(t+1) once dP is in registers]  ->  dQ += dS K  |  dV += P^T dO, dK += dS^T QU.
// TMEM: S [0,128)   BD [128,384), its first half re-used for dP once the skew has consumed it   accumulators [384,512).
// Skew scratch: a thread needs BD[r][127 - r + c] for the 16 columns c of a window; the coarse part of the per-row offset (multiples of 8) is
// resolved by register selects, the fine part by parking 24 columns in a private shared-memory row (pitch 26 floats, lanes 16-31 shifted by
// 8 words: 8-byte stores and 4-byte skewed reads are bank conflict free) and reading 16 back at the lane's offset, four times per tile.  The
// rows live in the warp's own 4 KB block of the P tile, which is dead between the dV MMA of step t-1 and the P store of step t (in the dQ
// kernel the P tile is not an operand: its place is the dBD staging area, below).  The mbarriers sit in the first 32 bytes of the 2432-byte
// slots behind the dS tile (the slots held scratch rows 20-31 while a row was 48 columns wide).
// dBD (dQ kernel): dS un-shifted back to position coordinates, dBD[i, T-1-i+j] = dS[i, j].  The two warps of a lane quarter stage their 32
// rows x 128 columns row-contiguously (256 B per row, 16-byte chunks XOR-swizzled by the row) and each writes 16 whole rows.  A row whose
// first destination column is odd is written one element to the left, the missing element being the last one of the same row from the
// previous key tile (kept in a register), so that every store is a full 4-byte word: two-lane half-word stores at the ends of the
// rows cost more than all the word stores together (measured: 4.4 k clk per step with them, 1.7 k without).
constexpr int kBThreads = 320;
constexpr int kPitch = 26;
constexpr int kScrSlot = 2432;
constexpr int oOps = 0, oP = oOps + 9 * kTileBytes, oDs = oP + kPBytes, oScr = oDs + kPBytes;
constexpr int kSmem = oScr + 8 * kScrSlot;
static_assert(kSmem <= 232448, "rel-pos attention backward: shared memory");
constexpr int kTmemCols = 512;
enum { bResFull = 0, bRingFull = 1, bRingEmpty = 3, bXFull = 5, bXEmpty = 6, bYFull = 7, bYEmpty = 8, bPwFull = 9, bPwEmpty = 12, bSFull = 15,
       bPReady = 16, bDpFull = 17, bDpRead = 18, bDsFull = 19, bFin = 20, bAccFull = 21, bPFree = 22, bCount = 23 };

}  // namespace bwd

template <bool kDq>
__global__ void __launch_bounds__(bwd::kBThreads, 1)
relattn_bwd_kernel(const __grid_constant__ CUtensorMap tmQU, const __grid_constant__ CUtensorMap tmQV,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmPos, const RelArgs ra) {
  using namespace bwd;
  const Args& a = ra.a;
  extern __shared__ __align__(1024) unsigned char smem[];
  auto bar = [&](int i) { return reinterpret_cast<uint64_t*>(smem + oScr + (i >> 2) * kScrSlot + (i & 3) * 8); };   // the slots' leading holes
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar(bCount));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = a.n_tiles;
  const long long stat_base = ((long long)b * a.H + h) * a.Nl;
  // operand tiles
  unsigned char* sRes = smem + oOps;                                   // dQ: QU, QV, dO     dKV: K, V
  unsigned char* sRing = smem + oOps + (kDq ? 3 : 2) * kTileBytes;     // dQ: K              dKV: QU
  unsigned char* sX = smem + oOps + (kDq ? 5 : 4) * kTileBytes;        // dQ: V              dKV: dO
  unsigned char* sY = smem + oOps + 5 * kTileBytes;                    //                    dKV: QV
  unsigned char* sPw = smem + oOps + 6 * kTileBytes;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("t4s relattn_bwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < bCount; ++i) ptx::mbar_init(bar(i), (i == bPReady || i == bDpRead || i == bDsFull) ? 8 : 1);
    ptx::fence_barrier_init();
  }
  if (warp == 8 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmQU);
    ptx::prefetch_tmap(&tmQV);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::prefetch_tmap(&tmDO);
    ptx::prefetch_tmap(&tmPos);
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    // Position chunk n (n = 0 .. n_tiles) goes to slot n % 3; step t reads chunks t and t + 1.
    //   dQ  kernel (query tile fixed, keys advance): chunk n = rows T - 128 - i0 + 128 n, (lower, upper) half of the window = (t, t + 1)
    //   dKV kernel (key tile fixed, queries advance): chunk n = rows T - 128 n + j0,      (lower, upper) = (t + 1, t)
    if (ptx::elect_one()) {
      auto load_pw = [&](int n) {
        const int slot = n % 3;
        if (n >= 3) ptx::mbar_wait(bar(bPwEmpty + slot), ((n / 3) - 1) & 1);
        ptx::mbar_arrive_expect_tx(bar(bPwFull + slot), kTileBytes);
        ptx::tma_load_4d(sPw + slot * kTileBytes, &tmPos, bar(bPwFull + slot), 0, kDq ? a.N - kTile - t0 + kTile * n : a.N - kTile * n + t0, h, 0);
      };
      auto load_ring = [&](int t) {
        const int slot = t & 1;
        if (t >= 2) ptx::mbar_wait(bar(bRingEmpty + slot), ((t >> 1) - 1) & 1);
        ptx::mbar_arrive_expect_tx(bar(bRingFull + slot), kTileBytes);
        ptx::tma_load_4d(sRing + slot * kTileBytes, kDq ? &tmK : &tmQU, bar(bRingFull + slot), 0, t * kTile, h, b);
      };
      auto load_y = [&](int t) {   // dKV only
        if (t >= 1) ptx::mbar_wait(bar(bYEmpty), (t - 1) & 1);
        ptx::mbar_arrive_expect_tx(bar(bYFull), kTileBytes);
        ptx::tma_load_4d(sY, &tmQV, bar(bYFull), 0, t * kTile, h, b);
      };
      auto load_x = [&](int t) {
        if (t >= 1) ptx::mbar_wait(bar(bXEmpty), (t - 1) & 1);
        ptx::mbar_arrive_expect_tx(bar(bXFull), kTileBytes);
        ptx::tma_load_4d(sX, kDq ? &tmV : &tmDO, bar(bXFull), 0, t * kTile, h, b);
      };
      if (kDq) {
        ptx::mbar_arrive_expect_tx(bar(bResFull), 3 * kTileBytes);
        ptx::tma_load_4d(sRes, &tmQU, bar(bResFull), 0, t0, h, b);
        ptx::tma_load_4d(sRes + kTileBytes, &tmQV, bar(bResFull), 0, t0, h, b);
        ptx::tma_load_4d(sRes + 2 * kTileBytes, &tmDO, bar(bResFull), 0, t0, h, b);
      } else {
        ptx::mbar_arrive_expect_tx(bar(bResFull), 2 * kTileBytes);
        ptx::tma_load_4d(sRes, &tmK, bar(bResFull), 0, t0, h, b);
        ptx::tma_load_4d(sRes + kTileBytes, &tmV, bar(bResFull), 0, t0, h, b);
      }
      load_pw(0);
      // the operands of step t, in the order in which their buffers fall free
      for (int t = 0; t < n_tiles; ++t) {
        load_pw(t + 1);
        if (!kDq) load_y(t);
        load_ring(t);
        load_x(t);
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    const uint32_t uRes = ptx::smem_u32(sRes), uRing = ptx::smem_u32(sRing), uX = ptx::smem_u32(sX), uY = ptx::smem_u32(sY), uPw = ptx::smem_u32(sPw),
                   uP = ptx::smem_u32(smem + oP), uDs = ptx::smem_u32(smem + oDs);
    const uint32_t uQV = kDq ? uRes + kTileBytes : uY;
    ptx::mbar_wait(bar(bResFull), 0);
    auto issue_s = [&](int t) {      // AC(t) = (q+u) k^T
      ptx::mbar_wait(bar(bRingFull + (t & 1)), (t >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t ring = uRing + (t & 1) * kTileBytes;
      if (ptx::elect_one()) mma_k64(tmem, kDq ? uRes : ring, kDq ? ring : uRes, kIdescS, false);
      __syncwarp();
    };
    // BD(t) = (q+v) Pw^T over chunks t and t + 1, as two N = 128 halves: the upper half of the window can be formed as soon as the skew of
    // step t-1 is over, the lower half shares its columns with dP(t-1) and waits until that is in registers.
    auto issue_bd_hi = [&](int t) {
      if (t == 0) ptx::mbar_wait(bar(bPwFull), 0);
      if (kDq) ptx::mbar_wait(bar(bPwFull + (t + 1) % 3), ((t + 1) / 3) & 1);
      else ptx::mbar_wait(bar(bYFull), t & 1);
      ptx::tc_fence_after();
      const uint32_t c0 = uPw + (t % 3) * kTileBytes, c1 = uPw + ((t + 1) % 3) * kTileBytes;
      if (ptx::elect_one()) mma_k64(tmem + 256, uQV, kDq ? c1 : c0, kIdescS, false);
      __syncwarp();
    };
    auto issue_bd_lo = [&](int t) {
      if (!kDq) ptx::mbar_wait(bar(bPwFull + (t + 1) % 3), ((t + 1) / 3) & 1);
      ptx::tc_fence_after();
      const uint32_t c0 = uPw + (t % 3) * kTileBytes, c1 = uPw + ((t + 1) % 3) * kTileBytes;
      if (ptx::elect_one()) {
        mma_k64(tmem + 128, uQV, kDq ? c0 : c1, kIdescS, false);
        ptx::tc_commit(bar(bSFull));
        ptx::tc_commit(bar(bPwEmpty + t % 3));     // chunk t leaves the window
        if (!kDq) ptx::tc_commit(bar(bYEmpty));
      }
      __syncwarp();
    };
    issue_s(0);
    issue_bd_hi(0);
    issue_bd_lo(0);
    for (int t = 0; t < n_tiles; ++t) {
      ptx::mbar_wait(bar(bPReady), t & 1);             // P is in registers: the S and BD columns are free
      ptx::mbar_wait(bar(bXFull), t & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        mma_k64(tmem + 128, kDq ? uRes + 2 * kTileBytes : uX, kDq ? uX : uRes + kTileBytes, kIdescS, false);   // dP = dO v^T
        ptx::tc_commit(bar(bDpFull));
        if (kDq) ptx::tc_commit(bar(bXEmpty));
      }
      __syncwarp();
      if (t + 1 < n_tiles) {
        issue_s(t + 1);
        issue_bd_hi(t + 1);
      }
      ptx::mbar_wait(bar(bDpRead), t & 1);             // dP is in registers
      if (t + 1 < n_tiles) issue_bd_lo(t + 1);
      ptx::mbar_wait(bar(bDsFull), t & 1);             // P / dS tiles are in shared memory
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t ring = uRing + (t & 1) * kTileBytes;
        if (kDq) {
          mma_k128_mn(tmem + 384, uDs, ring, kIdescPV, t > 0);   // d(q+u) += dS K
        } else {
          mma_k128_amn(tmem + 384, uP, uX, t > 0);               // dV += P^T dO
          ptx::tc_commit(bar(bPFree));                           // the P tile (= the next step's skew scratch) is free
          ptx::tc_commit(bar(bXEmpty));
          mma_k128_amn(tmem + 448, uDs, ring, t > 0);            // dK += dS^T (q+u)
        }
        ptx::tc_commit(bar(bRingEmpty + (t & 1)));
        ptx::tc_commit(bar(bFin));
        if (t == n_tiles - 1) ptx::tc_commit(bar(bAccFull));
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax warps: thread = (query row, column half) ----------------
    const int wq = warp & 3, g = warp >> 2;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
    // this warp's rows of the P tile; dQ kernel: its half of the lane quarter's 8 KB dBD staging area
    unsigned char* blkP = smem + oP + (kDq ? wq * 8192 + g * 4096 : g * kTileBytes + wq * 4096);
    uint32_t cw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // dQ kernel: last staged word of the odd rows this warp writes (previous key tile)
    // scratch row of this lane: 24 floats at pitch 26, lanes 16-31 shifted by 8 words (conflict-free 8-byte stores and skewed 4-byte reads)
    float* scr = reinterpret_cast<float*>(blkP) + lane * kPitch + ((lane & 16) >> 1);
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0;
    const float sl2 = a.sl2;
    const uint64_t sl2_2 = ptx::pack2(sl2, sl2);
    float my_lse = 0.f, my_delta = 0.f;
    if (kDq) {
      my_lse = a.lse[stat_base + t0 + r];
      my_delta = a.delta[stat_base + t0 + r];
    }
    for (int t = 0; t < n_tiles; ++t) {
      const int i0 = kDq ? t0 : t * kTile, j0 = kDq ? t * kTile : t0;
      if (!kDq) {
        my_lse = a.lse[stat_base + i0 + r];
        my_delta = a.delta[stat_base + i0 + r];
      }
      const int nvalid = a.N - j0 - 64 * g;  // key columns of this half that exist (may be <= 0 or >= 64)
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 0);
      ptx::mbar_wait(bar(bSFull), t & 1);
      ptx::tc_fence_after();
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 1);
      // phase A: P = exp2((AC + shift(BD)) * c - lse), kept packed in registers.
      // s[16 q + cc] += BD[r][127 - r + 64 g + 16 q + cc]: a 48-column window of this warp's rows is parked in the scratch row and read back
      // at the lane's offset, four times; the TMEM load of window q + 1 is in flight while window q goes through shared memory.
      float s[64];
      uint32_t xa0[32], xa1[16], xb0[32], xb1[16];
      auto ld_win = [&](int q, uint32_t (&y0)[32], uint32_t (&y1)[16]) {
        const int wb = 96 - 32 * wq + 64 * g + 16 * q;
        ptx::tmem_ld_32x32(t_lane + 128 + wb, y0);
        ptx::tmem_ld_32x16(t_lane + 128 + wb + 32, y1);
      };
      // Lane l reads its 16 values at window offset 31 - l = 8 (3 - (l >> 3)) + 7 - (l & 7): the coarse part is resolved in registers (two
      // levels of selects on lane bits 4 and 3), so that only 24 of the 48 window columns go through shared memory (12 STS.64 + 16 LDS per
      // window instead of 24 + 16: the phase is bound by the shared-memory pipe).
      auto skew = [&](int q, const uint32_t (&y0)[32], const uint32_t (&y1)[16]) {
        uint32_t a16[32], y[24];
#pragma unroll
        for (int k = 0; k < 32; ++k) a16[k] = b4 ? y0[k] : (k < 16 ? y0[k + 16] : y1[k - 16]);
#pragma unroll
        for (int k = 0; k < 24; ++k) y[k] = b3 ? a16[k] : a16[k + 8];
#pragma unroll
        for (int k = 0; k < 12; ++k) *reinterpret_cast<uint2*>(scr + 2 * k) = make_uint2(y[2 * k], y[2 * k + 1]);
        const volatile float* rd = scr + (7 - (lane & 7));
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) s[16 * q + cc] += rd[cc];
      };
      {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32(t_lane + 64 * g, v0);
        ptx::tmem_ld_32x32(t_lane + 64 * g + 32, v1);
        ld_win(0, xa0, xa1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = __uint_as_float(v0[i]);
          s[32 + i] = __uint_as_float(v1[i]);
        }
      }
      if (!kDq) ptx::mbar_wait(bar(bPFree), (t & 1) ^ 1);   // the previous step's dV MMA has finished with the P tile (= scratch rows 0-19)
      else ptx::bar_sync(1 + wq, 64);                       // the partner warp has written out the rows it read from this warp's staging half
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 2);
      ld_win(1, xb0, xb1);
      skew(0, xa0, xa1);
      ptx::tmem_ld_wait();
      ld_win(2, xa0, xa1);
      skew(1, xb0, xb1);
      ptx::tmem_ld_wait();
      ld_win(3, xb0, xb1);
      skew(2, xa0, xa1);
      ptx::tmem_ld_wait();
      skew(3, xb0, xb1);
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 3);
      const uint64_t nlse2 = ptx::pack2(-my_lse, -my_lse);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float t0_, t1_;
        ptx::unpack2(ptx::fma2(ptx::pack2(s[2 * i], s[2 * i + 1]), sl2_2, nlse2), t0_, t1_);
        float p0 = ex2(t0_), p1 = ex2(t1_);
        if (nvalid < 64) {
          if (2 * i >= nvalid) p0 = 0.f;
          if (2 * i + 1 >= nvalid) p1 = 0.f;
        }
        pk[i] = pack_bf16(p0, p1);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(bPReady));
      if (!kDq) {
        uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[0]);
        uint32_t (&hi)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[16]);
        store_row_chunk(smem + oP, r, 64 * g, lo);
        store_row_chunk(smem + oP, r, 64 * g + 32, hi);
      }
      // phase B: dS = P (dP - delta)
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 4);
      ptx::mbar_wait(bar(bDpFull), t & 1);
      ptx::tc_fence_after();
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 5);
      uint32_t pd[32];
      {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32(t_lane + 128 + 64 * g, v0);
        ptx::tmem_ld_32x32(t_lane + 128 + 64 * g + 32, v1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar(bDpRead));
        const uint64_t nd2 = ptx::pack2(-my_delta, -my_delta);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t (&v)[32] = (i < 16) ? v0 : v1;
          const int k = (i & 15) * 2;
          const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk[i]));
          float d0, d1;
          ptx::unpack2(ptx::mul2(ptx::pack2(p.x, p.y), ptx::add2(ptx::pack2(__uint_as_float(v[k]), __uint_as_float(v[k + 1])), nd2)), d0, d1);
          pd[i] = pack_bf16(d0, d1);
        }
      }
      ptx::mbar_wait(bar(bFin), (t & 1) ^ 1);   // the previous step's MMAs have finished with the dS tile
      {
        uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pd[0]);
        uint32_t (&hi)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pd[16]);
        store_row_chunk(smem + oDs, r, 64 * g, lo);
        store_row_chunk(smem + oDs, r, 64 * g + 32, hi);
      }
      if (kDq) {
        uint4* dst = reinterpret_cast<uint4*>(smem + oP + wq * 8192 + lane * 256) + 8 * g;
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q ^ (lane & 7)] = make_uint4(pd[4 * q], pd[4 * q + 1], pd[4 * q + 2], pd[4 * q + 3]);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(bDsFull));
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 6);
      if (kDq) {
        // dBD[b, h, i, T-1-i+j0 + c] <- dS[i, j0 + c], c < 128: rows 4 m + 2 g + {0, 1} of the lane quarter, lane = destination words l, l + 32
        ptx::bar_sync(1 + wq, 64);                         // both halves of the quarter's rows are staged
        // Branch-free: row rl goes to dbase + rl (ld - 1) - par with par = parity of its first destination column = P0 ^ (rl & 1), the
        // same for every key tile; the number of elements to store, N - j0 (+ the borrowed one), is the same for every row.
        const uint32_t* stq = reinterpret_cast<const uint32_t*>(smem + oP + wq * 8192);
        const int ib = i0 + wq * 32, X0 = a.N - 1 - ib + j0, P0 = X0 & 1;
        const int rows = a.N - ib, nkeys = a.N - j0, pitch1 = (int)ra.dbd_ld - 1;
        __nv_bfloat16* dbase = ra.dbd + (((long long)b * a.H + h) * a.N + ib) * ra.dbd_ld + X0 + 2 * g * pitch1;
        const int lg = lane ^ (g << 3), lgm = ((lane + 63) & 63) ^ (g << 3), lg31 = (lane + 31) ^ (g << 3);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int par = P0 ^ u, sw = (((m & 1) << 2) | u) << 2;
            const uint32_t* srow = stq + (4 * m + u) * 64 + g * 128;
            const uint32_t A = srow[lg ^ sw], Bw = srow[(lg ^ sw) + 32], Ams = srow[lgm ^ sw], Bm = srow[lg31 ^ sw];
            const uint32_t Am = lane ? Ams : cw[m];
            const int sh = par ? 16 : 32;
            const uint32_t o0 = __funnelshift_rc(Am, A, sh), o1 = __funnelshift_rc(Bm, Bw, sh);
            const uint32_t last = __shfl_sync(0xffffffffu, Bw, 31);
            if (par) cw[m] = last;
            uint32_t* d = reinterpret_cast<uint32_t*>(dbase + (4 * m + u) * pitch1 - par);
            const int lim = (4 * m + 2 * g + u < rows) ? nkeys + par : 0;
            if (2 * lane < lim) d[lane] = o0;
            if (2 * lane + 64 < lim) d[lane + 32] = o1;
            if (par && t == n_tiles - 1 && lane == 0 && 128 < lim) d[64] = last >> 16;   // N = 128 n: the band's very last element
          }
        }
      }
      T4S_TRACE_AT(warp + (kDq ? 0 : 10), t, 7);
    }
    // ---- accumulators: each column half writes 32 of the 64 head-dim columns ----
    ptx::mbar_wait(bar(bAccFull), 0);
    ptx::tc_fence_after();
    const int row = t0 + r;
    uint32_t v0[32];
    ptx::tmem_ld_32x32(t_lane + 384 + 32 * g, v0);
    ptx::tmem_ld_wait();
    if (kDq) {
      if (row < a.N) store_row32(ra.dqu + (long long)b * ra.dqu_bs + (long long)row * ra.dqu_ld + h * kHd + 32 * g, v0, a.scale);
    } else {
      if (row < a.N) store_row32(a.dv + (long long)b * a.dv_bs + (long long)row * a.dv_ld + h * kHd + 32 * g, v0, 1.f);
      ptx::tmem_ld_32x32(t_lane + 448 + 32 * g, v0);
      ptx::tmem_ld_wait();
      if (row < a.N) store_row32(a.dk + (long long)b * a.dk_bs + (long long)row * a.dk_ld + h * kHd + 32 * g, v0, a.scale);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, kTmemCols);
  }
}

static int check_rel(const T4sRelAttn* p) {
  T4S_REQUIRE(p, "t4s_relattn: null descriptor");
  T4S_REQUIRE(p->head_dim == kHd, "t4s_relattn: head_dim must be 64 (got %d)", p->head_dim);
  T4S_REQUIRE(p->batch > 0 && p->heads > 0 && p->tokens > 0, "t4s_relattn: batch, heads and tokens must be positive");
  T4S_REQUIRE(p->batch <= 65535 && p->heads <= 65535, "t4s_relattn: batch / heads exceed the grid limits");
  T4S_REQUIRE(p->qu && p->qv && p->k && p->v && p->pos && p->o && p->lse, "t4s_relattn: qu, qv, k, v, pos, o and lse are required");
  T4S_REQUIRE(!(reinterpret_cast<uintptr_t>(p->o) & 15) && !(p->o_ld % 8) && !(p->o_bs % 8), "t4s_relattn: o must be 16-byte aligned with pitches % 8 == 0");
  return T4S_OK;
}

static void fill_rel_args(Args& a, const T4sRelAttn* p) {
  a.N = p->tokens;
  a.n_tiles = (p->tokens + kTile - 1) / kTile;
  a.Nl = a.n_tiles * kTile;
  a.H = p->heads;
  a.scale = p->scale;
  a.sl2 = p->scale * 1.4426950408889634f;
  a.o = reinterpret_cast<__nv_bfloat16*>(p->o);
  a.o_ld = p->o_ld;
  a.o_bs = p->o_bs;
  a.lse = p->lse;
  a.o32 = p->o32;
  a.delta = nullptr;
  a.dq = a.dk = a.dv = nullptr;
  a.dq_ld = a.dq_bs = a.dk_ld = a.dk_bs = a.dv_ld = a.dv_bs = 0;
}

}  // namespace rel
}  // namespace attn
}  // namespace t4s

#ifdef T4S_TRACE
extern "C" int t4s_debug_trace_rel(long long* out, int n) {
  if (n > 4096) n = 4096;
  T4S_CUDA(cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * n));
  return T4S_OK;
}
#endif

extern "C" int t4s_relattn_fwd(const T4sRelAttn* p, void* stream) {
  using namespace t4s::attn;
  using namespace t4s::attn::rel;
  int rc = check_rel(p);
  if (rc) return rc;
  const int B = p->batch, H = p->heads, T = p->tokens;
  fwd2::Maps tm;
  if ((rc = make_map(&tm.q, p->qu, p->qu_ld, p->qu_bs, B, H, T, "qu"))) return rc;
  if ((rc = make_map(&tm.qv, p->qv, p->qv_ld, p->qv_bs, B, H, T, "qv"))) return rc;
  if ((rc = make_map(&tm.k, p->k, p->k_ld, p->k_bs, B, H, T, "k"))) return rc;
  if ((rc = make_map(&tm.v, p->v, p->v_ld, p->v_bs, B, H, T, "v"))) return rc;
  if ((rc = make_map(&tm.pos, p->pos, p->pos_ld, p->pos_ld * (2LL * T - 1), 1, H, 2 * T - 1, "pos", 256))) return rc;
  Args a;
  fill_rel_args(a, p);
  T4S_CUDA(cudaFuncSetAttribute(fwd2::attn_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd2::Layout<true>::kSmem));
  dim3 grid(a.n_tiles, H, B);
  fwd2::attn_fwd2_kernel<true><<<grid, fwd2::Layout<true>::kThreads, fwd2::Layout<true>::kSmem, t4s::as_stream(stream)>>>(tm, a);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

extern "C" int t4s_relattn_bwd(const T4sRelAttnBwd* p, void* stream) {
  using namespace t4s::attn;
  using namespace t4s::attn::rel;
  T4S_REQUIRE(p, "t4s_relattn_bwd: null descriptor");
  const T4sRelAttn* f = &p->fwd;
  int rc = check_rel(f);
  if (rc) return rc;
  T4S_REQUIRE(p->d_o && p->delta && p->dqu && p->dk && p->dv && p->dbd, "t4s_relattn_bwd: d_o, delta, dqu, dk, dv and dbd are required");
  for (const void* ptr : {(const void*)p->dqu, (const void*)p->dk, (const void*)p->dv})
    T4S_REQUIRE(!(reinterpret_cast<uintptr_t>(ptr) & 15), "t4s_relattn_bwd: dqu / dk / dv must be 16-byte aligned");
  T4S_REQUIRE(!(p->dqu_ld % 8) && !(p->dqu_bs % 8) && !(p->dk_ld % 8) && !(p->dk_bs % 8) && !(p->dv_ld % 8) && !(p->dv_bs % 8),
              "t4s_relattn_bwd: gradient pitches must be multiples of 8 elements");
  const int B = f->batch, H = f->heads, T = f->tokens;
  CUtensorMap tqu, tqv, tk, tv, tdo, tpos;
  if ((rc = make_map(&tqu, f->qu, f->qu_ld, f->qu_bs, B, H, T, "qu"))) return rc;
  if ((rc = make_map(&tqv, f->qv, f->qv_ld, f->qv_bs, B, H, T, "qv"))) return rc;
  if ((rc = make_map(&tk, f->k, f->k_ld, f->k_bs, B, H, T, "k"))) return rc;
  if ((rc = make_map(&tv, f->v, f->v_ld, f->v_bs, B, H, T, "v"))) return rc;
  if ((rc = make_map(&tdo, p->d_o, p->do_ld, p->do_bs, B, H, T, "d_o"))) return rc;
  if ((rc = make_map(&tpos, f->pos, f->pos_ld, f->pos_ld * (2LL * T - 1), 1, H, 2 * T - 1, "pos", kTile))) return rc;   // 128-row chunks
  T4S_REQUIRE(!(reinterpret_cast<uintptr_t>(p->dbd) & 15) && !(p->dbd_ld % 8) && p->dbd_ld >= 2LL * T - 1,
              "t4s_relattn_bwd: dbd needs a 16-byte aligned base and a row pitch >= 2T-1 that is a multiple of 8 elements");
  RelArgs ra;
  fill_rel_args(ra.a, f);
  ra.a.delta = p->delta;
  ra.a.dk = reinterpret_cast<__nv_bfloat16*>(p->dk); ra.a.dk_ld = p->dk_ld; ra.a.dk_bs = p->dk_bs;
  ra.a.dv = reinterpret_cast<__nv_bfloat16*>(p->dv); ra.a.dv_ld = p->dv_ld; ra.a.dv_bs = p->dv_bs;
  ra.dqu = reinterpret_cast<__nv_bfloat16*>(p->dqu); ra.dqu_ld = p->dqu_ld; ra.dqu_bs = p->dqu_bs;
  ra.dbd = reinterpret_cast<__nv_bfloat16*>(p->dbd); ra.dbd_ld = p->dbd_ld;
  cudaStream_t st = t4s::as_stream(stream);
  rc = launch_delta(f->o, f->o_ld, f->o_bs, f->o32, p->d_o, p->do_ld, p->do_bs, p->delta, B, H, T, ra.a.Nl, st);
  if (rc) return rc;
  dim3 grid(ra.a.n_tiles, H, B);
  T4S_CUDA(cudaFuncSetAttribute(relattn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmem));
  relattn_bwd_kernel<true><<<grid, bwd::kBThreads, bwd::kSmem, st>>>(tqu, tqv, tk, tv, tdo, tpos, ra);
  T4S_LAUNCH_CHECK();
  T4S_CUDA(cudaFuncSetAttribute(relattn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmem));
  relattn_bwd_kernel<false><<<grid, bwd::kBThreads, bwd::kSmem, st>>>(tqu, tqv, tk, tv, tdo, tpos, ra);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
