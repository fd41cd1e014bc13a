// Forward kernel of the fused multi-head self-attention (attn.cu), head_dim 64, bf16.
//
// CTA = TWO 128-query tiles (groups) of one (clip, head), walking the 128-key tiles; 19 warps:
//   warps 0-15  softmax: group gq = w >> 3, TMEM lane quarter wq = w & 3 (query rows 32 wq ..), column half g = (w >> 2) & 1.
//               A thread reads its 64 scores from TMEM once, keeps them in registers (mask, max, exponentials, row sum with packed
//               f32x2 math) and writes P (bf16 pairs) back into TMEM, where the tensor core reads it as the A operand of P V
//               (tcgen05.mma with A in tensor memory): the P tile never touches shared memory, whose bandwidth the S = Q K^T and
//               P V operand fetches already use up (r2 ncu: tensor-core reads 30 % + P stores 19 % of the data pipe with P in smem).
//               The two halves of a row agree on the running maximum through a 1 KB shared-memory exchange and a 64-thread named
//               barrier per tile, so both accumulate into ONE O tile (64 TMEM columns) and the group fits in 256 columns.
//   warp 16     TMA producer (Q tiles once, K / V rings shared by both groups).
//   warps 17,18 one tcgen05.mma issuer per group, so that neither group waits behind the other's barriers: S_{j+1} = Q K_{j+1}^T is
//               issued as soon as the group's 8 warps hold S_j in registers (it runs under tile j's exponentials, S and P have
//               their own columns), O += P_j V_j when P_j is handed over.  MMA / TMA instructions are issued under elect.sync.
//   The softmax warps therefore never wait for the tensor core in steady state; the two groups share the MUFU pipe, which is
//   the resource that bounds head size 64 (16384 exponentials = 1024 clk per tile against 512 clk of MMA).
//   Rescaling is lazy: a row's reference maximum only moves when the new maximum exceeds it by more than 2^8 (exactness is not
//   affected: the reference cancels in O / l); O is then rescaled in place (tcgen05.ld / .st), each half its 32 columns.
#pragma once
#include "attn_common.cuh"

#ifndef T4S_FWD_POLY_OF_8
#define T4S_FWD_POLY_OF_8 3
#endif

namespace t4s {
namespace attn {
namespace fwd3 {

constexpr int kThreads = 19 * 32;
constexpr float kRescaleThreshold = 8.0f;   // log2 units
constexpr int kPolyOf8 = T4S_FWD_POLY_OF_8;  // of every 8 column pairs, this many take the polynomial exp2 (FMA pipe), the rest MUFU
constexpr int oQ = 0, oK = oQ + 2 * kTileBytes, oV = oK + 2 * kTileBytes, oX = oV + 2 * kTileBytes, oBar = oX + 2 * 3072;
constexpr int kSmem = oBar + 256;
constexpr int kTmemCols = 512;   // group g: S [256 g, +128)  P [256 g + 128, +64) (bf16 pairs)  O [256 g + 192, +64)
enum { bQFull = 0, bKFull = 1, bKEmpty = 3, bVFull = 5, bVEmpty = 7, bSFull = 9, bSFree = 11, bPFull = 13, bOFull = 15 /* [group][parity] */,
       bCount = 19 };

struct Maps {
  CUtensorMap q, k, v;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd3_kernel(const __grid_constant__ Maps tm, const Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * kTile), h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = a.n_tiles;
  const int groups = (q0 + kTile < a.N) ? 2 : 1;   // the second query tile may lie entirely beyond N

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("t4s attn_fwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::mbar_init(&bars[bQFull], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars[bKFull + i], 1);
      ptx::mbar_init(&bars[bKEmpty + i], groups);
      ptx::mbar_init(&bars[bVFull + i], 1);
      ptx::mbar_init(&bars[bVEmpty + i], groups);
      ptx::mbar_init(&bars[bSFull + i], 1);
      ptx::mbar_init(&bars[bSFree + i], 8);
      ptx::mbar_init(&bars[bPFull + i], 8);
      ptx::mbar_init(&bars[bOFull + 2 * i], 1);
      ptx::mbar_init(&bars[bOFull + 2 * i + 1], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 16 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tm.q);
    ptx::prefetch_tmap(&tm.k);
    ptx::prefetch_tmap(&tm.v);
  }
  if (warp == 17) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  T4S_TRACE_AT(16, 15, warp == 0 ? 0 : 7);

  if (warp == 16) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&bars[bQFull], groups * kTileBytes);
      for (int g = 0; g < groups; ++g) ptx::tma_load_4d(smem + oQ + g * kTileBytes, &tm.q, &bars[bQFull], 0, q0 + g * kTile, h, b);
      auto load_k = [&](int j) {
        const int s = j & 1;
        ptx::mbar_wait(&bars[bKEmpty + s], ((j >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars[bKFull + s], kTileBytes);
        ptx::tma_load_4d(smem + oK + s * kTileBytes, &tm.k, &bars[bKFull + s], 0, j * kTile, h, b);
      };
      load_k(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) load_k(j + 1);
        const int s = j & 1;
        ptx::mbar_wait(&bars[bVEmpty + s], ((j >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&bars[bVFull + s], kTileBytes);
        ptx::tma_load_4d(smem + oV + s * kTileBytes, &tm.v, &bars[bVFull + s], 0, j * kTile, h, b);
      }
    }
  } else if (warp >= 17) {
    // ---------------- MMA issuer of group warp - 17 ----------------
    const int g = warp - 17;
    if (g < groups) {
      const uint32_t sQ = ptx::smem_u32(smem + oQ + g * kTileBytes), sK = ptx::smem_u32(smem + oK), sV = ptx::smem_u32(smem + oV);
      const uint32_t tS = tmem + 256 * g, tP = tS + 128, tO = tS + 192;
      auto issue_s = [&](int j) {
        const int s = j & 1;
        ptx::mbar_wait(&bars[bKFull + s], (j >> 1) & 1);
        if (j > 0) ptx::mbar_wait(&bars[bSFree + g], (j - 1) & 1);   // the group's warps hold S_{j-1} in registers
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          mma_k64(tS, sQ, sK + s * kTileBytes, kIdescS, false);
          ptx::tc_commit(&bars[bKEmpty + s]);
          ptx::tc_commit(&bars[bSFull + g]);
        }
        __syncwarp();
      };
      ptx::mbar_wait(&bars[bQFull], 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        const int s = j & 1;
        ptx::mbar_wait(&bars[bVFull + s], (j >> 1) & 1);
        T4S_TRACE_AT(17 + g, j, 0);
        ptx::mbar_wait(&bars[bPFull + g], j & 1);
        T4S_TRACE_AT(17 + g, j, 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // O (+)= P V: A = P from TMEM (128 keys = 64 columns of bf16 pairs), B = V consumed MN-major
          const uint64_t bdesc = ptx::umma_desc_sw128(sV + s * kTileBytes, 8192, 1024);
#pragma unroll
          for (int k = 0; k < 8; ++k) ptx::mma_f16_ts(tO, tP + 8 * k, bdesc + 128 * k, kIdescPV, (j > 0 || k > 0) ? 1u : 0u);
          ptx::tc_commit(&bars[bVEmpty + s]);
          ptx::tc_commit(&bars[bOFull + 2 * g + (j & 1)]);
        }
        __syncwarp();
        T4S_TRACE_AT(17 + g, j, 2);
      }
    }
  } else if ((warp >> 3) < groups) {
    // ---------------- softmax warps: thread = (query row, column half) of one group ----------------
    const int gq = warp >> 3, wq = warp & 3, g = (warp >> 2) & 1;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem + 256 * gq + ((uint32_t)(wq * 32) << 16);
    uint64_t* o_full = &bars[bOFull + 2 * gq];
    float* xm = reinterpret_cast<float*>(smem + oX + gq * 3072);   // [parity][half][row] tile maxima, then [half][row] row sums at 2048
    const int pair_bar = 1 + 4 * gq + wq;
    const float sl2 = a.sl2;
    const uint64_t sl2_2 = ptx::pack2(sl2, sl2);
    float m = -INFINITY, l = 0.f;

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
      // S_j is in registers: the tensor core may overwrite it with tile j+1
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bSFree + gq]);
      T4S_TRACE_AT(warp, j, 2);
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
      float mx = fmaxf(mx0, mx1);
      // the row maximum of the tile: exchange with the warp that holds the other 64 columns of the same rows
      float* xj = xm + (j & 1) * 256;
      xj[g * kTile + r] = mx;
      ptx::bar_sync(pair_bar, 64);
      mx = fmaxf(mx, xj[(g ^ 1) * kTile + r]);
      if (j > 0) {
        const bool need = (mx - m) * sl2 > kRescaleThreshold;    // both halves of a row take the same decision
        if (__any_sync(0xffffffffu, need)) {
          ptx::mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);   // the previous P V product has retired: O is stable
          ptx::tc_fence_after();
          const float alpha = need ? ex2((m - mx) * sl2) : 1.f;
          if (need) m = mx;
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {   // this half rescales O columns [32 g, 32 g + 32)
            uint32_t o[8];
            ptx::tmem_ld_32x8(t_lane + 192 + 32 * g + 8 * c, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            ptx::tmem_st_32x8(t_lane + 192 + 32 * g + 8 * c, o);
          }
          ptx::tmem_st_wait();
        }
      } else {
        m = mx;
      }
      T4S_TRACE_AT(warp, j, 3);
      const float mneg = (m == -INFINITY) ? 0.f : -m * sl2;
      const uint64_t mneg2 = ptx::pack2(mneg, mneg);
      uint64_t rs2 = ptx::pack2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float t0, t1;
        ptx::unpack2(ptx::fma2(ptx::pack2(s[2 * i], s[2 * i + 1]), sl2_2, mneg2), t0, t1);
        float p0, p1;
        if ((i & 7) < kPolyOf8) {
          exp2_poly2(t0, t1, p0, p1);
        } else {
          p0 = ex2(t0);
          p1 = ex2(t1);
        }
        rs2 = ptx::add2(rs2, ptx::pack2(p0, p1));
        pk[i] = pack_bf16(p0, p1);
      }
      float r0, r1;
      ptx::unpack2(rs2, r0, r1);
      l += r0 + r1;
      T4S_TRACE_AT(warp, j, 4);
      // P_j (bf16 pairs): 64 keys of this half -> 32 columns at [128 + 32 g, +32) of the group's TMEM block, once the previous
      // tile's P V product has finished reading them
      if (j > 0) {
        ptx::mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        ptx::tc_fence_after();
      }
      ptx::tmem_st_32x32(t_lane + 128 + 32 * g, pk);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bPFull + gq]);
      T4S_TRACE_AT(warp, j, 5);
    }

    T4S_TRACE_AT(warp, 15, 0);
    // ---- epilogue: O / (l_a + l_b), each half writes 32 of the 64 head-dim columns ----
    ptx::mbar_wait(&o_full[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
    ptx::tc_fence_after();
    float* xl = xm + 512;
    xl[g * kTile + r] = l;
    ptx::bar_sync(pair_bar, 64);
    const float Lsum = l + xl[(g ^ 1) * kTile + r];
    const float inv = 1.f / Lsum;
    uint32_t o[32];
    ptx::tmem_ld_32x32(t_lane + 192 + 32 * g, o);
    ptx::tmem_ld_wait();
    const int row = q0 + gq * kTile + r;
    if (g == 0) a.lse[((long long)b * a.H + h) * a.Nl + row] = fmaf(m, sl2, log2f(Lsum));
    if (row < a.N) {
      float out[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) out[i] = __uint_as_float(o[i]) * inv;
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
  if (warp == 17) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, kTmemCols);
  }
}

}  // namespace fwd3
}  // namespace attn
}  // namespace t4s
