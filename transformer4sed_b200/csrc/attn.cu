// K4: fused multi-head self-attention, forward and backward, on tcgen05 / TMEM / TMA (head_dim 64, bf16 operands).
//
// No [N, N] matrix ever reaches HBM.  Per (clip, head) the kernels walk 128 x 128 score tiles:
//
//   forward   attn_fwd_plain.cuh: CTA = two 128-query tiles, 16 softmax warps (thread = query row x column half), S read once
//             from TMEM into registers, P handed to the tensor core through TMEM, O accumulated in TMEM with lazy rescaling,
//             S_{j+1} issued under tile j's exponentials, a share of the exponentials on the FMA pipe.  Emits lse (log2 units).
//   backward  attn_bwd_fused_kernel (default): persistent CTAs over (clip, head, key tile) work items, thread = key row:
//             S^T = K Q_i^T, dP^T = V dO_i^T, P^T = exp2(S^T c - lse_i), dS^T = P^T (dP^T - delta_i) -> shared memory (bf16) ->
//             dV += P^T dO_i, dK += dS^T Q_i in TMEM, dQ_i tile = dS K_j reduce-added (TMA) into an fp32 buffer.
//             attn_bwd_kernel<0> / <1>: the deterministic two-kernel backward (dK/dV kernel + dQ kernel), selected when no fp32
//             dQ workspace is given.
//
// Out-of-range rows are handled by TMA zero fill (loads) and row predicates (stores); out-of-range key columns are masked.
#include <cstdlib>

#include "attn_common.cuh"
#include "attn_fwd_plain.cuh"

namespace t4s {
namespace attn {

// ======================================================================================================
// backward: delta = rowsum(dO * O)
// ======================================================================================================
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long o_ld, long long o_bs, const float* __restrict__ o32,
                                  const __nv_bfloat16* __restrict__ d_o, long long do_ld, long long do_bs,
                                  float* __restrict__ delta, int B, int H, int N, int Nl) {
  const long long wg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg >= (long long)B * Nl) return;
  const int b = (int)(wg / Nl), n = (int)(wg % Nl);
  if (n >= N) {
    for (int hh = lane; hh < H; hh += 32) delta[((long long)b * H + hh) * Nl + n] = 0.f;
    return;
  }
  const int chunks = H * 8;  // 16-byte chunks per token row
  const __nv_bfloat16* op = o + (long long)b * o_bs + (long long)n * o_ld;
  const __nv_bfloat16* dp = d_o + (long long)b * do_bs + (long long)n * do_ld;
  for (int c0 = 0; c0 < chunks; c0 += 32) {
    const int c = c0 + lane;
    float s = 0.f;
    if (c < chunks) {
      const uint4 y = reinterpret_cast<const uint4*>(dp)[c];
      const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&y);
      if (o32) {
        const float4* xp = reinterpret_cast<const float4*>(o32 + ((long long)b * N + n) * ((long long)H * kHd)) + 2 * c;
        const float4 x0 = xp[0], x1 = xp[1];
        const float2 y0 = __bfloat1622float2(yp[0]), y1 = __bfloat1622float2(yp[1]), y2 = __bfloat1622float2(yp[2]),
                     y3 = __bfloat1622float2(yp[3]);
        s = x0.x * y0.x + x0.y * y0.y + x0.z * y1.x + x0.w * y1.y + x1.x * y2.x + x1.y * y2.y + x1.z * y3.x + x1.w * y3.y;
      } else {
        const uint4 x = reinterpret_cast<const uint4*>(op)[c];
        const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&x);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xf = __bfloat1622float2(xp[i]), yf = __bfloat1622float2(yp[i]);
          s = fmaf(xf.x, yf.x, s);
          s = fmaf(xf.y, yf.y, s);
        }
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0 && c < chunks) delta[((long long)b * H + (c >> 3)) * Nl + n] = s;
  }
}

// ======================================================================================================
// backward kernels (shared layout)
// ======================================================================================================
namespace bwd {
constexpr int kThreads = 320;  // warps 0-7 softmax (two column halves), 8 TMA, 9 MMA
// resident pair (K,V for dKV / Q,dO for dQ), streamed pair x 2 stages, two [128x128] bf16 operand tiles, lse/delta stages
// Modes: 0 = dK/dV kernel, 1 = dQ kernel (the deterministic two-kernel backward), 2 = fused: the dK/dV kernel also forms
// dQ_i += dS_ij K_j per tile and adds it to an fp32 dQ buffer with TMA reduce-add (cp.reduce.async.bulk.tensor .add), so S, P
// and dS -- the exponential / softmax work that bounds these kernels -- are computed once instead of twice.
// P / dS operand tiles: two buffers in modes 0 / 1 (fill i+1 while the tensor core reads i), one in mode 2 (the room goes to
// the per-warp dQ staging boxes).
constexpr int kDqBox = 32 * 128;  // per softmax warp: 32 rows x 32 fp32, SWIZZLE_128B
constexpr int oRes = 0, oStr = oRes + 2 * kTileBytes, oP = oStr + 4 * kTileBytes, oDs = oP + 2 * kPBytes, oStat = oDs + 2 * kPBytes,
              oBar = oStat + 2 * 2 * kTile * 4;
constexpr int oDsF = oP + kPBytes, oDqF = oDsF + kPBytes;   // fused layout: P, dS single, then 8 dQ boxes (ends below oStat)
static_assert(oDqF + 8 * kDqBox <= oStat, "fused layout must fit in front of the statistics");
constexpr int kSmem = oBar + 128;
constexpr int kTmemCols = 512;  // S: [0,128)  dP: [128,256)  acc0: [256,320)  acc1: [320,384)  dQ tile (fused): [384,448)
enum { bResFull = 0, bStrFull = 1, bStrEmpty = 3, bSFull = 5, bSFree = 6, bPFull = 7, bPFree = 8 /* +1 */, bAccFull = 10, bDqFull = 11,
       bDqFree = 12, bCount = 13 };
constexpr uint32_t kIdescAmn = ptx::umma_idesc(1, 128, 64, 1, 1);  // A and B MN-major
// D[128 x 64] (+)= A^T . B: A = a [128 (K) x 128 (M)] tile written K-major, consumed MN-major (two 64-wide M blocks 16 KB apart);
// B = [128 rows (K) x 64] tile consumed MN-major
__device__ __forceinline__ void mma_k128_amn(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool accumulate) {
  const uint64_t adesc = ptx::umma_desc_sw128(a_addr, kTileBytes, 1024), bdesc = ptx::umma_desc_sw128(b_addr, 8192, 1024);
#pragma unroll
  for (int k = 0; k < 8; ++k) ptx::mma_f16(d_tmem, adesc + 128 * k, bdesc + 128 * k, kIdescAmn, (accumulate || k > 0) ? 1u : 0u);
}
// shared -> global tile reduce-add (fp32), bulk async-group completion; the box is clipped at the tensor bounds
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
}  // namespace bwd

// kDq = false: dK/dV kernel (resident K_j, V_j; streams Q_i, dO_i, lse_i, delta_i; thread = key row)
// kDq = true : dQ kernel    (resident Q_i, dO_i; streams K_j, V_j;                  thread = query row)
template <int kMode>
__global__ void __launch_bounds__(bwd::kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQ,
                const Args a) {
  using namespace bwd;
  constexpr bool kDq = kMode == 1, kFused = kMode == 2;
  constexpr int kPBuf = kFused ? 0 : kPBytes;          // stride between the two P / dS buffers (0: single buffer)
  constexpr int oDsM = kFused ? oDsF : oDs;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = a.n_tiles;
  const long long stat_base = ((long long)b * a.H + h) * a.Nl;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("t4s attn_bwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::mbar_init(&bars[bResFull], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars[bStrFull + i], 1);
      ptx::mbar_init(&bars[bStrEmpty + i], 1);
    }
    ptx::mbar_init(&bars[bSFull], 1);
    ptx::mbar_init(&bars[bSFree], 8);
    ptx::mbar_init(&bars[bPFull], 8);
    ptx::mbar_init(&bars[bPFree], 1);
    ptx::mbar_init(&bars[bPFree + 1], 1);
    ptx::mbar_init(&bars[bAccFull], 1);
    ptx::mbar_init(&bars[bDqFull], 1);
    ptx::mbar_init(&bars[bDqFree], 8);
    ptx::fence_barrier_init();
  }
  if (warp == 8 && ptx::elect_one()) {
    if (kFused) ptx::prefetch_tmap(&tmDQ);
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::prefetch_tmap(&tmDO);
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // shared-memory roles.  X = operand that pairs with K-major "S" product, Y = operand that pairs with the "dP" product.
  //   dKV: resident (K_j, V_j), streamed (Q_i, dO_i):  S^T = K Q^T,  dP^T = V dO^T,  dV += P^T dO,  dK += dS^T Q
  //   dQ : resident (Q_i, dO_i), streamed (K_j, V_j):  S   = Q K^T,  dP   = dO V^T,  dQ += dS K
  unsigned char* sRes0 = smem + oRes;               // K_j  | Q_i
  unsigned char* sRes1 = smem + oRes + kTileBytes;  // V_j  | dO_i
  float* sStat = reinterpret_cast<float*>(smem + oStat);  // [stage][lse(128) | delta(128)]  (dKV only)

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(&bars[bResFull], 2 * kTileBytes);
      if (kDq) {
        ptx::tma_load_4d(sRes0, &tmQ, &bars[bResFull], 0, t0, h, b);
        ptx::tma_load_4d(sRes1, &tmDO, &bars[bResFull], 0, t0, h, b);
      } else {
        ptx::tma_load_4d(sRes0, &tmK, &bars[bResFull], 0, t0, h, b);
        ptx::tma_load_4d(sRes1, &tmV, &bars[bResFull], 0, t0, h, b);
      }
      for (int i = 0; i < n_tiles; ++i) {
        const int s = i & 1;
        unsigned char* st0 = smem + oStr + s * 2 * kTileBytes;
        ptx::mbar_wait(&bars[bStrEmpty + s], ((i >> 1) & 1) ^ 1);
        if (kDq) {
          ptx::mbar_arrive_expect_tx(&bars[bStrFull + s], 2 * kTileBytes);
          ptx::tma_load_4d(st0, &tmK, &bars[bStrFull + s], 0, i * kTile, h, b);
          ptx::tma_load_4d(st0 + kTileBytes, &tmV, &bars[bStrFull + s], 0, i * kTile, h, b);
        } else {
          ptx::mbar_arrive_expect_tx(&bars[bStrFull + s], 2 * kTileBytes + 2 * kTile * 4);
          ptx::tma_load_4d(st0, &tmQ, &bars[bStrFull + s], 0, i * kTile, h, b);
          ptx::tma_load_4d(st0 + kTileBytes, &tmDO, &bars[bStrFull + s], 0, i * kTile, h, b);
          bulk_load(sStat + s * 2 * kTile, a.lse + stat_base + i * kTile, kTile * 4, &bars[bStrFull + s]);
          bulk_load(sStat + s * 2 * kTile + kTile, a.delta + stat_base + i * kTile, kTile * 4, &bars[bStrFull + s]);
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    const uint32_t r0 = ptx::smem_u32(sRes0), r1 = ptx::smem_u32(sRes1), str = ptx::smem_u32(smem + oStr),
                   sP = ptx::smem_u32(smem + oP), sDs = ptx::smem_u32(smem + oDsM);
    ptx::mbar_wait(&bars[bResFull], 0);
    ptx::mbar_wait(&bars[bStrFull + 0], 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      mma_k64(tmem, r0, str, kIdescS, false);                      // S / S^T
      mma_k64(tmem + 128, r1, str + kTileBytes, kIdescS, false);   // dP / dP^T
      ptx::tc_commit(&bars[bSFull]);
    }
    __syncwarp();
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i & 1;
      if (i + 1 < n_tiles) {
        const uint32_t nx = str + (s ^ 1) * 2 * kTileBytes;
        ptx::mbar_wait(&bars[bStrFull + (s ^ 1)], ((i + 1) >> 1) & 1);
        ptx::mbar_wait(&bars[bSFree], i & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          mma_k64(tmem, r0, nx, kIdescS, false);
          mma_k64(tmem + 128, r1, nx + kTileBytes, kIdescS, false);
          ptx::tc_commit(&bars[bSFull]);
        }
        __syncwarp();
      }
      ptx::mbar_wait(&bars[bPFull], i & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t cur = str + s * 2 * kTileBytes;
        const uint32_t pb = sP + s * kPBuf, db = sDs + s * kPBuf;
        if (kDq) {
          mma_k128_mn(tmem + 256, db, cur, kIdescPV, i > 0);                   // dQ += dS K_j
        } else {
          mma_k128_mn(tmem + 256, pb, cur + kTileBytes, kIdescPV, i > 0);      // dV += P^T dO_i
          mma_k128_mn(tmem + 320, db, cur, kIdescPV, i > 0);                   // dK += dS^T Q_i
        }
        ptx::tc_commit(&bars[bStrEmpty + s]);
        if (kFused) {
          // dQ_i tile = dS_ij K_j (A = the dS^T tile read MN-major, B = the resident K_j): fresh accumulator every tile
          ptx::mbar_wait(&bars[bDqFree], (i & 1) ^ 1);   // the softmax warps have drained tile i-1's product
          ptx::tc_fence_after();
          mma_k128_amn(tmem + 384, db, r0, false);
          ptx::tc_commit(&bars[bDqFull]);
        }
        ptx::tc_commit(&bars[bPFree + (kFused ? 0 : s)]);
        if (i == n_tiles - 1) ptx::tc_commit(&bars[bAccFull]);
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax warps: thread = row of this CTA's tile, column half g ----------------
    const int wq = warp & 3, g = warp >> 2;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
    const float sl2 = a.sl2;
    float my_lse = 0.f, my_delta = 0.f;
    if (kDq) {
      my_lse = a.lse[stat_base + t0 + r];
      my_delta = a.delta[stat_base + t0 + r];
    }
    // fused mode: move query tile `it`'s dQ product (this warp: rows 32 wq.., columns 32 g..) from TMEM to the warp's staging box
    // and add it to the fp32 dQ buffer with one TMA reduce
    unsigned char* dq_box = smem + oDqF + warp * kDqBox;
    auto drain_dq = [&](int it) {
      ptx::mbar_wait(&bars[bDqFull], it & 1);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_lane + 384 + 32 * g, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (ptx::elect_one()) {
        ptx::mbar_arrive(&bars[bDqFree]);
        ptx::bulk_wait_read_all();           // the previous reduce has finished reading the box
      }
      __syncwarp();
      uint4* dst = reinterpret_cast<uint4*>(dq_box + lane * 128);
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q ^ (lane & 7)] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      ptx::fence_proxy_async();
      __syncwarp();
      if (ptx::elect_one()) {
        tma_reduce_add_4d(&tmDQ, dq_box, 32 * g, it * kTile + 32 * wq, h, b);
        ptx::bulk_commit();
      }
    };
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i & 1;
      const float* st = sStat + s * 2 * kTile;
      if (!kDq) ptx::mbar_wait(&bars[bStrFull + s], (i >> 1) & 1);  // lse / delta of this query tile have landed
      ptx::mbar_wait(&bars[bSFull], i & 1);
      ptx::tc_fence_after();
      if (kFused && i > 0) drain_dq(i - 1);
      // the MMAs that last read this P / dS buffer (tile i-2, or i-1 with a single buffer) have finished
      ptx::mbar_wait(&bars[bPFree + (kFused ? 0 : s)], kFused ? ((i & 1) ^ 1) : (((i >> 1) & 1) ^ 1));
      const int nvalid = a.N - i * kTile;          // dQ: key columns that exist
      const bool full = nvalid >= kTile;
      // all four TMEM loads of this thread's 64 columns are issued before the first wait (the exposed tcgen05.ld round trips
      // were the longest stall of the softmax warps), and S / dP are handed back to the MMA warp right after they land
      uint32_t vs0[32], vp0[32], vs1[32], vp1[32];
      ptx::tmem_ld_32x32(t_lane + 64 * g, vs0);
      ptx::tmem_ld_32x32(t_lane + 128 + 64 * g, vp0);
      ptx::tmem_ld_32x32(t_lane + 64 * g + 32, vs1);
      ptx::tmem_ld_32x32(t_lane + 128 + 64 * g + 32, vp1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bSFree]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = 64 * g + 32 * c;
        const uint32_t (&vs)[32] = c ? vs1 : vs0;
        const uint32_t (&vp)[32] = c ? vp1 : vp0;
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float ls[4], dl[4];
          if (kDq) {
            ls[0] = ls[1] = ls[2] = ls[3] = my_lse;
            dl[0] = dl[1] = dl[2] = dl[3] = my_delta;
          } else {
            const float4 L = *reinterpret_cast<const float4*>(st + col0 + 4 * q);
            const float4 Dv = *reinterpret_cast<const float4*>(st + kTile + col0 + 4 * q);
            ls[0] = L.x; ls[1] = L.y; ls[2] = L.z; ls[3] = L.w;
            dl[0] = Dv.x; dl[1] = Dv.y; dl[2] = Dv.z; dl[3] = Dv.w;
          }
          float p[4], d[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[e] = ex2(fmaf(__uint_as_float(vs[4 * q + e]), sl2, -ls[e]));
            if (kDq && !full && col0 + 4 * q + e >= nvalid) p[e] = 0.f;
            d[e] = p[e] * (__uint_as_float(vp[4 * q + e]) - dl[e]);
          }
          pp[2 * q] = pack_bf16(p[0], p[1]);
          pp[2 * q + 1] = pack_bf16(p[2], p[3]);
          pd[2 * q] = pack_bf16(d[0], d[1]);
          pd[2 * q + 1] = pack_bf16(d[2], d[3]);
        }
        if (!kDq) store_row_chunk(smem + oP + s * kPBuf, r, col0, pp);
        store_row_chunk(smem + oDsM + s * kPBuf, r, col0, pd);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bPFull]);
    }
    if (kFused) drain_dq(n_tiles - 1);
    // ---- write the accumulators: each column half takes 32 of the 64 head-dim columns ----
    ptx::mbar_wait(&bars[bAccFull], 0);
    ptx::tc_fence_after();
    const int row = t0 + r;
    uint32_t v[32];
    ptx::tmem_ld_32x32(t_lane + 256 + 32 * g, v);
    ptx::tmem_ld_wait();
    if (kDq) {
      if (row < a.N) store_row32(a.dq + (long long)b * a.dq_bs + (long long)row * a.dq_ld + h * kHd + 32 * g, v, a.scale);
    } else {
      if (row < a.N) store_row32(a.dv + (long long)b * a.dv_bs + (long long)row * a.dv_ld + h * kHd + 32 * g, v, 1.f);
      ptx::tmem_ld_32x32(t_lane + 320 + 32 * g, v);
      ptx::tmem_ld_wait();
      if (row < a.N) store_row32(a.dk + (long long)b * a.dk_bs + (long long)row * a.dk_ld + h * kHd + 32 * g, v, a.scale);
    }
    if (kFused && ptx::elect_one()) ptx::bulk_wait_all();   // the staging boxes must outlive the last reduce
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, kTmemCols);
  }
}

// ======================================================================================================
// backward, one kernel (the default): persistent CTAs walk (clip, head, key tile) work items
// ======================================================================================================
// Per work item the CTA keeps K_j / V_j resident and streams the query tiles: S^T = K Q_i^T and dP^T = V dO_i^T land in TMEM,
// the 8 softmax warps (thread = key row, warp g = w >> 2 takes query columns 64 g ..) form P^T = exp2(S^T c - lse_i) and
// dS^T = P^T (dP^T - delta_i) with packed f32x2 math and write them (bf16) into SWIZZLE_128B operand tiles, the tensor core
// accumulates dV += P^T dO_i and dK += dS^T Q_i in TMEM and forms the dQ_i tile = dS K_j, which the softmax warps add into the
// fp32 dQ buffer with TMA reduce-add.  Everything is software-pipelined around the softmax warps, which never wait for the
// tensor core in steady state:
//   * S / dP of tile T+1 are issued as soon as tile T's S / dP are in registers (they run under tile T's exponentials);
//   * dS is double buffered and P is released by its own commit right after dV, so tile T+1's softmax overlaps tile T's MMAs;
//   * the dQ tile of T-1 is drained AFTER tile T's P / dS have been handed over (its MMA finished long before);
//   * the CTA is persistent: barriers, TMEM and tensor maps are set up once per SM and the accumulator write-back of one work
//     item overlaps the operand loads and first MMAs of the next (r1: 24 % of the warp samples sat in per-CTA prologue /
//     epilogue code with one CTA per SM).
#ifdef T4S_TRACE
#define T4S_TRACE_B(w, j, e)                                                                    \
  do {                                                                                          \
    if (blockIdx.x == 70 && (threadIdx.x & 31) == 0 && (j) >= 20 && (j) < 36)                   \
      g_trace[((w) * 16 + ((j) - 20)) * 8 + (e)] = clock64();                                   \
  } while (0)
#else
#define T4S_TRACE_B(w, j, e) do {} while (0)
#endif
namespace fbw {
constexpr int kThreads = 320;  // warps 0-7 softmax, 8 TMA, 9 MMA
constexpr int kDqBox = 32 * 128;
constexpr int kStr = 3;        // stages of the streamed (Q_i, dO_i) ring
constexpr int oRes = 0, oStr = oRes + 2 * kTileBytes, oP = oStr + kStr * 2 * kTileBytes, oDs = oP + kPBytes, oDq = oDs + kPBytes,
              oStat = oDq + 8 * kDqBox, oBar = oStat + 2 * 2 * kTile * 4;
constexpr int kSmem = oBar + 192;
static_assert(kSmem <= 232448, "fused attention backward: shared memory");
constexpr int kTmemCols = 512;  // S^T [0,128)  dP^T [128,256)  dV [256,320)  dK [320,384)  dQ tile [384,448)  P^T (bf16 pairs) [448,512)
enum { bResFull = 0, bStrFull = 1, bStrEmpty = 4, bStatFull = 7, bStatEmpty = 9, bSFull = 11, bSFree = 12, bPFull = 13, bPvFree = 14, bDsFree = 15,
       bDqFull = 16, bDqFree = 17, bAccFull = 18, bAccFree = 19, bCount = 20 };
}  // namespace fbw

__global__ void __launch_bounds__(fbw::kThreads, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                      const __grid_constant__ CUtensorMap tmDV, const Args a, const int n_items) {
  using namespace fbw;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = a.n_tiles;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("t4s attn_bwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::mbar_init(&bars[bResFull], 1);
    for (int i = 0; i < kStr; ++i) {
      ptx::mbar_init(&bars[bStrFull + i], 1);
      ptx::mbar_init(&bars[bStrEmpty + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars[bStatFull + i], 1);
      ptx::mbar_init(&bars[bStatEmpty + i], 8);
    }
    ptx::mbar_init(&bars[bSFull], 1);
    ptx::mbar_init(&bars[bSFree], 8);
    ptx::mbar_init(&bars[bPFull], 8);
    ptx::mbar_init(&bars[bPvFree], 1);
    ptx::mbar_init(&bars[bDsFree], 1);
    ptx::mbar_init(&bars[bDqFull], 1);
    ptx::mbar_init(&bars[bDqFree], 8);
    ptx::mbar_init(&bars[bAccFull], 1);
    ptx::mbar_init(&bars[bAccFree], 8);
    ptx::fence_barrier_init();
  }
  if (warp == 8 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmDQ);
    ptx::prefetch_tmap(&tmDK);
    ptx::prefetch_tmap(&tmDV);
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::prefetch_tmap(&tmDO);
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  unsigned char* sRes0 = smem + oRes;               // K_j
  unsigned char* sRes1 = smem + oRes + kTileBytes;  // V_j
  float* sStat = reinterpret_cast<float*>(smem + oStat);  // [stage][lse(128) | delta(128)]

  // work item -> (key tile, head, clip): key tiles of one (clip, head) run on neighbouring SMs and share Q / dO in L2
  auto decode = [&](int item, int& t0, int& h, int& b) {
    const int kt = item % n_tiles, bh = item / n_tiles;
    t0 = kt * kTile;
    h = bh % a.H;
    b = bh / a.H;
  };

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      int T = 0, W = 0, st = 0, sph = 0;   // st = T % kStr, sph = (T / kStr) & 1
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++W) {
        int t0, h, b;
        decode(item, t0, h, b);
        if (W > 0) ptx::mbar_wait(&bars[bAccFull], (W - 1) & 1);   // every MMA of the previous item has read K / V
        ptx::mbar_arrive_expect_tx(&bars[bResFull], 2 * kTileBytes);
        ptx::tma_load_4d(sRes0, &tmK, &bars[bResFull], 0, t0, h, b);
        ptx::tma_load_4d(sRes1, &tmV, &bars[bResFull], 0, t0, h, b);
        const long long stat_base = ((long long)b * a.H + h) * a.Nl;
        for (int i = 0; i < n_tiles; ++i, ++T) {
          unsigned char* st0 = smem + oStr + st * 2 * kTileBytes;
          ptx::mbar_wait(&bars[bStrEmpty + st], sph ^ 1);
          ptx::mbar_arrive_expect_tx(&bars[bStrFull + st], 2 * kTileBytes);
          ptx::tma_load_4d(st0, &tmQ, &bars[bStrFull + st], 0, i * kTile, h, b);
          ptx::tma_load_4d(st0 + kTileBytes, &tmDO, &bars[bStrFull + st], 0, i * kTile, h, b);
          const int s2 = T & 1;
          ptx::mbar_wait(&bars[bStatEmpty + s2], ((T >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&bars[bStatFull + s2], 2 * kTile * 4);
          bulk_load(sStat + s2 * 2 * kTile, a.lse + stat_base + i * kTile, kTile * 4, &bars[bStatFull + s2]);
          bulk_load(sStat + s2 * 2 * kTile + kTile, a.delta + stat_base + i * kTile, kTile * 4, &bars[bStatFull + s2]);
          if (++st == kStr) { st = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    const uint32_t r0 = ptx::smem_u32(sRes0), r1 = ptx::smem_u32(sRes1), str = ptx::smem_u32(smem + oStr), sP = ptx::smem_u32(smem + oP),
                   sDs = ptx::smem_u32(smem + oDs);
    int T = 0, W = 0;
    int st = 0, sph = 0;      // stage / phase of global tile T
    int stn = 0, sphn = 0;    // stage / phase of the next tile whose S / dP have not been issued yet
    int Tn = 0;               // that tile's global index
    auto issue_s = [&]() {   // S^T / dP^T of global tile Tn
      const uint32_t nx = str + stn * 2 * kTileBytes;
      ptx::mbar_wait(&bars[bStrFull + stn], sphn);
      if (Tn > 0) ptx::mbar_wait(&bars[bSFree], (Tn - 1) & 1);   // tile Tn-1's S / dP are in registers
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        mma_k64(tmem, r0, nx, kIdescS, false);
        mma_k64(tmem + 128, r1, nx + kTileBytes, kIdescS, false);
        ptx::tc_commit(&bars[bSFull]);
      }
      __syncwarp();
      ++Tn;
      if (++stn == kStr) { stn = 0; sphn ^= 1; }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++W) {
      ptx::mbar_wait(&bars[bResFull], W & 1);
      issue_s();
      for (int i = 0; i < n_tiles; ++i, ++T) {
        T4S_TRACE_B(9, T, 0);
        if (i + 1 < n_tiles) issue_s();
        T4S_TRACE_B(9, T, 1);
        ptx::mbar_wait(&bars[bPFull], T & 1);
        T4S_TRACE_B(9, T, 2);
        if (i == 0 && W > 0) ptx::mbar_wait(&bars[bAccFree], (W - 1) & 1);   // the previous item's dV / dK have been read out
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t cur = str + st * 2 * kTileBytes;
          mma_k128_mn(tmem + 320, sDs, cur, kIdescPV, i > 0);                            // dK += dS^T Q_i
          ptx::mbar_wait(&bars[bDqFree], (T & 1) ^ 1);                                   // tile T-1's dQ product has been read out
          ptx::tc_fence_after();
          bwd::mma_k128_amn(tmem + 384, sDs, r0, false);                                 // dQ_i tile = dS K_j
          ptx::tc_commit(&bars[bDsFree]);
          ptx::tc_commit(&bars[bDqFull]);
          {                                                                              // dV += P^T dO_i, P^T from TMEM
            const uint64_t bdesc = ptx::umma_desc_sw128(cur + kTileBytes, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 8; ++k) ptx::mma_f16_ts(tmem + 256, tmem + 448 + 8 * k, bdesc + 128 * k, kIdescPV, (i > 0 || k > 0) ? 1u : 0u);
          }
          ptx::tc_commit(&bars[bPvFree]);
          ptx::tc_commit(&bars[bStrEmpty + st]);
          if (i == n_tiles - 1) ptx::tc_commit(&bars[bAccFull]);
        }
        __syncwarp();
        T4S_TRACE_B(9, T, 3);
        if (++st == kStr) { st = 0; sph ^= 1; }
      }
    }
    (void)sph;
  } else {
    // ---------------- softmax warps: thread = key row of the resident tile, query-column half g ----------------
    const int wq = warp & 3, g = warp >> 2;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
    const uint64_t nsl2 = ptx::pack2(-a.sl2, -a.sl2);
    unsigned char* dq_box = smem + oDq + warp * kDqBox;
    int T = 0, W = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++W) {
      int t0, h, b;
      decode(item, t0, h, b);
      for (int i = 0; i <= n_tiles; ++i) {
        const bool tail = i == n_tiles;       // extra round at the end of the item: only the drain of the last tile
        uint32_t pp[2][16];
        if (!tail) {
          const int s2 = T & 1;
          const float* stt = sStat + s2 * 2 * kTile + 64 * g;
          T4S_TRACE_B(warp, T, 0);
          ptx::mbar_wait(&bars[bStatFull + s2], (T >> 1) & 1);   // lse / delta of this query tile have landed
          ptx::mbar_wait(&bars[bSFull], T & 1);
          ptx::tc_fence_after();
          T4S_TRACE_B(warp, T, 1);
          uint32_t vs0[32], vp0[32], vs1[32], vp1[32];
          ptx::tmem_ld_32x32(t_lane + 64 * g, vs0);
          ptx::tmem_ld_32x32(t_lane + 128 + 64 * g, vp0);
          ptx::tmem_ld_32x32(t_lane + 64 * g + 32, vs1);
          ptx::tmem_ld_32x32(t_lane + 128 + 64 * g + 32, vp1);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars[bSFree]);
          T4S_TRACE_B(warp, T, 2);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t (&vs)[32] = c ? vs1 : vs0;
            const uint32_t (&vp)[32] = c ? vp1 : vp0;
            uint32_t pd[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 Lq = *reinterpret_cast<const float4*>(stt + 32 * c + 4 * q);
              const float4 Dq = *reinterpret_cast<const float4*>(stt + kTile + 32 * c + 4 * q);
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const uint64_t s2v = ptx::pack2(__uint_as_float(vs[4 * q + 2 * e]), __uint_as_float(vs[4 * q + 2 * e + 1]));
                const uint64_t dp2 = ptx::pack2(__uint_as_float(vp[4 * q + 2 * e]), __uint_as_float(vp[4 * q + 2 * e + 1]));
                const uint64_t l2 = e ? ptx::pack2(Lq.z, Lq.w) : ptx::pack2(Lq.x, Lq.y);
                const uint64_t d2 = e ? ptx::pack2(Dq.z, Dq.w) : ptx::pack2(Dq.x, Dq.y);
                float u0, u1;
                ptx::unpack2(ptx::fma2(s2v, nsl2, l2), u0, u1);          // lse - s c
                const float p0 = ex2(-u0), p1 = ex2(-u1);
                float d0, d1;
                ptx::unpack2(ptx::mul2(ptx::pack2(p0, p1), ptx::sub2(dp2, d2)), d0, d1);
                pp[c][2 * q + e] = pack_bf16(p0, p1);
                pd[2 * q + e] = pack_bf16(d0, d1);
              }
            }
            if (c == 0) ptx::mbar_wait(&bars[bDsFree], (T & 1) ^ 1);   // tile T-1's dK / dQ MMAs have finished with the dS tile
            store_row_chunk(smem + oDs, r, 64 * g + 32 * c, pd);
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars[bStatEmpty + s2]);
          T4S_TRACE_B(warp, T, 3);
        }
        // dQ product of the previous tile (this warp: query rows 32 wq .., columns 32 g ..) -> staging box
        const bool drain = i > 0;   // tile i-1 of this item (the last one in the extra round)
        if (drain) {
          ptx::mbar_wait(&bars[bDqFull], (T - 1) & 1);
          ptx::tc_fence_after();
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_lane + 384 + 32 * g, v);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (ptx::elect_one()) {
            ptx::mbar_arrive(&bars[bDqFree]);
            ptx::bulk_wait_read_all();           // the previous reduce has finished reading the box
          }
          __syncwarp();
          uint4* dst = reinterpret_cast<uint4*>(dq_box + lane * 128);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q ^ (lane & 7)] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (!tail) T4S_TRACE_B(warp, T, 4);
        if (!tail) {
          // P^T (bf16 pairs: this row's 64 queries -> 32 columns) goes to TMEM and is the A operand of dV += P^T dO from there: no
          // shared-memory store and no operand fetch for it (the kernel is bound by the shared-memory pipe)
          ptx::mbar_wait(&bars[bPvFree], (T & 1) ^ 1);   // tile T-1's dV MMA has finished reading P
          ptx::tc_fence_after();
          ptx::tmem_st_32x32(t_lane + 448 + 32 * g, reinterpret_cast<const uint32_t (&)[32]>(pp));
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
        }
        ptx::fence_proxy_async();                        // one fence for P, dS and the dQ box
        __syncwarp();
        if (ptx::elect_one()) {
          if (!tail) ptx::mbar_arrive(&bars[bPFull]);
          if (drain) {
            bwd::tma_reduce_add_4d(&tmDQ, dq_box, 32 * g, (i - 1) * kTile + 32 * wq, h, b);
            ptx::bulk_commit();
          }
        }
        if (!tail) T4S_TRACE_B(warp, T, 5);
        if (!tail) ++T;
      }
      // ---- dV / dK of this work item: each column half writes 32 of the 64 head-dim columns ----
      T4S_TRACE_B(warp, T - 1, 6);
      ptx::mbar_wait(&bars[bAccFull], W & 1);
      T4S_TRACE_B(warp, T - 1, 7);
      ptx::tc_fence_after();
      uint32_t v[32], w[32];
      ptx::tmem_ld_32x32(t_lane + 256 + 32 * g, v);
      ptx::tmem_ld_32x32(t_lane + 320 + 32 * g, w);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bAccFree]);
      const int row = t0 + r;
      // the column sums first: they only need the registers, and the tail's dQ reduce finishes reading the staging box meanwhile
      if (a.colsum) {   // qkv bias gradient: column sums of this warp's 32 key rows
        const int D = a.H * kHd;
        const float sv = warp_colsum32(v, row < a.N, lane), sk = warp_colsum32(w, row < a.N, lane);
        atomicAdd(a.colsum + 2 * D + h * kHd + 32 * g + lane, sv);
        atomicAdd(a.colsum + D + h * kHd + 32 * g + lane, sk * a.scale);
      }
      T4S_TRACE_B(warp, T, 6);
      // This warp's 32 x 32 chunks of dV and dK leave through its (now idle) dQ staging box as two SWIZZLE_64B tiles and one TMA store
      // each: asynchronous and in whole lines, where 16-byte row stores from 32 different rows stalled the warps at every work-item
      // boundary (all CTAs reach it together).  Rows beyond the sequence are clipped by the tensor map.
      if (ptx::elect_one()) ptx::bulk_wait_read_all();   // the item's last dQ reduce has read the box
      __syncwarp();
      {
        uint4* dv4 = reinterpret_cast<uint4*>(dq_box + lane * 64);
        uint4* dk4 = reinterpret_cast<uint4*>(dq_box + 2048 + lane * 64);
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dv4[q ^ sw] = make_uint4(pack_bf16(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1])),
                                   pack_bf16(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                                   pack_bf16(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
                                   pack_bf16(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
          dk4[q ^ sw] = make_uint4(pack_bf16(__uint_as_float(w[8 * q]) * a.scale, __uint_as_float(w[8 * q + 1]) * a.scale),
                                   pack_bf16(__uint_as_float(w[8 * q + 2]) * a.scale, __uint_as_float(w[8 * q + 3]) * a.scale),
                                   pack_bf16(__uint_as_float(w[8 * q + 4]) * a.scale, __uint_as_float(w[8 * q + 5]) * a.scale),
                                   pack_bf16(__uint_as_float(w[8 * q + 6]) * a.scale, __uint_as_float(w[8 * q + 7]) * a.scale));
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (ptx::elect_one()) {
        ptx::tma_store_4d(&tmDV, dq_box, 32 * g, t0 + 32 * wq, h, b);
        ptx::tma_store_4d(&tmDK, dq_box + 2048, 32 * g, t0 + 32 * wq, h, b);
        ptx::bulk_commit();
      }
      T4S_TRACE_B(warp, T, 7);
    }
    if (ptx::elect_one()) ptx::bulk_wait_all();   // the staging boxes must outlive the last reduce
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, kTmemCols);
  }
}

// dq[b, n, h*64 + d] (bf16, strided) = scale * dq32[b, n, h*64 + d]   (fused backward: fp32 reduce buffer -> gradient slot)
__global__ void dq_finish_kernel(const float* __restrict__ dq32, __nv_bfloat16* __restrict__ dq, long long dq_ld, long long dq_bs, int N, int D,
                                 float scale, long long total8) {
  const int c8n = D >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)((unsigned)i / (unsigned)c8n), c = (int)i - row * c8n;   // total8 < 2^31 (checked by the host)
    const int b = row / N, n = row - b * N;
    const float* src = dq32 + i * 8;
    const float4 x = *reinterpret_cast<const float4*>(src), y = *reinterpret_cast<const float4*>(src + 4);
    uint4 u;
    u.x = pack_bf16(x.x * scale, x.y * scale);
    u.y = pack_bf16(x.z * scale, x.w * scale);
    u.z = pack_bf16(y.x * scale, y.y * scale);
    u.w = pack_bf16(y.z * scale, y.w * scale);
    *reinterpret_cast<uint4*>(dq + (long long)b * dq_bs + (long long)n * dq_ld + 8 * c) = u;
  }
}

// Same, plus the column sums of dq (the q third of the qkv bias gradient).  Block = rpp rows x (D / 8) chunks of 8 columns; a thread
// keeps four rows in flight (8 x 16-byte loads), accumulates its 8 columns in registers, the rpp row lanes are combined through shared
// memory and one lane per column issues the atomic.  `flat`: dq_bs == N * dq_ld (the q third of a packed qkv gradient), no row split.
__global__ void __launch_bounds__(384) dq_finish_colsum_kernel(const float* __restrict__ dq32, __nv_bfloat16* __restrict__ dq, long long dq_ld,
                                                               long long dq_bs, int N, int D, float scale, int rows, int rows_per_block,
                                                               int flat, float* __restrict__ colsum) {
  extern __shared__ float s_sum[];   // [rpp][D]
  const int c8n = D >> 3, rpp = blockDim.x / c8n;
  const int c = threadIdx.x % c8n, rr = threadIdx.x / c8n;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (rr < rpp) {
    for (int row = r0 + rr; row < r1; row += 4 * rpp) {
      float4 x[4], y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rw = row + u * rpp;
        if (rw < r1) {
          const float* src = dq32 + (long long)rw * D + 8 * c;
          x[u] = *reinterpret_cast<const float4*>(src);
          y[u] = *reinterpret_cast<const float4*>(src + 4);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rw = row + u * rpp;
        if (rw < r1) {
          const float f[8] = {x[u].x * scale, x[u].y * scale, x[u].z * scale, x[u].w * scale, y[u].x * scale, y[u].y * scale, y[u].z * scale, y[u].w * scale};
          uint4 o;
          o.x = pack_bf16(f[0], f[1]);
          o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]);
          o.w = pack_bf16(f[6], f[7]);
          long long off;
          if (flat) {
            off = (long long)rw * dq_ld;
          } else {
            const int bb = rw / N;
            off = (long long)bb * dq_bs + (long long)(rw - bb * N) * dq_ld;
          }
          *reinterpret_cast<uint4*>(dq + off + 8 * c) = o;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += f[e];
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s_sum[rr * D + 8 * c + e] = acc[e];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < rpp; ++q) t += s_sum[q * D + j];
    atomicAdd(colsum + j, t);
  }
}

// ---- host side ---------------------------------------------------------------------------------------
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_map(CUtensorMap* m, const void* ptr, long long ld, long long bs, int B, int H, int N, const char* name, int box_rows) {
  ensure_context();
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return T4S_ERR_CUDA;
  }
  if (!ptr || (reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8) || (bs % 8) || ld < (long long)H * kHd || bs <= 0) {
    set_error("t4s_attn: %s needs a 16-byte aligned base, pitches that are multiples of 8 elements and ld >= heads*64 (ld=%lld, bs=%lld)",
              name, ld, bs);
    return T4S_ERR_ARG;
  }
  cuuint64_t dims[4] = {(cuuint64_t)kHd, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)(ld * 2), (cuuint64_t)(kHd * 2), (cuuint64_t)(bs * 2)};
  cuuint32_t box[4] = {(cuuint32_t)kHd, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (N=%d H=%d B=%d ld=%lld bs=%lld)", name, (int)rc, N, H, B, ld, bs);
    return T4S_ERR_CUDA;
  }
  return T4S_OK;
}

int launch_delta(const void* o, long long o_ld, long long o_bs, const float* o32, const void* d_o, long long do_ld, long long do_bs, float* delta, int B,
                 int H, int N, int Nl, cudaStream_t st) {
  const long long warps = (long long)B * Nl;
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  attn_delta_kernel<<<(unsigned)blocks, threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(o), o_ld, o_bs, o32,
                                                          reinterpret_cast<const __nv_bfloat16*>(d_o), do_ld, do_bs, delta, B, H, N, Nl);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

static int check_common(const T4sAttn* p) {
  T4S_REQUIRE(p, "t4s_attn: null descriptor");
  T4S_REQUIRE(p->head_dim == kHd, "t4s_attn: head_dim must be 64 (got %d)", p->head_dim);
  T4S_REQUIRE(p->batch > 0 && p->heads > 0 && p->tokens > 0, "t4s_attn: batch, heads and tokens must be positive");
  T4S_REQUIRE(p->batch <= 65535 && p->heads <= 65535, "t4s_attn: batch / heads exceed the grid limits");
  T4S_REQUIRE(p->q && p->k && p->v && p->o && p->lse, "t4s_attn: q, k, v, o and lse are required");
  T4S_REQUIRE(!(reinterpret_cast<uintptr_t>(p->o) & 15) && !(p->o_ld % 8) && !(p->o_bs % 8), "t4s_attn: o must be 16-byte aligned with pitches % 8 == 0");
  return T4S_OK;
}

static void fill_args(Args& a, const T4sAttn* p) {
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
  a.colsum = nullptr;
}

}  // namespace attn
}  // namespace t4s

#ifdef T4S_TRACE
extern "C" int t4s_debug_trace(long long* out, int n) {
  if (n > 4096) n = 4096;
  T4S_CUDA(cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * n));
  return T4S_OK;
}
#endif

extern "C" int64_t t4s_attn_padded_len(int tokens) {
  return (int64_t)((tokens + t4s::attn::kTile - 1) / t4s::attn::kTile) * t4s::attn::kTile;
}

extern "C" int t4s_attn_fwd(const T4sAttn* p, void* stream) {
  using namespace t4s::attn;
  int rc = check_common(p);
  if (rc) return rc;
  fwd3::Maps tm;
  if ((rc = make_map(&tm.q, p->q, p->q_ld, p->q_bs, p->batch, p->heads, p->tokens, "q"))) return rc;
  if ((rc = make_map(&tm.k, p->k, p->k_ld, p->k_bs, p->batch, p->heads, p->tokens, "k"))) return rc;
  if ((rc = make_map(&tm.v, p->v, p->v_ld, p->v_bs, p->batch, p->heads, p->tokens, "v"))) return rc;
  Args a;
  fill_args(a, p);
  T4S_CUDA(cudaFuncSetAttribute(fwd3::attn_fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd3::kSmem));
  dim3 grid((a.n_tiles + 1) / 2, p->heads, p->batch);
  fwd3::attn_fwd3_kernel<<<grid, fwd3::kThreads, fwd3::kSmem, t4s::as_stream(stream)>>>(tm, a);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

extern "C" int t4s_attn_bwd(const T4sAttnBwd* p, void* stream) {
  using namespace t4s::attn;
  T4S_REQUIRE(p, "t4s_attn_bwd: null descriptor");
  const T4sAttn* f = &p->fwd;
  int rc = check_common(f);
  if (rc) return rc;
  T4S_REQUIRE(p->d_o && p->delta && p->dq && p->dk && p->dv, "t4s_attn_bwd: d_o, delta, dq, dk and dv are required");
  for (const void* ptr : {(const void*)p->dq, (const void*)p->dk, (const void*)p->dv})
    T4S_REQUIRE(!(reinterpret_cast<uintptr_t>(ptr) & 15), "t4s_attn_bwd: dq / dk / dv must be 16-byte aligned");
  T4S_REQUIRE(!(p->dq_ld % 8) && !(p->dq_bs % 8) && !(p->dk_ld % 8) && !(p->dk_bs % 8) && !(p->dv_ld % 8) && !(p->dv_bs % 8),
              "t4s_attn_bwd: gradient pitches must be multiples of 8 elements");
  CUtensorMap tq, tk, tv, tdo;
  if ((rc = make_map(&tq, f->q, f->q_ld, f->q_bs, f->batch, f->heads, f->tokens, "q"))) return rc;
  if ((rc = make_map(&tk, f->k, f->k_ld, f->k_bs, f->batch, f->heads, f->tokens, "k"))) return rc;
  if ((rc = make_map(&tv, f->v, f->v_ld, f->v_bs, f->batch, f->heads, f->tokens, "v"))) return rc;
  if ((rc = make_map(&tdo, p->d_o, p->do_ld, p->do_bs, f->batch, f->heads, f->tokens, "d_o"))) return rc;
  Args a;
  fill_args(a, f);
  a.delta = p->delta;
  a.dq = reinterpret_cast<__nv_bfloat16*>(p->dq); a.dq_ld = p->dq_ld; a.dq_bs = p->dq_bs;
  a.dk = reinterpret_cast<__nv_bfloat16*>(p->dk); a.dk_ld = p->dk_ld; a.dk_bs = p->dk_bs;
  a.dv = reinterpret_cast<__nv_bfloat16*>(p->dv); a.dv_ld = p->dv_ld; a.dv_bs = p->dv_bs;
  cudaStream_t st = t4s::as_stream(stream);
  if ((rc = launch_delta(f->o, f->o_ld, f->o_bs, f->o32, p->d_o, p->do_ld, p->do_bs, p->delta, f->batch, f->heads, f->tokens, a.Nl, st))) return rc;
  dim3 grid(a.n_tiles, f->heads, f->batch);
  static const bool fused_enabled = [] { const char* e = getenv("T4S_ATTN_FUSED_BWD"); return !(e && e[0] == '0'); }();
  if (p->dq32 && fused_enabled) {
    // one kernel: dK / dV as before, dQ tiles reduce-added (fp32, TMA) into dq32, then scaled and rounded into the dq slot
    const int D = f->heads * kHd;
    T4S_REQUIRE(!(reinterpret_cast<uintptr_t>(p->dq32) & 15), "t4s_attn_bwd: dq32 must be 16-byte aligned");
    T4S_REQUIRE(!p->dqkv_colsum || f->heads * kHd / 8 <= 256, "t4s_attn_bwd: dqkv_colsum supports up to 32 heads");
    a.colsum = p->dqkv_colsum;
    CUtensorMap tdq;
    {
      t4s::ensure_context();
      EncodeTiledFn enc = get_encode();
      T4S_REQUIRE(enc, "cuTensorMapEncodeTiled entry point not available");
      cuuint64_t dims[4] = {(cuuint64_t)kHd, (cuuint64_t)f->tokens, (cuuint64_t)f->heads, (cuuint64_t)f->batch};
      cuuint64_t strides[3] = {(cuuint64_t)D * 4, (cuuint64_t)kHd * 4, (cuuint64_t)f->tokens * D * 4};
      cuuint32_t box[4] = {32, 32, 1, 1}, estr[4] = {1, 1, 1, 1};
      CUresult cr = enc(&tdq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p->dq32, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      T4S_REQUIRE(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dq32) failed with CUresult %d", (int)cr);
    }
    // dK / dV leave through TMA tile stores: (d, token, head, clip) bf16 views with a [32 x 32] SWIZZLE_64B box
    CUtensorMap tdk, tdv;
    {
      EncodeTiledFn enc = get_encode();
      const struct { CUtensorMap* m; void* ptr; long long ld, bs; const char* name; } outs[2] = {{&tdk, p->dk, p->dk_ld, p->dk_bs, "dk"},
                                                                                                 {&tdv, p->dv, p->dv_ld, p->dv_bs, "dv"}};
      for (const auto& o : outs) {
        T4S_REQUIRE(o.ptr && !(reinterpret_cast<uintptr_t>(o.ptr) & 15) && o.ld % 8 == 0 && o.bs % 8 == 0 && o.ld >= (long long)f->heads * kHd,
                    "t4s_attn_bwd: %s needs a 16-byte aligned base and pitches that are multiples of 8 elements", o.name);
        cuuint64_t dims[4] = {(cuuint64_t)kHd, (cuuint64_t)f->tokens, (cuuint64_t)f->heads, (cuuint64_t)f->batch};
        cuuint64_t strides[3] = {(cuuint64_t)(o.ld * 2), (cuuint64_t)(kHd * 2), (cuuint64_t)(o.bs * 2)};
        cuuint32_t box[4] = {32, 32, 1, 1}, estr[4] = {1, 1, 1, 1};
        CUresult cr = enc(o.m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, o.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        T4S_REQUIRE(cr == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", o.name, (int)cr);
      }
    }
    T4S_CUDA(cudaMemsetAsync(p->dq32, 0, (size_t)f->batch * f->tokens * D * sizeof(float), st));
    T4S_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fbw::kSmem));
    const int n_items = a.n_tiles * f->heads * f->batch;
    const int pgrid = std::min(n_items, t4s::sm_count());
    attn_bwd_fused_kernel<<<pgrid, fbw::kThreads, fbw::kSmem, st>>>(tq, tk, tv, tdo, tdq, tdk, tdv, a, n_items);
    T4S_LAUNCH_CHECK();
    if (p->dqkv_colsum) {
      const long long rows = (long long)f->batch * f->tokens;
      T4S_REQUIRE(rows < (1LL << 31) && D <= 3072, "t4s_attn_bwd: batch * tokens / width exceed the dq-finish kernel's limits");
      const int c8n = D / 8, rpp = std::max(1, std::min(4, 384 / c8n));
      const int rpb = 16 * rpp;    // four passes of four rows per thread
      dq_finish_colsum_kernel<<<(int)((rows + rpb - 1) / rpb), c8n * rpp, (size_t)rpp * D * sizeof(float), st>>>(
          p->dq32, a.dq, a.dq_ld, a.dq_bs, f->tokens, D, a.scale, (int)rows, rpb, a.dq_bs == (long long)f->tokens * a.dq_ld ? 1 : 0, p->dqkv_colsum);
      T4S_LAUNCH_CHECK();
      return T4S_OK;
    }
    const long long total8 = (long long)f->batch * f->tokens * D / 8;
    T4S_REQUIRE(total8 < (1LL << 31), "t4s_attn_bwd: batch * tokens * width exceeds the dq-finish kernel's limit");
    const int fgrid = (int)std::min<long long>((total8 + 255) / 256, (long long)t4s::sm_count() * 16);
    dq_finish_kernel<<<fgrid, 256, 0, st>>>(p->dq32, a.dq, a.dq_ld, a.dq_bs, f->tokens, D, a.scale, total8);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_REQUIRE(!p->dqkv_colsum, "t4s_attn_bwd: dqkv_colsum needs the one-kernel backward (dq32 workspace)");
  T4S_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmem));
  attn_bwd_kernel<0><<<grid, bwd::kThreads, bwd::kSmem, st>>>(tq, tk, tv, tdo, tq, a);
  T4S_LAUNCH_CHECK();
  T4S_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd::kSmem));
  attn_bwd_kernel<1><<<grid, bwd::kThreads, bwd::kSmem, st>>>(tq, tk, tv, tdo, tq, a);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
