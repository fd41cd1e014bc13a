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

constexpr int kPwBytes = 256 * 128;    // position window: 256 rows x 64 bf16
constexpr uint32_t kIdescBD = ptx::umma_idesc(1, 128, 256, 0, 0);
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
// operand region (112 KB): dQ  kernel: resident [QU][QV][dO], streamed [K][V][Pw]
//                          dKV kernel: resident [K][V],       streamed [QU][QV][dO][Pw]
// Skew scratch: every thread parks a 48-column BD window of its row (pitch 50 floats: 8-byte stores and 4-byte skewed reads are bank
// conflict free) and reads 16 shifted values back, four times per tile.  dQ kernel: the P tile's place holds the dBD staging rows.
constexpr int kBThreads = 320;
constexpr int kPitch = 50;
constexpr int kScr = 32 * kPitch * 4;   // 6400 B per warp
constexpr int oOps = 0, oP = oOps + 7 * kTileBytes, oDs = oP + kPBytes, oScr = oDs + kPBytes, oBar = oScr + 8 * kScr;
constexpr int kSmem = oBar + 128;
static_assert(kSmem <= 232448, "rel-pos attention backward: shared memory");
constexpr int kTmemCols = 512;  // S / dP: [0,128)  BD: [128,384)  acc0: [384,448)  acc1: [448,512)
enum { bResFull = 0, bStrFull = 1, bStrEmpty = 2, bSFull = 3, bPReady = 4, bDpFull = 5, bDsFull = 6, bFin = 7, bAccFull = 8, bCount = 9 };
}  // namespace bwd

template <bool kDq>
__global__ void __launch_bounds__(bwd::kBThreads, 1)
relattn_bwd_kernel(const __grid_constant__ CUtensorMap tmQU, const __grid_constant__ CUtensorMap tmQV,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmPos, const RelArgs ra) {
  using namespace bwd;
  const Args& a = ra.a;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + bCount);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * kTile, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = a.n_tiles;
  const long long stat_base = ((long long)b * a.H + h) * a.Nl;
  // operand tiles
  unsigned char* sQU = smem + oOps + (kDq ? 0 : 2) * kTileBytes;
  unsigned char* sQV = smem + oOps + (kDq ? 1 : 3) * kTileBytes;
  unsigned char* sDO = smem + oOps + (kDq ? 2 : 4) * kTileBytes;
  unsigned char* sK = smem + oOps + (kDq ? 3 : 0) * kTileBytes;
  unsigned char* sV = smem + oOps + (kDq ? 4 : 1) * kTileBytes;
  unsigned char* sPw = smem + oOps + 5 * kTileBytes;

  if (threadIdx.x == 0) {
    if (ptx::smem_u32(smem) & 1023u) {
      printf("t4s relattn_bwd: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    ptx::mbar_init(&bars[bResFull], 1);
    ptx::mbar_init(&bars[bStrFull], 1);
    ptx::mbar_init(&bars[bStrEmpty], 1);
    ptx::mbar_init(&bars[bSFull], 1);
    ptx::mbar_init(&bars[bPReady], 8);
    ptx::mbar_init(&bars[bDpFull], 1);
    ptx::mbar_init(&bars[bDsFull], 8);
    ptx::mbar_init(&bars[bFin], 1);
    ptx::mbar_init(&bars[bAccFull], 1);
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
    if (ptx::elect_one()) {
      if (kDq) {
        ptx::mbar_arrive_expect_tx(&bars[bResFull], 3 * kTileBytes);
        ptx::tma_load_4d(sQU, &tmQU, &bars[bResFull], 0, t0, h, b);
        ptx::tma_load_4d(sQV, &tmQV, &bars[bResFull], 0, t0, h, b);
        ptx::tma_load_4d(sDO, &tmDO, &bars[bResFull], 0, t0, h, b);
      } else {
        ptx::mbar_arrive_expect_tx(&bars[bResFull], 2 * kTileBytes);
        ptx::tma_load_4d(sK, &tmK, &bars[bResFull], 0, t0, h, b);
        ptx::tma_load_4d(sV, &tmV, &bars[bResFull], 0, t0, h, b);
      }
      for (int t = 0; t < n_tiles; ++t) {
        ptx::mbar_wait(&bars[bStrEmpty], (t & 1) ^ 1);
        const int i0 = kDq ? t0 : t * kTile, j0 = kDq ? t * kTile : t0;
        if (kDq) {
          ptx::mbar_arrive_expect_tx(&bars[bStrFull], 2 * kTileBytes + kPwBytes);
          ptx::tma_load_4d(sK, &tmK, &bars[bStrFull], 0, j0, h, b);
          ptx::tma_load_4d(sV, &tmV, &bars[bStrFull], 0, j0, h, b);
        } else {
          ptx::mbar_arrive_expect_tx(&bars[bStrFull], 3 * kTileBytes + kPwBytes);
          ptx::tma_load_4d(sQU, &tmQU, &bars[bStrFull], 0, i0, h, b);
          ptx::tma_load_4d(sQV, &tmQV, &bars[bStrFull], 0, i0, h, b);
          ptx::tma_load_4d(sDO, &tmDO, &bars[bStrFull], 0, i0, h, b);
        }
        ptx::tma_load_4d(sPw, &tmPos, &bars[bStrFull], 0, a.N - kTile - i0 + j0, h, 0);
      }
    }
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    const uint32_t uQU = ptx::smem_u32(sQU), uQV = ptx::smem_u32(sQV), uDO = ptx::smem_u32(sDO), uK = ptx::smem_u32(sK),
                   uV = ptx::smem_u32(sV), uPw = ptx::smem_u32(sPw), uP = ptx::smem_u32(smem + oP), uDs = ptx::smem_u32(smem + oDs);
    ptx::mbar_wait(&bars[bResFull], 0);
    for (int t = 0; t < n_tiles; ++t) {
      ptx::mbar_wait(&bars[bStrFull], t & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        mma_k64(tmem, uQU, uK, kIdescS, false);          // AC = (q+u) k^T
        mma_k64(tmem + 128, uQV, uPw, kIdescBD, false);  // BD = (q+v) Pw^T
        ptx::tc_commit(&bars[bSFull]);
      }
      __syncwarp();
      ptx::mbar_wait(&bars[bPReady], t & 1);             // P is in registers: the S columns are free
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        mma_k64(tmem, uDO, uV, kIdescS, false);          // dP = dO v^T
        ptx::tc_commit(&bars[bDpFull]);
      }
      __syncwarp();
      ptx::mbar_wait(&bars[bDsFull], t & 1);             // P / dS tiles are in shared memory
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        if (kDq) {
          mma_k128_mn(tmem + 384, uDs, uK, kIdescPV, t > 0);   // d(q+u) += dS K
        } else {
          mma_k128_amn(tmem + 384, uP, uDO, t > 0);            // dV += P^T dO
          mma_k128_amn(tmem + 448, uDs, uQU, t > 0);           // dK += dS^T (q+u)
        }
        ptx::tc_commit(&bars[bStrEmpty]);
        ptx::tc_commit(&bars[bFin]);
        if (t == n_tiles - 1) ptx::tc_commit(&bars[bAccFull]);
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax warps: thread = (query row, column half) ----------------
    const int wq = warp & 3, g = warp >> 2;
    const int r = wq * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
    float* scr = reinterpret_cast<float*>(smem + oScr + warp * kScr) + lane * kPitch;
    unsigned char* stage = smem + oP + warp * 4096;          // dQ kernel: 32 staged dS rows x 128 B (this warp's 64 columns)
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
      ptx::mbar_wait(&bars[bSFull], t & 1);
      ptx::tc_fence_after();
      // phase A: P = exp2((AC + shift(BD)) * c - lse), kept packed in registers
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
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // s[16 q + cc] += BD[r][127 - r + 64 g + 16 q + cc]: a 48-column window of this warp's rows, read back at the lane's offset
        const int wb = 96 - 32 * wq + 64 * g + 16 * q;
        uint32_t x0[32], x1[16];
        ptx::tmem_ld_32x32(t_lane + 128 + wb, x0);
        ptx::tmem_ld_32x16(t_lane + 128 + wb + 32, x1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) *reinterpret_cast<uint2*>(scr + 2 * k) = make_uint2(x0[2 * k], x0[2 * k + 1]);
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<uint2*>(scr + 32 + 2 * k) = make_uint2(x1[2 * k], x1[2 * k + 1]);
        const volatile float* rd = scr + (31 - lane);
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) s[16 * q + cc] += rd[cc];
      }
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
      if (lane == 0) ptx::mbar_arrive(&bars[bPReady]);
      // the previous step's MMAs have finished with the P / dS tiles
      ptx::mbar_wait(&bars[bFin], (t & 1) ^ 1);
      if (!kDq) {
        uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[0]);
        uint32_t (&hi)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[16]);
        store_row_chunk(smem + oP, r, 64 * g, lo);
        store_row_chunk(smem + oP, r, 64 * g + 32, hi);
      }
      // phase B: dS = P (dP - delta)
      ptx::mbar_wait(&bars[bDpFull], t & 1);
      ptx::tc_fence_after();
      uint32_t pd[32];
      {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32(t_lane + 64 * g, v0);
        ptx::tmem_ld_32x32(t_lane + 64 * g + 32, v1);
        ptx::tmem_ld_wait();
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
      {
        uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pd[0]);
        uint32_t (&hi)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pd[16]);
        store_row_chunk(smem + oDs, r, 64 * g, lo);
        store_row_chunk(smem + oDs, r, 64 * g + 32, hi);
      }
      if (kDq) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stage + lane * 128 + q * 16) = make_uint4(pd[4 * q], pd[4 * q + 1], pd[4 * q + 2], pd[4 * q + 3]);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[bDsFull]);
      if (kDq) {
        // dBD[b, h, i, T-1-i+j0 + 64 g + e] <- dS[i, j0 + 64 g + e]: this warp's 32 staged rows, 32 destination words per row
        const uint32_t* st32 = reinterpret_cast<const uint32_t*>(stage);
        const int nv = min(nvalid, 64);
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const int i = i0 + wq * 32 + rr;
          if (i >= a.N) break;
          const int x0 = a.N - 1 - i + j0 + 64 * g;
          __nv_bfloat16* drow = ra.dbd + (((long long)b * a.H + h) * a.N + i) * ra.dbd_ld + x0;
          const int par = x0 & 1;  // rows start 16-byte aligned, so the word alignment of the destination is the parity of x0
          const uint32_t* srow = st32 + rr * 32;
          const int e = par + 2 * lane;  // first source element of destination word `lane`
          const uint32_t nxt = (lane < 31) ? srow[lane + 1] : 0u;
          const uint32_t val = __funnelshift_r(srow[lane], nxt, 16 * par);
          if (e + 1 < nv) *reinterpret_cast<uint32_t*>(drow + e) = val;
          else if (e < nv) *reinterpret_cast<unsigned short*>(drow + e) = (unsigned short)(val & 0xffffu);
          if (par && lane == 0 && nv > 0) *reinterpret_cast<unsigned short*>(drow) = (unsigned short)(srow[0] & 0xffffu);
        }
        __syncwarp();
      }
    }
    // ---- accumulators: each column half writes 32 of the 64 head-dim columns ----
    ptx::mbar_wait(&bars[bAccFull], 0);
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
  if ((rc = make_map(&tpos, f->pos, f->pos_ld, f->pos_ld * (2LL * T - 1), 1, H, 2 * T - 1, "pos", 256))) return rc;
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
