// Shared device helpers of the fused attention kernels (attn.cu: plain MHSA, attn_rel.cu: Transformer-XL rel-pos MHSA).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

// Optional pipeline trace (build with -DT4S_TRACE: scripts/trace_attn.sh): one CTA records clock64() at the hand-over points of
// its warps; read back with t4s_debug_trace().  Compiled out of the product library.
#ifdef T4S_TRACE
static __device__ long long g_trace[4096];
#define T4S_TRACE_AT(w, j, e)                                                                                  \
  do {                                                                                                         \
    if (blockIdx.x == 2 && blockIdx.y == 0 && blockIdx.z == 5 && (threadIdx.x & 31) == 0 && (j) < 16)           \
      g_trace[((w) * 16 + (j)) * 8 + (e)] = clock64();                                                         \
  } while (0)
#else
#define T4S_TRACE_AT(w, j, e) do {} while (0)
#endif

namespace t4s {
namespace attn {

constexpr int kHd = 64;
constexpr int kTile = 128;
constexpr int kTileBytes = kTile * kHd * 2;  // one [128 rows x 64] bf16 SWIZZLE_128B tile
constexpr int kPBytes = 2 * kTileBytes;      // one [128 x 128] bf16 tile = two 64-column sub-tiles

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a packed pair on the FMA / ALU pipes instead of MUFU (head size 64 attention is bound by the 16 exp2 / clk / SM of the
// special-function unit: 1024 clk per 128 x 128 tile against 512 clk of MMA, so a fixed share of the exponentials is moved
// here).  Cody-Waite split by the 1.5 * 2^23 rounding trick, degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (relative error
// 7.5e-5, far below the bf16 rounding of P), exponent inserted with an integer shift-add.  Arguments below -126 are clamped.
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
  const uint64_t magic = ptx::pack2(12582912.f, 12582912.f);
  const uint64_t x = ptx::pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t = ptx::add2(x, magic);
  const uint64_t f = ptx::sub2(x, ptx::sub2(t, magic));
  uint64_t p = ptx::fma2(f, ptx::pack2(0.055171654f, 0.055171654f), ptx::pack2(0.24261113f, 0.24261113f));
  p = ptx::fma2(p, f, ptx::pack2(0.69326097f, 0.69326097f));
  p = ptx::fma2(p, f, ptx::pack2(0.99992806f, 0.99992806f));
  float ta, tb, pa, pb;
  ptx::unpack2(t, ta, tb);
  ptx::unpack2(p, pa, pb);
  p0 = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(ta) << 23));
  p1 = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(tb) << 23));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// Row r, columns [col0, col0 + 32) of a K-major [128 x 128] bf16 operand tile (two SWIZZLE_128B sub-tiles of 64 columns).
__device__ __forceinline__ void store_row_chunk(unsigned char* tile, int r, int col0, const uint32_t (&pk)[16]) {
  unsigned char* base = tile + (col0 >> 6) * kTileBytes + r * 128;
  const int ch0 = (col0 & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int ch = (ch0 + q) ^ (r & 7);
    *reinterpret_cast<uint4*>(base + ch * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
// D[128 x N] (+)= A[128 x 64] . B[N x 64]^T, both K-major tiles (4 instructions of K = 16)
__device__ __forceinline__ void mma_k64(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool accumulate) {
  const uint64_t adesc = ptx::umma_desc_sw128(a_addr, 16, 1024), bdesc = ptx::umma_desc_sw128(b_addr, 16, 1024);
#pragma unroll
  for (int k = 0; k < 4; ++k) ptx::mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
}
// D[128 x 64] (+)= A[128 x 128] . B, A a K-major [128 x 128] tile written by the softmax warps, B a [128 rows(K) x 64] tile
// consumed MN-major (8 instructions of K = 16 rows = 2048 bytes each).
__device__ __forceinline__ void mma_k128_mn(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool accumulate) {
  const uint64_t bdesc = ptx::umma_desc_sw128(b_addr, 8192, 1024);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t adesc = ptx::umma_desc_sw128(a_addr + (k >> 2) * kTileBytes, 16, 1024) + 2 * (k & 3);
    ptx::mma_f16(d_tmem, adesc, bdesc + 128 * k, idesc, (accumulate || k > 0) ? 1u : 0u);
  }
}
constexpr uint32_t kIdescS = ptx::umma_idesc(1, 128, 128, 0, 0);   // 128 x 128, both K-major
constexpr uint32_t kIdescPV = ptx::umma_idesc(1, 128, 64, 0, 1);   // 128 x 64, B MN-major

struct Args {
  int N, Nl, n_tiles, H;
  float sl2;    // scale * log2(e)
  float scale;
  __nv_bfloat16* o;  long long o_ld, o_bs;
  float* lse;        // [B, H, Nl], log2 units
  float* o32;        // optional fp32 copy of o, [B, N, H*64] contiguous
  const float* delta;
  __nv_bfloat16* dq; long long dq_ld, dq_bs;
  __nv_bfloat16* dk; long long dk_ld, dk_bs;
  __nv_bfloat16* dv; long long dv_ld, dv_bs;
  float* colsum;     // optional [3 * H * 64]: column sums of dq | dk | dv (qkv bias gradient)
};

// Column sums of a warp's 32 x 32 fp32 chunk (thread = row, v = its 32 columns) by a butterfly reduce-scatter over the lanes
// (31 shuffles); lane c returns the sum of column c.  `ok` = this lane's row exists.
__device__ __forceinline__ float warp_colsum32(const uint32_t (&v)[32], bool ok, int lane) {
  float y[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) y[i] = ok ? __uint_as_float(v[i]) : 0.f;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? y[i] : y[i + o], keep = up ? y[i + o] : y[i];
      y[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return y[0];
}

__device__ __forceinline__ void store_row64(__nv_bfloat16* dst, const float (&v)[64], float mul) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 u;
    u.x = pack_bf16(v[8 * q] * mul, v[8 * q + 1] * mul);
    u.y = pack_bf16(v[8 * q + 2] * mul, v[8 * q + 3] * mul);
    u.z = pack_bf16(v[8 * q + 4] * mul, v[8 * q + 5] * mul);
    u.w = pack_bf16(v[8 * q + 6] * mul, v[8 * q + 7] * mul);
    reinterpret_cast<uint4*>(dst)[q] = u;
  }
}
__device__ __forceinline__ void store_row64_f32(float* dst, const float (&v)[64], float mul) {
#pragma unroll
  for (int q = 0; q < 16; ++q)
    reinterpret_cast<float4*>(dst)[q] = make_float4(v[4 * q] * mul, v[4 * q + 1] * mul, v[4 * q + 2] * mul, v[4 * q + 3] * mul);
}
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const uint32_t (&v)[32], float mul) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16(__uint_as_float(v[8 * q]) * mul, __uint_as_float(v[8 * q + 1]) * mul);
    u.y = pack_bf16(__uint_as_float(v[8 * q + 2]) * mul, __uint_as_float(v[8 * q + 3]) * mul);
    u.z = pack_bf16(__uint_as_float(v[8 * q + 4]) * mul, __uint_as_float(v[8 * q + 5]) * mul);
    u.w = pack_bf16(__uint_as_float(v[8 * q + 6]) * mul, __uint_as_float(v[8 * q + 7]) * mul);
    reinterpret_cast<uint4*>(dst)[q] = u;
  }
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
// (d, token, head, clip) bf16 view with a [64 x box_rows x 1 x 1] SWIZZLE_128B box
// delta[b, h, n] = sum_d dO * O (0 for the padded rows n >= N)
int launch_delta(const void* o, long long o_ld, long long o_bs, const float* o32, const void* d_o, long long do_ld, long long do_bs, float* delta, int B,
                 int H, int N, int Nl, cudaStream_t st);
int make_map(CUtensorMap* m, const void* ptr, long long ld, long long bs, int B, int H, int N, const char* name, int box_rows = kTile);

}  // namespace attn
}  // namespace t4s
