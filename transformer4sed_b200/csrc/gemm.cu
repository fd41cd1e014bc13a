// K3: persistent warp-specialised tcgen05 GEMM with TMA-fed shared-memory pipeline and a fused epilogue.
//
//   C[z] = act(alpha * A[z] . B[z]^T + bias) + residual[z]      A:[M,K]  B:[N,K]  (K- or MN-major), fp32 accumulate in TMEM
//
// CTA = 12 warps, one CTA per SM, walking 128 x BN output tiles in a grouped n-fastest order (a wave streams each A row-block
// once and keeps its B blocks in L2):
//   warp 0    TMA producer: 4-D tensor maps (k, row, batch1, batch2), SWIZZLE_128B boxes -> 4..6-stage smem ring
//   warp 1    MMA issuer:   one lane issues tcgen05.mma (M=128, N=BN, K=32 B per instruction), commits to mbarriers
//   warp 2    TMEM allocator (2 x BN fp32 columns: accumulator double buffer, epilogue overlaps the next mainloop)
//   warps 4-11 epilogue: two warps per TMEM lane quarter, each draining half of the tile's columns in 32-column chunks:
//             tcgen05.ld 32x32b (next chunk in flight while this one is processed) -> bias / GELU / GELU' / residual in registers
//             (packed f32x2 math) -> staged variant (bf16 outputs): the chunk goes to a per-warp SWIZZLE_64B shared-memory box
//             and leaves through ONE TMA tile store, the residual / pre-activation chunk arrives the same way by TMA load;
//             direct variant (fp32 outputs, odd layouts): 16-byte row stores.  The accumulator buffer is handed back to the
//             MMA warp as soon as its last chunk is in registers.  Row / column predicates and TMA clipping / zero-fill make
//             ragged M/N tiles compose.
// Pair mode (bf16, BN = 256): the two CTAs of a cluster (one TPC) share one 256 x 256 tile through tcgen05.mma.cta_group::2 --
// each stages its 128 rows of A and half of the B tile, the leader issues the MMAs, commits are multicast to both CTAs.
// bf16 inputs use kind::f16, fp32 inputs use kind::tf32 (tensor map type TFLOAT32 rounds on load).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace t4s {
namespace gemm {

constexpr int kBM = 128;
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;

struct MatArg {
  void* ptr;
  long long ld, s1, s2;
  int dtype;
  int vec;  // 16-byte accesses are legal at every (row, column % 8 == 0) position
};

struct Args {
  int M, N, K, nb1, nb2;
  int a_b1, a_b2, b_b1, b_b2;  // 1 if the operand really has that batch level (else coordinate 0)
  int tiles_m, tiles_n;
  int group_n;                 // rasterisation: n-tiles per group (n fastest inside a group, then m, then the next group)
  int split_k, kb_per_split;   // K range of split s: k-blocks [s*kb_per_split, min(kblocks, (s+1)*kb_per_split))
  long long c_split;           // elements between partial outputs
  long long total_tiles;
  MatArg C, aux, res;
  const float* bias;
  int bias_vec;  // bias base is 16-byte aligned
  float alpha;
  int act;
  int res_tma;   // staged epilogue: the residual / pre-activation tile arrives by TMA (tensor map tmR) instead of per-row loads
  float* colsum; // staged epilogue: column sums of the output, accumulated with atomics
  int band_lo, band_hi;  // band_hi > 0: A[m, k] == 0 unless band_lo <= m + k < band_hi
};

template <int BN>
struct Cfg {
  static constexpr int kStageA = kBM * 128;
  static constexpr int kStageB = BN * 128;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmem = 1024 /*align slack*/ + kStages * kStage + 256 /*barriers*/;
  // staged epilogue: per epilogue warp two 2 KB SWIZZLE_64B boxes (32 rows x 32 bf16) for TMA tile stores
  static constexpr int kStageOff = kStages * kStage + 512;
  static constexpr int kSmemStaged = 1024 + kStageOff + kEpiWarps * 2 * 2048;
  // CTA-pair mode (cta_group::2, BN = 256): each CTA stages its 128 rows of A and HALF of the B tile (128 of the 256 N rows):
  // 32 KB per stage instead of 48, i.e. a third less operand traffic from L2 per FLOP and six stages in the same budget
  static constexpr int kStageB2 = kStageB / 2;
  static constexpr int kStage2 = kStageA + kStageB2;
  static constexpr int kStages2 = 6;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// ---- bf16-output GELU / GELU' on packed fp32 pairs (FFMA2 / FMUL2 / FADD2: two lanes per issue slot) ---------------------------
// Phi(-a) = exp2(-q(a)), q a degree-5 polynomial fitted to -log2(Phi(-a)) on [0, 6.5] (monotone beyond it): |error| of
// x*Phi(x) < 7e-7 absolute and < 3.3e-3 of the value down to x = -4, i.e. below bf16 resolution.  One MUFU and 4.5 issue slots
// per element; the Abramowitz-Stegun erfc it replaces cost 2 MUFU + ~15 slots and made the fc1 / fc2-dgrad epilogues the
// bottleneck of those GEMMs (tensor pipe 30 % active).  Used only when the output is bf16; fp32 outputs keep erff.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2f(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
#define T4S_K2(c) pack2(c, c)
// -q(a) - shift: exp2 of it is Phi(-a) * 2^-shift
__device__ __forceinline__ uint64_t neg_log2_tail(uint64_t a, float c0) {
  uint64_t q = fma2(a, T4S_K2(-4.73970344e-04f), T4S_K2(7.08921017e-03f));
  q = fma2(q, a, T4S_K2(-5.18392308e-02f));
  q = fma2(q, a, T4S_K2(-4.59979016e-01f));
  q = fma2(q, a, T4S_K2(-1.15079439e+00f));
  return fma2(q, a, T4S_K2(c0));
}
// x Phi(x) = relu(x) - a Phi(-a),  a = |x|
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const uint64_t a = pack2(fabsf(x0), fabsf(x1)), na = pack2(-fabsf(x0), -fabsf(x1)), x = pack2(x0, x1);
  float q0, q1;
  unpack2(neg_log2_tail(a, -1.00003661e+00f + 1.0f), q0, q1);   // exp2 -> 2 Phi(-a)
  const uint64_t e = pack2(ex2f(q0), ex2f(q1));
  const uint64_t y = mul2(fma2(na, e, add2(x, a)), T4S_K2(0.5f));
  unpack2(y, x0, x1);
}
// d/dx [x Phi(x)] = Phi(x) + x phi(x);  Phi(x) = 0.5 + copysign(0.5 - Phi(-a), x).  Multiplies g in place.
__device__ __forceinline__ void gelu_grad_fast2(float h0, float h1, float& g0, float& g1) {
  const uint64_t a = pack2(fabsf(h0), fabsf(h1)), x = pack2(h0, h1);
  float q0, q1;
  unpack2(neg_log2_tail(a, -1.00003661e+00f), q0, q1);
  const uint64_t tail = pack2(ex2f(q0), ex2f(q1));                 // Phi(-a)
  float u0, u1;
  unpack2(fma2(tail, T4S_K2(-1.0f), T4S_K2(0.5f)), u0, u1);        // 0.5 - Phi(-a) >= 0
  u0 = __uint_as_float(__float_as_uint(u0) | (__float_as_uint(h0) & 0x80000000u));
  u1 = __uint_as_float(__float_as_uint(u1) | (__float_as_uint(h1) & 0x80000000u));
  const uint64_t cdf = add2(pack2(u0, u1), T4S_K2(0.5f));
  float w0, w1;
  unpack2(mul2(mul2(a, T4S_K2(-0.72134752044448170f)), a), w0, w1);   // -a^2 / 2 in log2 units
  const uint64_t pdf = pack2(ex2f(w0), ex2f(w1));
  const uint64_t d = fma2(mul2(x, T4S_K2(0.3989422804014327f)), pdf, cdf);
  unpack2(mul2(pack2(g0, g1), d), g0, g1);
}
__device__ __forceinline__ float gelu_grad_erf(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// 32 consecutive columns of one output row, from / to registers.  `nvalid` = columns that exist (may exceed 32).
__device__ __forceinline__ void store32(const MatArg& m, long long off, int nvalid, const float (&x)[32]) {
  if (m.dtype == T4S_F32) {
    float* p = reinterpret_cast<float*>(m.ptr) + off;
    if (m.vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) p[i] = x[i];
    }
  } else {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(m.ptr) + off;
    if (m.vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(x[8 * i], x[8 * i + 1]);     u.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(x[8 * i + 2], x[8 * i + 3]); u.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(x[8 * i + 4], x[8 * i + 5]); u.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(x[8 * i + 6], x[8 * i + 7]); u.w = *reinterpret_cast<uint32_t*>(&t);
        reinterpret_cast<uint4*>(p)[i] = u;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) p[i] = __float2bfloat16_rn(x[i]);
    }
  }
}

__device__ __forceinline__ void load32(const MatArg& m, long long off, int nvalid, float (&x)[32]) {
  if (m.dtype == T4S_F32) {
    const float* p = reinterpret_cast<const float*>(m.ptr) + off;
    if (m.vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = reinterpret_cast<const float4*>(p)[i];
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = i < nvalid ? p[i] : 0.f;
    }
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(m.ptr) + off;
    if (m.vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = reinterpret_cast<const uint4*>(p)[i];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          x[8 * i + 2 * j] = f.x;
          x[8 * i + 2 * j + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = i < nvalid ? __bfloat162float(p[i]) : 0.f;
    }
  }
}

__device__ __forceinline__ void add32(const MatArg& m, long long off, int nvalid, float (&x)[32]) {
  if (m.dtype == T4S_F32) {
    const float* p = reinterpret_cast<const float*>(m.ptr) + off;
    if (m.vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = reinterpret_cast<const float4*>(p)[i];
        x[4 * i] += v.x; x[4 * i + 1] += v.y; x[4 * i + 2] += v.z; x[4 * i + 3] += v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) x[i] += p[i];
    }
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(m.ptr) + off;
    if (m.vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = reinterpret_cast<const uint4*>(p)[i];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          x[8 * i + 2 * j] += f.x;
          x[8 * i + 2 * j + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nvalid) x[i] += __bfloat162float(p[i]);
    }
  }
}

// One 32-column chunk of the epilogue for the row this thread owns.
__device__ __forceinline__ void epilogue_chunk(const Args& a, const uint32_t (&v)[32], int grow, int gcol, long long c_row,
                                               long long x_row, long long r_row) {
  const int nvalid = a.N - gcol;
  if (grow >= a.M || nvalid <= 0) return;
  float x[32];
  if (a.bias) {
    if (a.bias_vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + gcol) + i);
        x[4 * i] = fmaf(a.alpha, __uint_as_float(v[4 * i]), b4.x);
        x[4 * i + 1] = fmaf(a.alpha, __uint_as_float(v[4 * i + 1]), b4.y);
        x[4 * i + 2] = fmaf(a.alpha, __uint_as_float(v[4 * i + 2]), b4.z);
        x[4 * i + 3] = fmaf(a.alpha, __uint_as_float(v[4 * i + 3]), b4.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaf(a.alpha, __uint_as_float(v[i]), i < nvalid ? __ldg(a.bias + gcol + i) : 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = a.alpha * __uint_as_float(v[i]);
  }
  if (a.aux.ptr) store32(a.aux, x_row + gcol, nvalid, x);
  if (a.act == T4S_ACT_GELU) {
    if (a.C.dtype == T4S_BF16) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_fast2(x[i], x[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = gelu_erf(x[i]);
    }
  }
  if (a.act == T4S_ACT_GELU_GRAD) {
    float h[32];
    load32(a.res, r_row + gcol, nvalid, h);
    if (a.C.dtype == T4S_BF16) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_grad_fast2(h[i], h[i + 1], x[i], x[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] *= gelu_grad_erf(h[i]);
    }
  } else if (a.res.ptr) {
    add32(a.res, r_row + gcol, nvalid, x);
  }
  store32(a.C, c_row + gcol, nvalid, x);
}

// Staged variant of epilogue_chunk for bf16 outputs: the 32 x 32 chunk of a warp is written to a SWIZZLE_64B shared-memory box
// (16-byte stores, conflict free) and leaves through ONE TMA tile store, instead of 4 STG.128 per thread that each touch 32
// different rows (32 L1 wavefronts per instruction: with two outputs the fc1 epilogue needed more LSU cycles than the tile's
// MMAs).  Ragged edges are clipped by the tensor map.  `seq` alternates the warp's two staging boxes; lane 0 owns the bulk groups.
__device__ __forceinline__ void stage_store(const CUtensorMap* map, unsigned char* boxes, uint32_t& seq, int lane, const float (&x)[32],
                                            int gcol, int row0, int z1, int z2, bool single_box) {
  unsigned char* box = boxes + (seq & 1) * 2048;
  ++seq;
  if (ptx::elect_one()) {   // the store that last used this box has drained it (same elected lane commits and waits)
    if (single_box) ptx::bulk_wait_read_all();
    else ptx::bulk_wait_read_1();
  }
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(box + lane * 64);
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(x[8 * i], x[8 * i + 1]);     u.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[8 * i + 2], x[8 * i + 3]); u.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[8 * i + 4], x[8 * i + 5]); u.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[8 * i + 6], x[8 * i + 7]); u.w = *reinterpret_cast<uint32_t*>(&t);
    dst[i ^ sw] = u;
  }
  ptx::fence_proxy_async();
  __syncwarp();
  if (ptx::elect_one()) {
    ptx::tma_store_4d(map, box, gcol, row0, z1, z2);
    ptx::bulk_commit();
  }
}

// Pull the 64 bytes of this thread's residual row that chunk `gcol` will read into L2 ahead of time (the direct per-row loads
// are latency-bound otherwise: each one is a DRAM round trip inside the chunk's critical path).
__device__ __forceinline__ void prefetch_res(const Args& a, long long r_row, int gcol, bool row_ok) {
  if (a.res.ptr && a.res.dtype == T4S_BF16 && row_ok && gcol < a.N)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __nv_bfloat16*>(a.res.ptr) + r_row + gcol));
}

// warp-uniform control flow (every lane stages its row; rows / columns outside the matrix are clipped by the TMA store)
// 32 bf16 of this lane's row out of a SWIZZLE_64B box filled by a TMA load
__device__ __forceinline__ void read_box_row(const unsigned char* box, int lane, float (&h)[32]) {
  const uint4* src = reinterpret_cast<const uint4*>(box + lane * 64);
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = src[i ^ sw];
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(hh[j]);
      h[8 * i + 2 * j] = f.x;
      h[8 * i + 2 * j + 1] = f.y;
    }
  }
}

__device__ __forceinline__ void epilogue_chunk_staged(const Args& a, const CUtensorMap* mapC, const CUtensorMap* mapX, unsigned char* boxes,
                                                      uint32_t& seq, int lane, const uint32_t (&v)[32], int row0, int gcol, long long r_row,
                                                      int z1, int z2, const float (&rtile)[32], bool have_rtile /* residual chunk already in registers */) {
  const int nvalid = a.N - gcol;
  if (row0 >= a.M || nvalid <= 0) return;
  const bool row_ok = row0 + lane < a.M;
  float x[32];
  if (a.bias) {
    if (a.bias_vec && nvalid >= 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + gcol) + i);
        x[4 * i] = fmaf(a.alpha, __uint_as_float(v[4 * i]), b4.x);
        x[4 * i + 1] = fmaf(a.alpha, __uint_as_float(v[4 * i + 1]), b4.y);
        x[4 * i + 2] = fmaf(a.alpha, __uint_as_float(v[4 * i + 2]), b4.z);
        x[4 * i + 3] = fmaf(a.alpha, __uint_as_float(v[4 * i + 3]), b4.w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaf(a.alpha, __uint_as_float(v[i]), i < nvalid ? __ldg(a.bias + gcol + i) : 0.f);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = a.alpha * __uint_as_float(v[i]);
  }
  if (a.aux.ptr) stage_store(mapX, boxes, seq, lane, x, gcol, row0, z1, z2, a.res_tma != 0);
  if (a.act == T4S_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) gelu_fast2(x[i], x[i + 1]);
  }
  if (a.act == T4S_ACT_GELU_GRAD) {
    if (have_rtile) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_grad_fast2(rtile[i], rtile[i + 1], x[i], x[i + 1]);
    } else {
      float h[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) h[i] = 0.f;
      if (row_ok) load32(a.res, r_row + gcol, nvalid, h);
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_grad_fast2(h[i], h[i + 1], x[i], x[i + 1]);
    }
  } else if (a.res.ptr) {
    if (have_rtile) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] += rtile[i];
    } else if (row_ok) {
      add32(a.res, r_row + gcol, nvalid, x);
    }
  }
  if (a.colsum) {
    // column sums of this warp's 32 x 32 chunk: butterfly reduce-scatter over the lanes (31 shuffles), lane c ends with column c
    float y[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) y[i] = row_ok ? x[i] : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const float send = up ? y[i] : y[i + o], keep = up ? y[i + o] : y[i];
        y[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    if (lane < nvalid) atomicAdd(a.colsum + gcol + lane, y[0]);
  }
  stage_store(mapC, boxes, seq, lane, x, gcol, row0, z1, z2, a.res_tma != 0);
}

// Tile r of a batch -> (m tile, n tile).  Tiles that run concurrently (148 consecutive indices) form a compact
// group_n x (148 / group_n) rectangle, so a wave streams each A row-block once and keeps its few B blocks in L2; the
// m-fastest order this replaces re-read the activation operand from HBM once per n tile (12x for N = 3072).
__device__ __forceinline__ void tile_coords(const Args& a, int r, int& tm, int& tn) {
  const int group_tiles = a.group_n * a.tiles_m;
  const int g = r / group_tiles, rem = r - g * group_tiles;
  const int width = min(a.group_n, a.tiles_n - g * a.group_n);
  tm = rem / width;
  tn = g * a.group_n + (rem - tm * width);
}

// k-blocks [kb0, kb1) of split `sp` for the M tile `tm`: the split's share of K, clipped to the blocks that meet the tile's part of the
// anti-diagonal band when the caller declared one.  Never empty (an all-zero block stands in): the accumulator is always written.
__device__ __forceinline__ void k_range(const Args& a, int sp, int tm, int tile_rows, int kblocks_all, int bk, int& kb0, int& kb1) {
  kb0 = sp * a.kb_per_split;
  kb1 = min(kblocks_all, kb0 + a.kb_per_split);
  if (a.band_hi > 0) {
    const int m0 = tm * tile_rows, m1 = min(a.M, m0 + tile_rows) - 1;
    const int lo = max(0, a.band_lo - m1) / bk, hi = (min(a.K, a.band_hi - m0) + bk - 1) / bk;
    const int c0 = max(kb0, lo), c1 = min(kb1, hi);
    if (c1 > c0) {
      kb0 = c0;
      kb1 = c1;
    } else {
      kb1 = kb0 + 1;
    }
  }
}

// kPair: the CTAs of a 2-CTA cluster (one TPC) work on one 256 x BN tile with tcgen05.mma.cta_group::2: CTA rank r owns rows
// [256 tm + 128 r, +128) of A / C and stages rows [n0 + 128 r, +128) of B; the leader (rank 0) issues the MMAs for both, its
// `full` barriers collect the TMA bytes of both CTAs, commits are multicast to both CTAs' `empty` / `tfull` barriers, and the
// peer's epilogue warps hand their accumulator buffer back on the leader's `tempty` barrier.
template <int BN, bool kTf32, bool kAMn, bool kBMn, bool kStaged, bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
            const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR, const Args a) {
  using C = Cfg<BN>;
  static_assert(!kPair || (BN == 256 && !kTf32), "pair mode: bf16, BN = 256");
  constexpr int kBK = kTf32 ? 32 : 64;   // K elements per pipeline stage
  constexpr int kRowEl = kTf32 ? 32 : 64; // elements per 128-byte swizzle row
  constexpr int kBoxBytes = kBK * 128;    // one MN-major box: kBK rows (K) x 128 B (M/N)
  constexpr int kNStages = kPair ? C::kStages2 : C::kStages;
  constexpr int kStageBytesB = kPair ? C::kStageB2 : C::kStageB;
  constexpr int kStageBytes = C::kStageA + kStageBytesB;
  constexpr int kRowsB = kPair ? BN / 2 : BN;      // B rows staged by this CTA
  constexpr int kTileM = kPair ? 2 * kBM : kBM;    // rows of C per scheduled tile
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  unsigned char* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  unsigned char* sA = smem;
  unsigned char* sB = smem + kNStages * C::kStageA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kNStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kNStages;
  uint64_t* tfull = bars + 2 * kNStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* rbars = bars + 24;   // one per epilogue warp: residual-tile TMA loads (staged epilogue)
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  const long long tile_first = kPair ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long tile_step = kPair ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks_all = (a.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if (kStaged) {
      ptx::prefetch_tmap(&tmC);
      if (a.aux.ptr) ptx::prefetch_tmap(&tmX);
      if (a.res_tma) ptx::prefetch_tmap(&tmR);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kNStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], kPair ? 2 * kEpiWarps : kEpiWarps);
    }
    if (kStaged)
      for (int i = 0; i < kEpiWarps; ++i) ptx::mbar_init(&rbars[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) {
      ptx::tmem_alloc2(tmem_slot, C::kTmemCols);
      ptx::tmem_relinquish2();
    } else {
      ptx::tmem_alloc(tmem_slot, C::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync();
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long tiles_per_batch = (long long)a.tiles_m * a.tiles_n;

  if (warp == 0) {
    // ================= TMA producer =================
    if (ptx::elect_one()) {
      uint32_t it = 0;
      for (long long tile = tile_first; tile < a.total_tiles; tile += tile_step) {
        const int zs = (int)(tile / tiles_per_batch);
        const int r = (int)(tile - (long long)zs * tiles_per_batch);
        int tm, tn;
        tile_coords(a, r, tm, tn);
        const int m0 = tm * kTileM + (int)rank * kBM, n0 = tn * BN + (int)rank * (kPair ? kRowsB : 0);
        const int nbz = a.nb1 * a.nb2;
        const int sp = zs / nbz, z = zs - sp * nbz;
        const int z1 = z % a.nb1, z2 = z / a.nb1;
        int kb0, kb1;
        k_range(a, sp, tm, kTileM, kblocks_all, kBK, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % kNStages;
          const uint32_t ph = (it / kNStages) & 1;
          ptx::mbar_wait(&empty[s], ph ^ 1);
          const int az1 = a.a_b1 ? z1 : 0, az2 = a.a_b2 ? z2 : 0, bz1 = a.b_b1 ? z1 : 0, bz2 = a.b_b2 ? z2 : 0;
          if (kPair) {
            // both CTAs' bytes are counted on the leader's barrier
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * kStageBytes);
            const uint32_t bar = ptx::mapa(ptx::smem_u32(&full[s]), 0);
            if (kAMn) {
#pragma unroll
              for (int j = 0; j < kBM / kRowEl; ++j)
                ptx::tma_load_4d_pair(sA + s * C::kStageA + j * kBoxBytes, &tmA, bar, m0 + j * kRowEl, kb * kBK, az1, az2);
            } else {
              ptx::tma_load_4d_pair(sA + s * C::kStageA, &tmA, bar, kb * kBK, m0, az1, az2);
            }
            if (kBMn) {
#pragma unroll
              for (int j = 0; j < kRowsB / kRowEl; ++j)
                ptx::tma_load_4d_pair(sB + s * kStageBytesB + j * kBoxBytes, &tmB, bar, n0 + j * kRowEl, kb * kBK, bz1, bz2);
            } else {
              ptx::tma_load_4d_pair(sB + s * kStageBytesB, &tmB, bar, kb * kBK, n0, bz1, bz2);
            }
          } else {
            ptx::mbar_arrive_expect_tx(&full[s], C::kStage);
            if (kAMn) {
#pragma unroll
              for (int j = 0; j < kBM / kRowEl; ++j)
                ptx::tma_load_4d(sA + s * C::kStageA + j * kBoxBytes, &tmA, &full[s], m0 + j * kRowEl, kb * kBK, az1, az2);
            } else {
              ptx::tma_load_4d(sA + s * C::kStageA, &tmA, &full[s], kb * kBK, m0, az1, az2);
            }
            if (kBMn) {
#pragma unroll
              for (int j = 0; j < BN / kRowEl; ++j)
                ptx::tma_load_4d(sB + s * C::kStageB + j * kBoxBytes, &tmB, &full[s], n0 + j * kRowEl, kb * kBK, bz1, bz2);
            } else {
              ptx::tma_load_4d(sB + s * C::kStageB, &tmB, &full[s], kb * kBK, n0, bz1, bz2);
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer (pair mode: the leader CTA only) =================
    constexpr uint32_t idesc = ptx::umma_idesc(kTf32 ? 2 : 1, kTileM, BN, kAMn ? 1 : 0, kBMn ? 1 : 0);
    // per-instruction K step (16 bf16 / 8 tf32 = 32 bytes of K): K-major advances 32 B inside the swizzled row,
    // MN-major advances whole rows (16 or 8 rows of 128 B); in 16-byte descriptor units.
    constexpr uint32_t kStepA = kAMn ? (kTf32 ? 64u : 128u) : 2u;
    constexpr uint32_t kStepB = kBMn ? (kTf32 ? 64u : 128u) : 2u;
    uint32_t it = 0, ai = 0;
    for (long long tile = tile_first; tile < a.total_tiles; tile += tile_step, ++ai) {
      const int as = ai & 1;
      const uint32_t aph = (ai >> 1) & 1;
      const int sp = (int)(tile / (tiles_per_batch * a.nb1 * a.nb2));
      int kb0, kb1, tm = 0, tn = 0;
      if (a.band_hi > 0) tile_coords(a, (int)(tile % tiles_per_batch), tm, tn);
      k_range(a, sp, tm, kTileM, kblocks_all, kBK, kb0, kb1);
      ptx::mbar_wait(&tempty[as], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % kNStages;
        const uint32_t ph = (it / kNStages) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // tf32 MN-major: 32-byte-atom swizzle, 4-row K groups (512 B); everything else: SWIZZLE_128B, 8-row groups
          const uint64_t adesc = ptx::umma_desc_sw128(ptx::smem_u32(sA + s * C::kStageA), kAMn ? kBoxBytes : 16,
                                                      (kAMn && kTf32) ? 512 : 1024, (kAMn && kTf32) ? 1 : 2);
          const uint64_t bdesc = ptx::umma_desc_sw128(ptx::smem_u32(sB + s * kStageBytesB), kBMn ? kBoxBytes : 16,
                                                      (kBMn && kTf32) ? 512 : 1024, (kBMn && kTf32) ? 1 : 2);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
            if (kPair)
              ptx::mma_f16_pair(d_tmem, adesc + kStepA * k, bdesc + kStepB * k, idesc, acc);
            else if (kTf32)
              ptx::mma_tf32(d_tmem, adesc + kStepA * k, bdesc + kStepB * k, idesc, acc);
            else
              ptx::mma_f16(d_tmem, adesc + kStepA * k, bdesc + kStepB * k, idesc, acc);
          }
          if (kPair) {
            ptx::tc_commit2(&empty[s]);                     // both CTAs' producers may refill the stage
            if (kb == kb1 - 1) ptx::tc_commit2(&tfull[as]);  // both CTAs' epilogues may drain their half
          } else {
            ptx::tc_commit(&empty[s]);                    // frees the smem stage when these MMAs retire
            if (kb == kb1 - 1) ptx::tc_commit(&tfull[as]);  // accumulator complete
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which half of the tile's columns this warp drains
    constexpr int kHalfCols = BN / 2;
    constexpr int kChunks = kHalfCols / 32;
    unsigned char* boxes = smem + C::kStageOff + (warp - 4) * 4096;   // staged epilogue only
    // with a TMA-fed residual, box 0 receives the residual chunks and box 1 alone stages the stores (seq stays odd)
    const bool rtma = kStaged && a.res_tma;
    uint64_t* rbar = &rbars[warp - 4];
    uint32_t rph = 0;
    uint32_t seq = rtma ? 1u : 0u;
    uint32_t ai = 0;
    for (long long tile = tile_first; tile < a.total_tiles; tile += tile_step, ++ai) {
      const int zs = (int)(tile / tiles_per_batch);
      const int r = (int)(tile - (long long)zs * tiles_per_batch);
      int tm, tn;
      tile_coords(a, r, tm, tn);
      const int m0 = tm * kTileM + (int)rank * kBM, n0 = tn * BN;
      const int nbz = a.nb1 * a.nb2;
      const int sp = zs / nbz, z = zs - sp * nbz;
      const int z1 = z % a.nb1, z2 = z / a.nb1;
      const int as = ai & 1;
      const uint32_t aph = (ai >> 1) & 1;
      const int grow = m0 + q * 32 + lane;
      const long long c_row = (long long)z1 * a.C.s1 + (long long)z2 * a.C.s2 + (long long)sp * a.c_split + (long long)grow * a.C.ld;
      const long long x_row = (long long)z1 * a.aux.s1 + (long long)z2 * a.aux.s2 + (long long)grow * a.aux.ld;
      const long long r_row = (long long)z1 * a.res.s1 + (long long)z2 * a.res.s2 + (long long)grow * a.res.ld;
      const int col_base = n0 + half * kHalfCols;
      const bool tile_live = m0 + q * 32 < a.M;      // warp-uniform: this warp's 32 rows exist
      if (rtma) {
        if (tile_live && col_base < a.N && ptx::elect_one()) {
          ptx::fence_proxy_async();
          ptx::mbar_arrive_expect_tx(rbar, 2048);
          ptx::tma_load_4d(boxes, &tmR, rbar, col_base, m0 + q * 32, z1, z2);
        }
      } else if (kStaged) {
        prefetch_res(a, r_row, col_base, grow < a.M);
      }
      ptx::mbar_wait(&tfull[as], aph);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * kHalfCols;
      uint32_t va[32], vb[32];
      ptx::tmem_ld_32x32(t_row, va);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < kChunks) {
          if (c & 1) ptx::tmem_ld_32x32(t_row + 32 * (c + 1), va);
          else ptx::tmem_ld_32x32(t_row + 32 * (c + 1), vb);
          if (kStaged && !rtma) prefetch_res(a, r_row, col_base + 32 * (c + 1), grow < a.M);
        } else {
          // the whole half-tile is in registers: hand the accumulator buffer back before the global stores
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tempty[as]), 0));   // the leader's barrier counts both CTAs
            else ptx::mbar_arrive(&tempty[as]);
          }
        }
        if (kStaged) {
          float rt[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) rt[i] = 0.f;
          const bool chunk_live = tile_live && col_base + 32 * c < a.N;
          if (rtma && chunk_live) {
            ptx::mbar_wait(rbar, rph);
            rph ^= 1;
            read_box_row(boxes, lane, rt);
            __syncwarp();
            if (c + 1 < kChunks && col_base + 32 * (c + 1) < a.N && ptx::elect_one()) {   // next chunk's residual lands while this one is processed
              ptx::fence_proxy_async();
              ptx::mbar_arrive_expect_tx(rbar, 2048);
              ptx::tma_load_4d(boxes, &tmR, rbar, col_base + 32 * (c + 1), m0 + q * 32, z1, z2);
            }
          }
          const bool have_rt = rtma && chunk_live;
          if (c & 1) epilogue_chunk_staged(a, &tmC, &tmX, boxes, seq, lane, vb, m0 + q * 32, col_base + 32 * c, r_row, z1, z2, rt, have_rt);
          else epilogue_chunk_staged(a, &tmC, &tmX, boxes, seq, lane, va, m0 + q * 32, col_base + 32 * c, r_row, z1, z2, rt, have_rt);
          if (rtma) seq |= 1u;   // keep using box 1 for the stores
        } else {
          if (c & 1) epilogue_chunk(a, vb, grow, col_base + 32 * c, c_row, x_row, r_row);
          else epilogue_chunk(a, va, grow, col_base + 32 * c, c_row, x_row, r_row);
        }
      }
    }
    if (kStaged && ptx::elect_one()) ptx::bulk_wait_all();   // the staging boxes must outlive the last stores
  }

  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync();   // the peer's shared memory and barriers stay alive until both CTAs are done
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (kPair) ptx::tmem_dealloc2(tmem_base, C::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 4-D map (k, row, b1, b2) with a [box_k x box_rows x 1 x 1] SWIZZLE_128B box; out-of-range elements read as zero.
static int make_map(CUtensorMap* m, const T4sOperand& op, int K, int esize, bool tf32, int box_k, int box_rows, const char* name) {
  const bool mn = op.mn_major != 0;
  const long long inner = mn ? op.rows : K, outer = mn ? K : op.rows;  // inner = contiguous dimension
  ensure_context();
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return T4S_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) || (op.ld * esize) % 16 || op.ld < inner) {
    set_error("t4s_gemm: operand %s needs a 16-byte aligned base and pitch >= its contiguous extent (ld=%lld, extent=%lld)", name,
              (long long)op.ld, inner);
    return T4S_ERR_ARG;
  }
  const long long nb1 = std::max<long long>(1, op.nb1), nb2 = std::max<long long>(1, op.nb2);
  long long s1 = nb1 > 1 ? op.stride1 : op.ld, s2 = nb2 > 1 ? op.stride2 : op.ld;
  if ((s1 * esize) % 16 || (s2 * esize) % 16 || s1 <= 0 || s2 <= 0) {
    set_error("t4s_gemm: operand %s batch strides must be positive multiples of 16 bytes", name);
    return T4S_ERR_ARG;
  }
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)nb1, (cuuint64_t)nb2};
  cuuint64_t strides[3] = {(cuuint64_t)(op.ld * esize), (cuuint64_t)(s1 * esize), (cuuint64_t)(s2 * esize)};
  // K-major: box = [box_k of K (128 B)] x [box_rows rows]; MN-major: box = [128 B of rows] x [box_k K-rows]
  cuuint32_t box[4] = {(cuuint32_t)(mn ? 128 / esize : box_k), (cuuint32_t)(mn ? box_k : box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMapDataType dt = esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : (tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult rc = enc(m, dt, 4, const_cast<void*>(op.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    (mn && esize == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (K=%d rows=%lld ld=%lld nb=%lldx%lld)", name, (int)rc, K,
              (long long)op.rows, (long long)op.ld, nb1, nb2);
    return T4S_ERR_CUDA;
  }
  return T4S_OK;
}

static MatArg mat_arg(const T4sMatrix& m) {
  MatArg o;
  o.ptr = m.ptr;
  o.ld = m.ld;
  o.s1 = m.stride1;
  o.s2 = m.stride2;
  o.dtype = m.dtype;
  const int q = m.dtype == T4S_F32 ? 4 : 8;  // elements per 16 bytes
  o.vec = m.ptr && !(reinterpret_cast<uintptr_t>(m.ptr) % 16) && !(m.ld % q) && !(m.stride1 % q) && !(m.stride2 % q);
  return o;
}

// 4-D store map (col, row, b1, b2) over a bf16 output, box 32 x 32, SWIZZLE_64B.  Returns false when the matrix cannot be
// described (alignment / strides): the caller then uses the register-direct epilogue.
static bool make_store_map(CUtensorMap* m, const T4sMatrix& c, int M, int N, int nb1, int nb2) {
  if (!c.ptr || c.dtype != T4S_BF16 || (reinterpret_cast<uintptr_t>(c.ptr) & 15) || (c.ld * 2) % 16 || c.ld < N) return false;
  const long long s1 = nb1 > 1 ? c.stride1 : c.ld, s2 = nb2 > 1 ? c.stride2 : c.ld;
  if (s1 <= 0 || s2 <= 0 || (s1 * 2) % 16 || (s2 * 2) % 16) return false;
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  ensure_context();
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)nb1, (cuuint64_t)nb2};
  cuuint64_t strides[3] = {(cuuint64_t)(c.ld * 2), (cuuint64_t)(s1 * 2), (cuuint64_t)(s2 * 2)};
  cuuint32_t box[4] = {32, 32, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, c.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename Kern>
static int launch_kernel(Kern kern, int grid, int smem, bool pair, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const CUtensorMap& tmC, const CUtensorMap& tmX, const CUtensorMap& tmR, const Args& a) {
  T4S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (!pair) {
    kern<<<grid, kThreads, smem, st>>>(tmA, tmB, tmC, tmX, tmR, a);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    T4S_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmX, tmR, a));
  }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

template <int BN, bool kTf32, bool kAMn, bool kBMn>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap* tmC, const CUtensorMap* tmX, const CUtensorMap* tmR, Args& a,
                  cudaStream_t st, bool pair) {
  a.tiles_m = pair ? (a.M + 2 * kBM - 1) / (2 * kBM) : (a.M + kBM - 1) / kBM;
  a.tiles_n = (a.N + BN - 1) / BN;
  a.group_n = std::min(a.tiles_n, 16);
  const int bk = kTf32 ? 32 : 64;
  const int kblocks = (a.K + bk - 1) / bk;
  a.kb_per_split = (kblocks + a.split_k - 1) / a.split_k;
  if ((kblocks + a.kb_per_split - 1) / a.kb_per_split != a.split_k) {
    // an empty trailing split would leave its partial output unwritten: the caller must pick a split count that divides evenly
    set_error("t4s_gemm: split_k=%d leaves empty splits for %d k-blocks (use ceil(kblocks / ceil(kblocks / split_k)))", a.split_k, kblocks);
    return T4S_ERR_ARG;
  }
  a.total_tiles = (long long)a.tiles_m * a.tiles_n * a.nb1 * a.nb2 * a.split_k;
  const int grid = (int)std::min<long long>(a.total_tiles, sm_count());
  if constexpr (!kTf32 && BN == 256) {
    if (pair) {
      const int grid2 = 2 * (int)std::min<long long>(a.total_tiles, sm_count() / 2);   // one CTA pair per TPC
      if (tmC) return launch_kernel(gemm_kernel<BN, kTf32, kAMn, kBMn, true, true>, grid2, Cfg<BN>::kSmemStaged, true, st, tmA, tmB, *tmC,
                                    tmX ? *tmX : *tmC, tmR ? *tmR : *tmC, a);
      return launch_kernel(gemm_kernel<BN, kTf32, kAMn, kBMn, false, true>, grid2, Cfg<BN>::kSmem, true, st, tmA, tmB, tmA, tmA, tmA, a);
    }
  }
  if constexpr (!kTf32) {
    if (tmC) return launch_kernel(gemm_kernel<BN, kTf32, kAMn, kBMn, true, false>, grid, Cfg<BN>::kSmemStaged, false, st, tmA, tmB, *tmC,
                                  tmX ? *tmX : *tmC, tmR ? *tmR : *tmC, a);
  }
  return launch_kernel(gemm_kernel<BN, kTf32, kAMn, kBMn, false, false>, grid, Cfg<BN>::kSmem, false, st, tmA, tmB, tmA, tmA, tmA, a);
}

}  // namespace gemm
}  // namespace t4s

extern "C" int t4s_gemm(const T4sGemm* g, void* stream) {
  using namespace t4s::gemm;
  T4S_REQUIRE(g, "t4s_gemm: null descriptor");
  T4S_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0 && g->nb1 > 0 && g->nb2 > 0, "t4s_gemm: M, N, K and batch counts must be positive");
  T4S_REQUIRE(g->A.ptr && g->B.ptr && g->C.ptr, "t4s_gemm: A, B and C are required");
  T4S_REQUIRE(g->in_dtype == T4S_BF16 || g->in_dtype == T4S_F32, "t4s_gemm: in_dtype must be T4S_BF16 or T4S_F32");
  T4S_REQUIRE(g->A.rows == g->M && g->B.rows == g->N, "t4s_gemm: operand rows must equal M / N");
  for (const T4sMatrix* m : {&g->C, &g->aux, &g->residual})
    T4S_REQUIRE(!m->ptr || m->dtype == T4S_F32 || m->dtype == T4S_BF16, "t4s_gemm: bad output dtype");
  const bool tf32 = g->in_dtype == T4S_F32;
  const int esize = tf32 ? 4 : 2, bk = tf32 ? 32 : 64;
  const int BN = g->N > 128 ? 256 : (g->N > 64 ? 128 : 64);
  // CTA-pair (cta_group::2) tiles for bf16 GEMMs with full-width N tiles and at least one 256-row tile (T4S_GEMM_PAIR=0 disables)
  static const bool pair_enabled = [] { const char* e = getenv("T4S_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  // Measured on B200: the pair tile pays off when the contraction is deep (wgrad, fc2, the K >= 2304 dgrads: +4..10 %) or the output
  // narrow; wide-output / short-K products (fc1, qkv, the GELU' dgrad: N >= 2048 with K <= 1024) are epilogue-bound and run ~7 %
  // faster on single-CTA tiles (ncu: 291 us vs 315 us for M=76160 N=3072 K=768), so those keep them.
  const bool pair = pair_enabled && !tf32 && BN == 256 && g->M >= 256 && !(g->K <= 1024 && g->N >= 2048);
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, g->A, g->K, esize, tf32, bk, kBM, "A");
  if (rc) return rc;
  rc = make_map(&tmB, g->B, g->K, esize, tf32, bk, pair ? BN / 2 : BN, "B");
  if (rc) return rc;
  Args a;
  a.M = g->M; a.N = g->N; a.K = g->K; a.nb1 = g->nb1; a.nb2 = g->nb2;
  a.a_b1 = g->A.nb1 > 1; a.a_b2 = g->A.nb2 > 1; a.b_b1 = g->B.nb1 > 1; a.b_b2 = g->B.nb2 > 1;
  T4S_REQUIRE((!a.a_b1 || g->A.nb1 == g->nb1) && (!a.a_b2 || g->A.nb2 == g->nb2) && (!a.b_b1 || g->B.nb1 == g->nb1) &&
                  (!a.b_b2 || g->B.nb2 == g->nb2), "t4s_gemm: operand batch extents must be 1 or match nb1/nb2");
  a.C = mat_arg(g->C); a.aux = mat_arg(g->aux); a.res = mat_arg(g->residual);
  a.bias = g->bias; a.alpha = g->alpha; a.act = g->act;
  a.colsum = g->colsum;
  a.band_lo = g->band_hi > 0 ? g->band_lo : 0;
  a.band_hi = g->band_hi > 0 ? g->band_hi : 0;
  T4S_REQUIRE(g->band_hi <= 0 || g->band_hi > g->band_lo, "t4s_gemm: band_hi must exceed band_lo");
  a.bias_vec = g->bias && !(reinterpret_cast<uintptr_t>(g->bias) & 15);
  a.split_k = g->split_k > 1 ? g->split_k : 1;
  a.c_split = g->c_split_stride;
  T4S_REQUIRE(g->act != T4S_ACT_GELU_GRAD || g->residual.ptr, "t4s_gemm: T4S_ACT_GELU_GRAD needs the pre-activation in `residual`");
  T4S_REQUIRE(a.split_k == 1 || (g->c_split_stride > 0 && !g->bias && !g->residual.ptr && !g->aux.ptr && g->act == T4S_ACT_NONE),
              "t4s_gemm: split_k needs c_split_stride and a plain epilogue");
  cudaStream_t st = t4s::as_stream(stream);
  // bf16 outputs of un-split GEMMs leave through TMA tile stores (staged epilogue) whenever the layout allows it
  CUtensorMap tmCs, tmXs, tmRs;
  const CUtensorMap *pC = nullptr, *pX = nullptr, *pR = nullptr;
  a.res_tma = 0;
  if (!tf32 && a.split_k == 1 && g->M >= 32 && make_store_map(&tmCs, g->C, g->M, g->N, g->nb1, g->nb2) &&
      (!g->aux.ptr || make_store_map(&tmXs, g->aux, g->M, g->N, g->nb1, g->nb2)) &&
      !(g->aux.ptr && g->residual.ptr)) {
    pC = &tmCs;
    pX = g->aux.ptr ? &tmXs : nullptr;
    // the residual / pre-activation tile rides in by TMA too when it has the layout for it (a stride-0 broadcast does not)
    static const bool res_tma_enabled = [] { const char* e = getenv("T4S_GEMM_RES_TMA"); return !(e && e[0] == '0'); }();
    if (res_tma_enabled && g->residual.ptr && (g->nb1 == 1 || g->residual.stride1 > 0) && (g->nb2 == 1 || g->residual.stride2 > 0) &&
        make_store_map(&tmRs, g->residual, g->M, g->N, g->nb1, g->nb2)) {
      pR = &tmRs;
      a.res_tma = 1;
    }
  }
  T4S_REQUIRE(!g->colsum || (pC && g->nb1 * g->nb2 == 1), "t4s_gemm: colsum needs an un-batched, un-split bf16 output with a TMA-storable layout");
  const int variant = (tf32 ? 4 : 0) | (g->A.mn_major ? 2 : 0) | (g->B.mn_major ? 1 : 0);
#define T4S_GEMM_CASE(V, TF, AM, BM_)                                                \
  case V:                                                                            \
    if (BN == 256) return launch<256, TF, AM, BM_>(tmA, tmB, pC, pX, pR, a, st, pair);   \
    if (BN == 128) return launch<128, TF, AM, BM_>(tmA, tmB, pC, pX, pR, a, st, false);  \
    return launch<64, TF, AM, BM_>(tmA, tmB, pC, pX, pR, a, st, false);
  switch (variant) {
    T4S_GEMM_CASE(0, false, false, false)
    T4S_GEMM_CASE(1, false, false, true)
    T4S_GEMM_CASE(2, false, true, false)
    T4S_GEMM_CASE(3, false, true, true)
    T4S_GEMM_CASE(4, true, false, false)
    T4S_GEMM_CASE(5, true, false, true)
    T4S_GEMM_CASE(6, true, true, false)
    T4S_GEMM_CASE(7, true, true, true)
  }
#undef T4S_GEMM_CASE
  return T4S_ERR_ARG;
}

namespace t4s {
namespace gemm {
__global__ void reduce_splits_kernel(const float* __restrict__ ws, int splits, size_t n, float* __restrict__ out, int accumulate) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 acc = accumulate ? *reinterpret_cast<const float4*>(out + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < splits; ++s) {
        float4 v = *reinterpret_cast<const float4*>(ws + (size_t)s * n + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(out + i) = acc;
    } else {
      for (size_t j = i; j < n; ++j) {
        float acc = accumulate ? out[j] : 0.f;
        for (int s = 0; s < splits; ++s) acc += ws[(size_t)s * n + j];
        out[j] = acc;
      }
    }
  }
}
}  // namespace gemm
}  // namespace t4s

extern "C" int t4s_reduce_splits(const float* ws, int splits, size_t n, float* out, int accumulate, void* stream) {
  T4S_REQUIRE(ws && out && splits > 0, "t4s_reduce_splits: bad arguments");
  T4S_REQUIRE(n % 4 == 0 && !(reinterpret_cast<uintptr_t>(ws) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15),
              "t4s_reduce_splits: n must be a multiple of 4 and buffers 16-byte aligned");
  if (n == 0) return T4S_OK;
  const int grid = (int)std::min<size_t>((n / 4 + 255) / 256, (size_t)t4s::sm_count() * 8);
  t4s::gemm::reduce_splits_kernel<<<grid, 256, 0, t4s::as_stream(stream)>>>(ws, splits, n, out, accumulate);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
