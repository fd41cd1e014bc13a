// K2 and friends: row-wise kernels (one warp per row, values held in registers, fp32 statistics):
//   LayerNorm forward / backward (+ fused residual-gradient add), row softmax forward / backward,
//   relative-position softmax with the Transformer-XL shift folded into the read, column sums (bias / gamma grads),
//   GELU backward.  All take fp32 or bf16 activations; statistics and parameters are fp32.
#include <algorithm>
#include <initializer_list>

#include "common.cuh"
#include "ptx.cuh"

namespace t4s {
namespace rowops {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&u.x), hi = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// ------------------------------------------------------------------------------------------------------------
// LayerNorm.  Row r lives at x + (r / n_inner) * bstride + (r % n_inner) * cols  (a [B, skip:, C] slice is a view).
// ------------------------------------------------------------------------------------------------------------
constexpr int kLnMaxV = 8;  // float4 per lane -> cols <= 1024
// the common case (one contiguous [rows, cols] block) skips the 64-bit divisions of the sliced-view address
__device__ __forceinline__ long long ln_row_offset(long long r, long long rows, int cols, long long n_inner, long long bstride) {
  return n_inner >= rows ? r * cols : (r / n_inner) * bstride + (r % n_inner) * cols;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, T* __restrict__ y,
                                                          float* __restrict__ mean, float* __restrict__ rstd, long long rows,
                                                          int cols, float eps, float in_scale, long long n_inner,
                                                          long long bstride) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int nv = cols >> 2;
  for (long long r = warp_global; r < rows; r += nwarps) {
    const T* xr = x + ln_row_offset(r, rows, cols, n_inner, bstride);
    float4 v[kLnMaxV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        v[i] = Vec4<T>::load(xr + 4 * c);
        v[i].x *= in_scale; v[i].y *= in_scale; v[i].z *= in_scale; v[i].w *= in_scale;
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mu = warp_sum(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float a = v[i].x - mu, b = v[i].y - mu, c2 = v[i].z - mu, d = v[i].w - mu;
        q += (a * a + b * b) + (c2 * c2 + d * d);
      }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)cols + eps);
    T* yr = y + r * cols;
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c), b = *reinterpret_cast<const float4*>(beta + 4 * c);
        float4 o;
        o.x = (v[i].x - mu) * rs * g.x + b.x;
        o.y = (v[i].y - mu) * rs * g.y + b.y;
        o.z = (v[i].z - mu) * rs * g.z + b.z;
        o.w = (v[i].w - mu) * rs * g.w + b.w;
        Vec4<T>::store(yr + 4 * c, o);
      }
    }
    if (lane == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
  }
}

// dx = in_scale * rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) (+ dx_add);  partial dgamma/dbeta per block.
template <typename T>
__global__ void __launch_bounds__(kThreads) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                          const float* __restrict__ gamma, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, const T* __restrict__ dx_add,
                                                          T* __restrict__ dx, float* __restrict__ part /*[grid][2][cols]*/,
                                                          long long rows, int cols, float in_scale, long long n_inner,
                                                          long long bstride) {
  extern __shared__ float s_part[];  // [kWarpsPerBlock][2][cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + warp;
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int nv = cols >> 2;
  float4 ag[kLnMaxV], ab[kLnMaxV];
#pragma unroll
  for (int i = 0; i < kLnMaxV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = warp_global; r < rows; r += nwarps) {
    const long long xoff = ln_row_offset(r, rows, cols, n_inner, bstride);
    const T* xr = x + xoff;
    const T* dyr = dy + r * cols;
    const float mu = mean[r], rs = rstd[r];
    float4 xh[kLnMaxV], gd[kLnMaxV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float4 xv = Vec4<T>::load(xr + 4 * c), dv = Vec4<T>::load(dyr + 4 * c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        xh[i] = make_float4((xv.x * in_scale - mu) * rs, (xv.y * in_scale - mu) * rs, (xv.z * in_scale - mu) * rs,
                            (xv.w * in_scale - mu) * rs);
        gd[i] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
        s1 += (gd[i].x + gd[i].y) + (gd[i].z + gd[i].w);
        s2 += (gd[i].x * xh[i].x + gd[i].y * xh[i].y) + (gd[i].z * xh[i].z + gd[i].w * xh[i].w);
        ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
        ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
      }
    }
    const float m1 = warp_sum(s1) / (float)cols, m2 = warp_sum(s2) / (float)cols;
    const float k = rs * in_scale;
    T* dxr = dx + xoff;
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        float4 o;
        o.x = k * (gd[i].x - m1 - xh[i].x * m2);
        o.y = k * (gd[i].y - m1 - xh[i].y * m2);
        o.z = k * (gd[i].z - m1 - xh[i].z * m2);
        o.w = k * (gd[i].w - m1 - xh[i].w * m2);
        if (dx_add) {
          float4 a = Vec4<T>::load(dx_add + xoff + 4 * c);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        Vec4<T>::store(dxr + 4 * c, o);
      }
    }
  }
  if (part) {
    float* sp = s_part + warp * 2 * cols;
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        *reinterpret_cast<float4*>(sp + 4 * c) = ag[i];
        *reinterpret_cast<float4*>(sp + cols + 4 * c) = ab[i];
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 2 * cols; j += kThreads) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) acc += s_part[w * 2 * cols + j];
      part[(size_t)blockIdx.x * 2 * cols + j] = acc;
    }
  }
}


constexpr int kLnMaxC8 = 4;  // bf16 fast paths: 16-byte chunks per lane -> cols <= 1024

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
  return u;
}

// kDxSum (staged kernel below): additionally accumulates the column sums of the OUTPUT dx (third partial row): dx is the gradient of
// the residual stream, i.e. the output gradient of the linear layer (proj / fc2) that wrote that stream, so this is that layer's bias
// gradient.
// ------------------------------------------------------------------------------------------------------------
// bf16 LayerNorm with rows staged through shared memory by the bulk-copy engine (cols % 8 == 0, cols <= 1024, 16-byte aligned rows).
// A warp owns a ring of row slots; lane 0 issues `cp.async.bulk` copies a few rows ahead (completion on a per-slot mbarrier), so the
// memory latency of a row is hidden behind the arithmetic of the rows before it and does not occupy registers.  gamma (/ beta) stay
// in registers for the whole kernel.  Statistics and the arithmetic order per row are those of the register kernels above.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
constexpr int kFwdStages = 4, kBwdStages = 3;
static size_t ln_fwd_staged_smem(int cols) { return (size_t)kWarpsPerBlock * kFwdStages * (cols * 2 + 8); }
static size_t ln_bwd_staged_smem(int cols) { return (size_t)kWarpsPerBlock * kBwdStages * (3 * cols * 2 + 8); }

template <int kC8>
__global__ void __launch_bounds__(kThreads, 2)
ln_fwd_bf16_staged_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          __nv_bfloat16* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, long long rows, int cols, float eps,
                          float in_scale, long long n_inner, long long bstride) {
  extern __shared__ __align__(128) unsigned char ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + warp, nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int nc = cols >> 3, row_bytes = cols * 2;
  unsigned char* ring = ln_smem + warp * kFwdStages * row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + kWarpsPerBlock * kFwdStages * row_bytes) + warp * kFwdStages;
  if (lane == 0) {
    for (int st = 0; st < kFwdStages; ++st) ptx::mbar_init(&bars[st], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();
  auto issue = [&](long long r, int st) {
    ptx::mbar_arrive_expect_tx(&bars[st], row_bytes);
    bulk_g2s(ring + st * row_bytes, x + ln_row_offset(r, rows, cols, n_inner, bstride), row_bytes, &bars[st]);
  };
  if (lane == 0)
    for (int st = 0; st < kFwdStages; ++st)
      if (warp_global + st * nwarps < rows) issue(warp_global + st * nwarps, st);
  float g[kC8][8], b[kC8][8];
#pragma unroll
  for (int i = 0; i < kC8; ++i) {
    const int c = lane + 32 * i;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      g[i][e] = c < nc ? gamma[8 * c + e] : 0.f;
      b[i][e] = c < nc ? beta[8 * c + e] : 0.f;
    }
  }
  const float inv_cols = 1.0f / (float)cols;
  int it = 0;
  for (long long r = warp_global; r < rows; r += nwarps, ++it) {
    const int st = it % kFwdStages;
    ptx::mbar_wait(&bars[st], (it / kFwdStages) & 1);
    const uint4* xs = reinterpret_cast<const uint4*>(ring + st * row_bytes);
    float v[kC8][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kC8; ++i) {
      const int c = lane + 32 * i;
      if (c < nc) {
        unpack8(xs[c], v[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] *= in_scale;
        s += ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
      }
    }
    __syncwarp();                                     // every lane has its values: the slot can be refilled
    const long long rn = r + kFwdStages * nwarps;
    if (lane == 0 && rn < rows) issue(rn, st);
    const float mu = warp_sum(s) * inv_cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kC8; ++i) {
      if (lane + 32 * i < nc) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[i][e] - mu;
          q = fmaf(d, d, q);
        }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_cols + eps);
    uint4* yr = reinterpret_cast<uint4*>(y + r * cols);
#pragma unroll
    for (int i = 0; i < kC8; ++i) {
      const int c = lane + 32 * i;
      if (c < nc) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf((v[i][e] - mu) * rs, g[i][e], b[i][e]);
        yr[c] = pack8(o);
      }
    }
    if (lane == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
  }
}

template <int kC8, bool kDxSum>
__global__ void __launch_bounds__(kThreads, 2)
ln_bwd_bf16_staged_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dx_add,
                          __nv_bfloat16* __restrict__ dx, float* __restrict__ part /*[grid][2 or 3][cols]*/, long long rows, int cols,
                          float in_scale, long long n_inner, long long bstride) {
  extern __shared__ __align__(128) unsigned char ln_smem[];
  constexpr int kRows = kDxSum ? 3 : 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + warp, nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int nc = cols >> 3, row_bytes = cols * 2, slot_bytes = 3 * row_bytes;
  unsigned char* ring = ln_smem + warp * kBwdStages * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + kWarpsPerBlock * kBwdStages * slot_bytes) + warp * kBwdStages;
  if (lane == 0) {
    for (int st = 0; st < kBwdStages; ++st) ptx::mbar_init(&bars[st], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();
  auto issue = [&](long long r, int st) {      // slot = [x | dy | dx_add]
    const long long xoff = ln_row_offset(r, rows, cols, n_inner, bstride);
    unsigned char* slot = ring + st * slot_bytes;
    ptx::mbar_arrive_expect_tx(&bars[st], (dx_add ? 3 : 2) * row_bytes);
    bulk_g2s(slot, x + xoff, row_bytes, &bars[st]);
    bulk_g2s(slot + row_bytes, dy + r * cols, row_bytes, &bars[st]);
    if (dx_add) bulk_g2s(slot + 2 * row_bytes, dx_add + xoff, row_bytes, &bars[st]);
  };
  if (lane == 0)
    for (int st = 0; st < kBwdStages; ++st)
      if (warp_global + st * nwarps < rows) issue(warp_global + st * nwarps, st);
  // packed f32x2 arithmetic: pair j of a 16-byte chunk = elements (2j, 2j+1); the accumulators are pairs as well
  uint64_t g2[kC8][4], ag2[kC8][4], ab2[kC8][4], ad2[kDxSum ? kC8 : 1][4];
#pragma unroll
  for (int i = 0; i < kC8; ++i) {
    const int c = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      g2[i][j] = c < nc ? ptx::pack2(gamma[8 * c + 2 * j], gamma[8 * c + 2 * j + 1]) : ptx::pack2(0.f, 0.f);
      ag2[i][j] = ab2[i][j] = ptx::pack2(0.f, 0.f);
      if (kDxSum) ad2[i][j] = ptx::pack2(0.f, 0.f);
    }
  }
  auto pair_of = [](uint32_t w) { return ptx::pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); };   // two bf16 -> f32x2
  const float inv_cols = 1.0f / (float)cols;
  int it = 0;
  for (long long r = warp_global; r < rows; r += nwarps, ++it) {
    const int st = it % kBwdStages;
    const float mu = mean[r], rs = rstd[r];
    ptx::mbar_wait(&bars[st], (it / kBwdStages) & 1);
    const uint4* xs = reinterpret_cast<const uint4*>(ring + st * slot_bytes);
    const uint4* dys = xs + nc;
    const uint4* adds = dys + nc;
    // xhat = x * (in_scale * rstd) - mean * rstd
    const uint64_t xa = ptx::pack2(in_scale * rs, in_scale * rs), xb = ptx::pack2(-mu * rs, -mu * rs);
    uint64_t s1p = ptx::pack2(0.f, 0.f), s2p = ptx::pack2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kC8; ++i) {
      const int c = lane + 32 * i;
      if (c < nc) {
        const uint4 xw = xs[c], dw = dys[c];
        const uint32_t xv[4] = {xw.x, xw.y, xw.z, xw.w}, dv[4] = {dw.x, dw.y, dw.z, dw.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t xh = ptx::fma2(pair_of(xv[j]), xa, xb), d = pair_of(dv[j]), gd = ptx::mul2(d, g2[i][j]);
          s1p = ptx::add2(s1p, gd);
          s2p = ptx::fma2(gd, xh, s2p);
          ag2[i][j] = ptx::fma2(d, xh, ag2[i][j]);
          ab2[i][j] = ptx::add2(ab2[i][j], d);
        }
      }
    }
    float s1a, s1b, s2a, s2b;
    ptx::unpack2(s1p, s1a, s1b);
    ptx::unpack2(s2p, s2a, s2b);
    const float m1 = warp_sum(s1a + s1b) * inv_cols, m2 = warp_sum(s2a + s2b) * inv_cols;
    // dx = k (g dy - m1 - xhat m2) = (g dy) k - k m1 - xhat (k m2),  k = rstd * in_scale
    const float k = rs * in_scale;
    const uint64_t kk = ptx::pack2(k, k), nkm1 = ptx::pack2(-k * m1, -k * m1), nkm2 = ptx::pack2(-k * m2, -k * m2);
    uint4* dxr = reinterpret_cast<uint4*>(dx + ln_row_offset(r, rows, cols, n_inner, bstride));
#pragma unroll
    for (int i = 0; i < kC8; ++i) {
      const int c = lane + 32 * i;
      if (c < nc) {
        const uint4 xw = xs[c], dw = dys[c];
        const uint32_t xv[4] = {xw.x, xw.y, xw.z, xw.w}, dv[4] = {dw.x, dw.y, dw.z, dw.w};
        uint32_t av[4] = {0u, 0u, 0u, 0u};
        if (dx_add) {
          const uint4 aw = adds[c];
          av[0] = aw.x; av[1] = aw.y; av[2] = aw.z; av[3] = aw.w;
        }
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t xh = ptx::fma2(pair_of(xv[j]), xa, xb);
          uint64_t o = ptx::fma2(ptx::mul2(pair_of(dv[j]), g2[i][j]), kk, nkm1);
          o = ptx::fma2(xh, nkm2, o);
          if (dx_add) o = ptx::add2(o, pair_of(av[j]));
          if (kDxSum) ad2[i][j] = ptx::add2(ad2[i][j], o);
          float oa, ob;
          ptx::unpack2(o, oa, ob);
          __nv_bfloat162 t = __floats2bfloat162_rn(oa, ob);
          ow[j] = *reinterpret_cast<uint32_t*>(&t);
        }
        dxr[c] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
    __syncwarp();                                     // every lane is done with the slot
    const long long rn = r + kBwdStages * nwarps;
    if (lane == 0 && rn < rows) issue(rn, st);
  }
  float ag[kC8][8], ab[kC8][8], ad[kDxSum ? kC8 : 1][8];
#pragma unroll
  for (int i = 0; i < kC8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ptx::unpack2(ag2[i][j], ag[i][2 * j], ag[i][2 * j + 1]);
      ptx::unpack2(ab2[i][j], ab[i][2 * j], ab[i][2 * j + 1]);
      if (kDxSum) ptx::unpack2(ad2[i][j], ad[i][2 * j], ad[i][2 * j + 1]);
    }
  if (part) {
    __syncthreads();                                  // the rings are dead: their place takes the per-warp partial sums
    float* s_part = reinterpret_cast<float*>(ln_smem);
    float* sp = s_part + warp * kRows * cols;
#pragma unroll
    for (int i = 0; i < kC8; ++i) {
      const int c = lane + 32 * i;
      if (c < nc) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sp[8 * c + e] = ag[i][e];
          sp[cols + 8 * c + e] = ab[i][e];
          if (kDxSum) sp[2 * cols + 8 * c + e] = ad[i][e];
        }
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kRows * cols; j += kThreads) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) acc += s_part[w * kRows * cols + j];
      part[(size_t)blockIdx.x * kRows * cols + j] = acc;
    }
  }
}

// out0[j] (+)= sum_p part[p * stride + j] for j < n0;  out1[j - n0] likewise for n0 <= j < n0 + n1.
// Block = 32 columns x 8 partial-sum lanes (deterministic: fixed order inside a lane, fixed tree across lanes).
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float* __restrict__ part, int nparts, long long stride, int n0, int n1,
                                                           float* __restrict__ out0, float* __restrict__ out1, int accumulate) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (j < n0 + n1) {
    int p = ty;
    for (; p + 24 < nparts; p += 32) {
      const float a = part[(size_t)p * stride + j], b = part[(size_t)(p + 8) * stride + j], c = part[(size_t)(p + 16) * stride + j],
                  d = part[(size_t)(p + 24) * stride + j];
      acc += (a + b) + (c + d);
    }
    for (; p < nparts; p += 8) acc += part[(size_t)p * stride + j];
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < n0 + n1) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][tx];
    float* o = j < n0 ? out0 + j : out1 + (j - n0);
    *o = (accumulate ? *o : 0.f) + t;
  }
}

// Column sums, stage 1 (bf16 / fp32, 16-byte loads): block = 32 lanes x 8 warps; a lane owns 8 (bf16) or 4 (fp32) adjacent
// columns, the 8 warps walk interleaved rows of the block's strip with 4 loads in flight, then combine through shared memory.
template <typename T>
__global__ void __launch_bounds__(256) colsum_strip_kernel(const T* __restrict__ x, long long rows, int cols, long long ld,
                                                           float* __restrict__ part, int rows_per_block) {
  constexpr int kE = 16 / sizeof(T);
  __shared__ float sm[8][32 * kE + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * kE;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc[kE];
#pragma unroll
  for (int e = 0; e < kE; ++e) acc[e] = 0.f;
  if (c < cols) {
    auto add = [&](const uint4& u) {
      if (sizeof(T) == 2) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t = __bfloat1622float2(h[i]);
          acc[(2 * i) % kE] += t.x;
          acc[(2 * i + 1) % kE] += t.y;
        }
      } else {
        const float* f = reinterpret_cast<const float*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i % kE] += f[i];
      }
    };
    long long r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      const uint4 a = *reinterpret_cast<const uint4*>(x + r * ld + c), b = *reinterpret_cast<const uint4*>(x + (r + 8) * ld + c),
                  d = *reinterpret_cast<const uint4*>(x + (r + 16) * ld + c), e = *reinterpret_cast<const uint4*>(x + (r + 24) * ld + c);
      add(a); add(b); add(d); add(e);
    }
    for (; r < r1; r += 8) add(*reinterpret_cast<const uint4*>(x + r * ld + c));
  }
#pragma unroll
  for (int e = 0; e < kE; ++e) sm[warp][lane * kE + e] = acc[e];
  __syncthreads();
  for (int j = threadIdx.x; j < 32 * kE; j += 256) {
    const int col = blockIdx.x * 32 * kE + j;
    if (col < cols) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sm[w][j];
      part[(size_t)blockIdx.y * cols + col] = t;
    }
  }
}

// out[j] (+)= sum_p part[p * stride + j]
__global__ void reduce_rows_kernel(const float* __restrict__ part, int nparts, long long stride, int n, float* __restrict__ out,
                                   int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float acc = accumulate ? out[j] : 0.f;
  for (int p = 0; p < nparts; ++p) acc += part[(size_t)p * stride + j];
  out[j] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// Column sums: out[c] (+)= sum_r x[r, c].  Stage 1: block (256 threads x 4 columns) over a strip of rows.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, long long rows, int cols, long long ld,
                                                             float* __restrict__ part, int rows_per_block) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = r0; r < r1; ++r) {
    float4 v = Vec4<T>::load(x + r * ld + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(part + (size_t)blockIdx.y * cols + c) = acc;
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_scalar_kernel(const T* __restrict__ x, long long rows, int cols, long long ld,
                                                                    float* __restrict__ part, int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += to_f32<T>(x[r * ld + c]);
  part[(size_t)blockIdx.y * cols + c] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// GELU backward: dh = dy * gelu'(h)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ h, T* __restrict__ dh, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 d = Vec4<T>::load(dy + 4 * i), x = Vec4<T>::load(h + 4 * i), o;
    auto g = [](float v) { return 0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * 0.3989422804014327f * __expf(-0.5f * v * v); };
    o.x = d.x * g(x.x); o.y = d.y * g(x.y); o.z = d.z * g(x.z); o.w = d.w * g(x.w);
    Vec4<T>::store(dh + 4 * i, o);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Row softmax.  kRel: logits = ac[i, j] + bd[i, T-1-i+j]  (Transformer-XL rel_shift folded into the address).
// ------------------------------------------------------------------------------------------------------------
constexpr int kSmMaxE = 64;  // elements per lane -> cols <= 2048

template <typename T, bool kRel>
__global__ void __launch_bounds__(kThreads) softmax_fwd_kernel(const T* __restrict__ s, const T* __restrict__ bd, T* __restrict__ p,
                                                               long long rows, int cols, long long ld_s, long long ld_bd,
                                                               long long ld_p, int T_len) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp_global; r < rows; r += nwarps) {
    const T* sr = s + r * ld_s;
    const T* br = kRel ? bd + r * ld_bd + (T_len - 1 - (int)(r % T_len)) : nullptr;
    float v[kSmMaxE];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kSmMaxE; ++i) {
      const int c = lane + 32 * i;
      if (c < cols) {
        float a = to_f32<T>(sr[c]);
        if (kRel) a += to_f32<T>(br[c]);
        v[i] = a;
        mx = fmaxf(mx, a);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kSmMaxE; ++i) {
      const int c = lane + 32 * i;
      if (c < cols) {
        v[i] = __expf(v[i] - mx);
        sum += v[i];
      }
    }
    const float inv = 1.0f / warp_sum(sum);
    T* pr = p + r * ld_p;
#pragma unroll
    for (int i = 0; i < kSmMaxE; ++i) {
      const int c = lane + 32 * i;
      if (c < cols) pr[c] = from_f32<T>(v[i] * inv);
    }
  }
}

// ds = p * (dp - sum_j p*dp); written over dp.  kRel additionally scatters ds into the (zero-filled) shifted bd-gradient row.
template <typename T, bool kRel>
__global__ void __launch_bounds__(kThreads) softmax_bwd_kernel(const T* __restrict__ p, T* __restrict__ dp, T* __restrict__ dbd,
                                                               long long rows, int cols, long long ld_p, long long ld_dp,
                                                               long long ld_bd, int T_len) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp_global; r < rows; r += nwarps) {
    const T* pr = p + r * ld_p;
    T* dr = dp + r * ld_dp;
    float pv[kSmMaxE], dv[kSmMaxE];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kSmMaxE; ++i) {
      const int c = lane + 32 * i;
      if (c < cols) {
        pv[i] = to_f32<T>(pr[c]);
        dv[i] = to_f32<T>(dr[c]);
        dot += pv[i] * dv[i];
      }
    }
    dot = warp_sum(dot);
    const int shift = kRel ? T_len - 1 - (int)(r % T_len) : 0;
    T* br = kRel ? dbd + r * ld_bd : nullptr;
    if (kRel) {
      const int width = 2 * T_len - 1;
      for (int c = lane; c < shift; c += 32) br[c] = from_f32<T>(0.f);
      for (int c = shift + cols + lane; c < width; c += 32) br[c] = from_f32<T>(0.f);
    }
#pragma unroll
    for (int i = 0; i < kSmMaxE; ++i) {
      const int c = lane + 32 * i;
      if (c < cols) {
        const T o = from_f32<T>(pv[i] * (dv[i] - dot));
        dr[c] = o;
        if (kRel) br[shift + c] = o;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// bf16 row softmax with 16-byte accesses (cols % 8 == 0, cols <= 1024; one warp per row, 4 vectors of 8 per lane).
// The un-fused attention path (head sizes other than 64: the d = 384 PMAM / DASM decoders) spends its time here; the scalar
// kernels above move 2 bytes per thread per instruction.  kRel: the rel_shift is an unaligned 2-byte offset (T-1-i) into the
// position-score row, so that row goes through a per-warp shared-memory buffer: aligned 16-byte traffic to global memory,
// the odd offset is absorbed by 2-byte shared-memory accesses.
// ------------------------------------------------------------------------------------------------------------
constexpr int kVecMaxV = 4;                 // 8-element vectors per lane  -> cols <= 1024
constexpr int kRowBuf = 1024 + 16;          // bf16 elements per warp buffer

template <bool kRel>
__global__ void __launch_bounds__(kThreads) softmax_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ s, const __nv_bfloat16* __restrict__ bd,
                                                                    __nv_bfloat16* __restrict__ p, long long rows, int cols, long long ld_s,
                                                                    long long ld_bd, long long ld_p, int T_len) {
  __shared__ __align__(16) __nv_bfloat16 sbuf[kRel ? kWarpsPerBlock * kRowBuf : 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + warp;
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int nvec = cols >> 3;
  __nv_bfloat16* wb = sbuf + (kRel ? warp * kRowBuf : 0);
  for (long long r = warp_global; r < rows; r += nwarps) {
    const uint4* sr = reinterpret_cast<const uint4*>(s + r * ld_s);
    int lo = 0;
    if (kRel) {
      // stage bd[r, shift .. shift + cols) : aligned vectors [shift/8, (shift + cols + 7)/8)
      const int shift = T_len - 1 - (int)(r % T_len);
      const int v0 = shift >> 3, v1 = (shift + cols + 7) >> 3;
      lo = shift - 8 * v0;
      const uint4* br = reinterpret_cast<const uint4*>(bd + r * ld_bd);
      __syncwarp();
      for (int v = v0 + lane; v < v1; v += 32) reinterpret_cast<uint4*>(wb)[v - v0] = br[v];
      __syncwarp();
    }
    float x[kVecMaxV][8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kVecMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        unpack8(sr[v], x[i]);
        if (kRel) {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[i][e] += __bfloat162float(wb[lo + 8 * v + e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) mx = fmaxf(mx, x[i][e]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kVecMaxV; ++i) {
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          x[i][e] = __expf(x[i][e] - mx);
          sum += x[i][e];
        }
      }
    }
    const float inv = 1.0f / warp_sum(sum);
    uint4* pr = reinterpret_cast<uint4*>(p + r * ld_p);
#pragma unroll
    for (int i = 0; i < kVecMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[i][e] *= inv;
        pr[v] = pack8(x[i]);
      }
    }
  }
}

// ds = p * (dp - sum_j p dp), written over dp; kRel also writes the whole shifted position-gradient row dbd[r, 0 .. ld_bd)
// (ds at columns [shift, shift + cols), zeros elsewhere) with aligned 16-byte stores.
template <bool kRel>
__global__ void __launch_bounds__(kThreads) softmax_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ p, __nv_bfloat16* __restrict__ dp,
                                                                    __nv_bfloat16* __restrict__ dbd, long long rows, int cols, long long ld_p,
                                                                    long long ld_dp, long long ld_bd, int T_len) {
  __shared__ __align__(16) __nv_bfloat16 sbuf[kRel ? kWarpsPerBlock * kRowBuf : 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warp_global = (long long)blockIdx.x * kWarpsPerBlock + warp;
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  const int nvec = cols >> 3;
  __nv_bfloat16* wb = sbuf + (kRel ? warp * kRowBuf : 0);
  for (long long r = warp_global; r < rows; r += nwarps) {
    const uint4* pr = reinterpret_cast<const uint4*>(p + r * ld_p);
    uint4* dr = reinterpret_cast<uint4*>(dp + r * ld_dp);
    float pv[kVecMaxV][8], dv[kVecMaxV][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kVecMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        unpack8(pr[v], pv[i]);
        unpack8(dr[v], dv[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) dot += pv[i][e] * dv[i][e];
      }
    }
    dot = warp_sum(dot);
    if (kRel) __syncwarp();
#pragma unroll
    for (int i = 0; i < kVecMaxV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) pv[i][e] *= dv[i][e] - dot;
        const uint4 o = pack8(pv[i]);
        dr[v] = o;
        if (kRel) reinterpret_cast<uint4*>(wb)[v] = o;
      }
    }
    if (kRel) {
      __syncwarp();
      const int shift = T_len - 1 - (int)(r % T_len);
      uint4* br = reinterpret_cast<uint4*>(dbd + r * ld_bd);
      const int nout = (int)(ld_bd >> 3);
      for (int v = lane; v < nout; v += 32) {
        const int c0 = 8 * v - shift;            // source column of element 0
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (c0 > -8 && c0 < cols) {
          __align__(16) __nv_bfloat16 t[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c0 + e;
            t[e] = (c >= 0 && c < cols) ? wb[c] : __float2bfloat16_rn(0.f);
          }
          o = *reinterpret_cast<const uint4*>(t);
        }
        br[v] = o;
      }
    }
  }
}

static bool softmax_vec_ok(int dtype, int cols, std::initializer_list<const void*> ptrs, std::initializer_list<long long> lds) {
  if (dtype != T4S_BF16 || cols % 8 || cols > 8 * 32 * kVecMaxV) return false;
  for (const void* q : ptrs)
    if (q && (reinterpret_cast<uintptr_t>(q) & 15)) return false;
  for (long long l : lds)
    if (l % 8) return false;
  return true;
}

static int grid_for_rows(long long rows) {
  long long blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)sm_count() * 8));
}

}  // namespace rowops
}  // namespace t4s

using namespace t4s::rowops;

#define T4S_DISPATCH_DTYPE(dtype, ...)                                   \
  do {                                                                   \
    if ((dtype) == T4S_F32) { using T = float; __VA_ARGS__; }            \
    else if ((dtype) == T4S_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { t4s::set_error("bad dtype %d", (int)(dtype)); return T4S_ERR_ARG; } \
  } while (0)

extern "C" {

int t4s_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows,
                      int cols, float eps, float in_scale, int dtype, int64_t n_inner, int64_t x_bstride, void* stream) {
  T4S_REQUIRE(x && gamma && beta && y && rows > 0, "t4s_layernorm_fwd: bad arguments");
  T4S_REQUIRE(cols % 4 == 0 && cols > 0 && cols <= 128 * kLnMaxV, "t4s_layernorm_fwd: cols must be a multiple of 4 and <= %d", 128 * kLnMaxV);
  if (n_inner <= 0) { n_inner = rows; x_bstride = 0; }
  T4S_REQUIRE(x_bstride % 4 == 0, "t4s_layernorm_fwd: batch stride must be a multiple of 4 elements");
  const bool fast = dtype == T4S_BF16 && cols % 8 == 0 && cols <= 256 * kLnMaxC8 && x_bstride % 8 == 0 &&
                    !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15);
  if (fast) {
    using B16 = __nv_bfloat16;
    const size_t smem = ln_fwd_staged_smem(cols);
    const int grid = (int)std::max<long long>(1, std::min<long long>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock, 2L * t4s::sm_count()));
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024) T4S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, kThreads, smem, t4s::as_stream(stream)>>>(static_cast<const B16*>(x), gamma, beta, static_cast<B16*>(y), mean, rstd, rows, cols, eps,
                                                              in_scale, n_inner, x_bstride);
      return T4S_OK;
    };
    int rc;
    if (cols <= 256) rc = launch(ln_fwd_bf16_staged_kernel<1>);
    else if (cols <= 512) rc = launch(ln_fwd_bf16_staged_kernel<2>);
    else if (cols <= 768) rc = launch(ln_fwd_bf16_staged_kernel<3>);
    else rc = launch(ln_fwd_bf16_staged_kernel<4>);
    if (rc) return rc;
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (ln_fwd_kernel<T><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(x), gamma, beta, static_cast<T*>(y), mean, rstd, rows, cols, eps, in_scale,
                                n_inner, x_bstride)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

size_t t4s_layernorm_bwd_workspace(int64_t rows, int cols) {
  const int grid = (int)std::min<long long>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock, 2L * t4s::sm_count());
  return (size_t)std::max(grid, 1) * 3 * cols * sizeof(float);
}

int t4s_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dx_add,
                      void* dx, float* dgamma, float* dbeta, float* dx_colsum, float* ws, size_t ws_bytes, int64_t rows, int cols,
                      float in_scale, int dtype, int64_t n_inner, int64_t x_bstride, void* stream) {
  T4S_REQUIRE(dy && x && gamma && mean && rstd && dx && rows > 0, "t4s_layernorm_bwd: bad arguments");
  T4S_REQUIRE(cols % 4 == 0 && cols > 0 && cols <= 128 * kLnMaxV, "t4s_layernorm_bwd: cols must be a multiple of 4 and <= %d", 128 * kLnMaxV);
  if (n_inner <= 0) { n_inner = rows; x_bstride = 0; }
  const bool want_params = dgamma != nullptr || dbeta != nullptr || dx_colsum != nullptr;
  const int prows = dx_colsum ? 3 : 2;
  int grid = (int)std::max<long long>(1, std::min<long long>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock, 2L * t4s::sm_count()));
  if (want_params) T4S_REQUIRE(ws && ws_bytes >= (size_t)grid * prows * cols * sizeof(float), "t4s_layernorm_bwd: workspace too small");
  size_t smem = want_params ? (size_t)kWarpsPerBlock * prows * cols * sizeof(float) : 0;
  cudaStream_t st = t4s::as_stream(stream);
  const bool fast = dtype == T4S_BF16 && cols % 8 == 0 && cols <= 256 * kLnMaxC8 && x_bstride % 8 == 0 &&
                    !((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                       reinterpret_cast<uintptr_t>(dx_add) | reinterpret_cast<uintptr_t>(gamma)) & 15);
  if (fast) {
    using B16 = __nv_bfloat16;
    smem = ln_bwd_staged_smem(cols);     // the row rings; the partial sums reuse their place
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024) T4S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, kThreads, smem, st>>>(static_cast<const B16*>(dy), static_cast<const B16*>(x), gamma, mean, rstd,
                                         static_cast<const B16*>(dx_add), static_cast<B16*>(dx), want_params ? ws : nullptr, rows, cols,
                                         in_scale, n_inner, x_bstride);
      return T4S_OK;
    };
    int rc;
    if (dx_colsum) {
      if (cols <= 256) rc = launch(ln_bwd_bf16_staged_kernel<1, true>);
      else if (cols <= 512) rc = launch(ln_bwd_bf16_staged_kernel<2, true>);
      else if (cols <= 768) rc = launch(ln_bwd_bf16_staged_kernel<3, true>);
      else rc = launch(ln_bwd_bf16_staged_kernel<4, true>);
    } else {
      if (cols <= 256) rc = launch(ln_bwd_bf16_staged_kernel<1, false>);
      else if (cols <= 512) rc = launch(ln_bwd_bf16_staged_kernel<2, false>);
      else if (cols <= 768) rc = launch(ln_bwd_bf16_staged_kernel<3, false>);
      else rc = launch(ln_bwd_bf16_staged_kernel<4, false>);
    }
    if (rc) return rc;
  } else {
    T4S_REQUIRE(!dx_colsum, "t4s_layernorm_bwd: dx_colsum needs the bf16 fast path (16-byte aligned bf16 tensors, cols %% 8 == 0)");
    T4S_DISPATCH_DTYPE(dtype, {
      if (smem > 48 * 1024) T4S_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ln_bwd_kernel<T><<<grid, kThreads, smem, st>>>(static_cast<const T*>(dy), static_cast<const T*>(x), gamma, mean, rstd,
                                                     static_cast<const T*>(dx_add), static_cast<T*>(dx), want_params ? ws : nullptr,
                                                     rows, cols, in_scale, n_inner, x_bstride);
    });
  }
  T4S_LAUNCH_CHECK();
  if (want_params) {
    const long long pstride = (long long)prows * cols;
    if (dgamma && dbeta) {
      reduce_parts_kernel<<<(2 * cols + 31) / 32, 256, 0, st>>>(ws, grid, pstride, cols, cols, dgamma, dbeta, 0);
      T4S_LAUNCH_CHECK();
    } else if (dgamma) {
      reduce_parts_kernel<<<(cols + 31) / 32, 256, 0, st>>>(ws, grid, pstride, cols, 0, dgamma, nullptr, 0);
      T4S_LAUNCH_CHECK();
    } else if (dbeta) {
      reduce_parts_kernel<<<(cols + 31) / 32, 256, 0, st>>>(ws + cols, grid, pstride, cols, 0, dbeta, nullptr, 0);
      T4S_LAUNCH_CHECK();
    }
    if (dx_colsum) {
      reduce_parts_kernel<<<(cols + 31) / 32, 256, 0, st>>>(ws + 2 * cols, grid, pstride, cols, 0, dx_colsum, nullptr, 0);
      T4S_LAUNCH_CHECK();
    }
  }
  return T4S_OK;
}

static int colsum_parts(int64_t rows, int cols, int elems_per_lane) {
  // enough blocks for ~4 per SM, each strip at least 64 rows
  const long long col_blocks = (cols + 32 * elems_per_lane - 1) / (32 * elems_per_lane);
  const long long want = std::max<long long>(1, (4LL * t4s::sm_count() + col_blocks - 1) / col_blocks);
  return (int)std::max<long long>(1, std::min<long long>((rows + 63) / 64, want));
}

size_t t4s_colsum_workspace(int64_t rows, int cols) {
  // sized for the most finely split variant (fp32 lanes own 4 columns; the scalar fallback uses the same count)
  const int parts = std::max(colsum_parts(rows, cols, 4), colsum_parts(rows, cols, 8));
  return (size_t)parts * cols * sizeof(float);
}

int t4s_colsum(const void* x, int dtype, int64_t rows, int cols, int64_t ld, float* ws, size_t ws_bytes, float* out, int accumulate,
               void* stream) {
  T4S_REQUIRE(x && out && ws && rows > 0 && cols > 0, "t4s_colsum: bad arguments");
  T4S_REQUIRE(dtype == T4S_F32 || dtype == T4S_BF16, "t4s_colsum: bad dtype");
  const int kE = dtype == T4S_F32 ? 4 : 8;
  const bool vec = cols % kE == 0 && ld % kE == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int parts = colsum_parts(rows, cols, vec ? kE : 4);
  T4S_REQUIRE(ws_bytes >= (size_t)parts * cols * sizeof(float), "t4s_colsum: workspace too small");
  const int rpb = (int)((rows + parts - 1) / parts);
  const int nparts = (int)((rows + rpb - 1) / rpb);
  cudaStream_t st = t4s::as_stream(stream);
  if (vec) {
    dim3 grid((cols + 32 * kE - 1) / (32 * kE), nparts);
    T4S_DISPATCH_DTYPE(dtype, (colsum_strip_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x), rows, cols, ld, ws, rpb)));
  } else {
    dim3 grid((cols + 255) / 256, nparts);
    T4S_DISPATCH_DTYPE(dtype, (colsum_partial_scalar_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x), rows, cols, ld, ws, rpb)));
  }
  T4S_LAUNCH_CHECK();
  reduce_parts_kernel<<<(cols + 31) / 32, 256, 0, st>>>(ws, nparts, cols, cols, 0, out, nullptr, accumulate);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_gelu_bwd(const void* dy, const void* h, void* dh, size_t n, int dtype, void* stream) {
  T4S_REQUIRE(dy && h && dh && n % 4 == 0, "t4s_gelu_bwd: n must be a multiple of 4");
  if (n == 0) return T4S_OK;
  const int grid = (int)std::min<size_t>((n / 4 + 255) / 256, (size_t)t4s::sm_count() * 16);
  T4S_DISPATCH_DTYPE(dtype, (gelu_bwd_kernel<T><<<grid, 256, 0, t4s::as_stream(stream)>>>(static_cast<const T*>(dy), static_cast<const T*>(h),
                                                                                       static_cast<T*>(dh), n / 4)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_softmax_fwd(const void* s, void* p, int64_t rows, int cols, int64_t ld_s, int64_t ld_p, int dtype, void* stream) {
  T4S_REQUIRE(s && p && rows > 0 && cols > 0 && cols <= 32 * kSmMaxE, "t4s_softmax_fwd: cols must be in 1..%d", 32 * kSmMaxE);
  if (softmax_vec_ok(dtype, cols, {s, (const void*)p}, {(long long)ld_s, (long long)ld_p})) {
    softmax_fwd_bf16_kernel<false><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(s), nullptr, static_cast<__nv_bfloat16*>(p), rows, cols, ld_s, 0, ld_p, 0);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (softmax_fwd_kernel<T, false><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(s), nullptr, static_cast<T*>(p), rows, cols, ld_s, 0, ld_p, 0)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_softmax_bwd(const void* p, void* dp, int64_t rows, int cols, int64_t ld_p, int64_t ld_dp, int dtype, void* stream) {
  T4S_REQUIRE(p && dp && rows > 0 && cols > 0 && cols <= 32 * kSmMaxE, "t4s_softmax_bwd: cols must be in 1..%d", 32 * kSmMaxE);
  if (softmax_vec_ok(dtype, cols, {p, dp}, {(long long)ld_p, (long long)ld_dp})) {
    softmax_bwd_bf16_kernel<false><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(p), static_cast<__nv_bfloat16*>(dp), nullptr, rows, cols, ld_p, ld_dp, 0, 0);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (softmax_bwd_kernel<T, false><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(p), static_cast<T*>(dp), nullptr, rows, cols, ld_p, ld_dp, 0, 0)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_relpos_softmax_fwd(const void* ac, const void* bd, void* p, int64_t rows, int T_len, int64_t ld_ac, int64_t ld_bd,
                           int64_t ld_p, int dtype, void* stream) {
  T4S_REQUIRE(ac && bd && p && rows > 0 && T_len > 0 && T_len <= 32 * kSmMaxE && rows % T_len == 0 && ld_bd >= 2 * T_len - 1,
              "t4s_relpos_softmax_fwd: bad arguments");
  if (softmax_vec_ok(dtype, T_len, {ac, bd, p}, {(long long)ld_ac, (long long)ld_bd, (long long)ld_p}) && ld_bd >= ((2 * T_len - 1 + 7) & ~7)) {
    softmax_fwd_bf16_kernel<true><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(ac), static_cast<const __nv_bfloat16*>(bd), static_cast<__nv_bfloat16*>(p), rows, T_len, ld_ac, ld_bd, ld_p,
        T_len);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (softmax_fwd_kernel<T, true><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(ac), static_cast<const T*>(bd), static_cast<T*>(p), rows, T_len, ld_ac, ld_bd,
                                ld_p, T_len)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_relpos_softmax_bwd(const void* p, void* dp, void* dbd, int64_t rows, int T_len, int64_t ld_p, int64_t ld_dp, int64_t ld_bd,
                           int dtype, void* stream) {
  T4S_REQUIRE(p && dp && dbd && rows > 0 && T_len > 0 && T_len <= 32 * kSmMaxE && rows % T_len == 0 && ld_bd >= 2 * T_len - 1,
              "t4s_relpos_softmax_bwd: bad arguments");
  if (softmax_vec_ok(dtype, T_len, {p, dp, dbd}, {(long long)ld_p, (long long)ld_dp, (long long)ld_bd})) {
    softmax_bwd_bf16_kernel<true><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(p), static_cast<__nv_bfloat16*>(dp), static_cast<__nv_bfloat16*>(dbd), rows, T_len, ld_p, ld_dp, ld_bd,
        T_len);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (softmax_bwd_kernel<T, true><<<grid_for_rows(rows), kThreads, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(p), static_cast<T*>(dp), static_cast<T*>(dbd), rows, T_len, ld_p, ld_dp, ld_bd,
                                T_len)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
