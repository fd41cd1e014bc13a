// Layout / glue kernels of the MAT-SED path (all HBM-bound, coalesced along the channel dimension):
//   patch im2col, cls/dist token rows, positional-embedding table, patch-embed small gradients,
//   frequency mean-pool, pad + linear x`ratio` interpolation (forward / backward), strided row-vector add,
//   dtype conversion, and the hi/lo tf32 split used by the error-compensated (3xTF32) parity mode.
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace misc {

template <typename T> __device__ __forceinline__ float ld(const T* p) { return to_f32<T>(*p); }

// ---- patch im2col: A[(b, f, t), i*P + j] = img[b, f*S + i, t*S + j]  (reference passt.py:302-315: Conv2d(1, D, 16, stride 10))
template <typename TI, typename TO>
__global__ void im2col_kernel(const TI* __restrict__ img, TO* __restrict__ out, int B, int H, int W, int F, int Tp, int P, int S) {
  const long long total = (long long)B * F * Tp * P;  // one thread per (b, f, t, i): copies P contiguous pixels
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % P);
    long long r = idx / P;
    const int t = (int)(r % Tp);
    r /= Tp;
    const int f = (int)(r % F);
    const int b = (int)(r / F);
    const TI* src = img + ((long long)b * H + f * S + i) * W + t * S;
    TO* dst = out + (((long long)b * F + f) * Tp + t) * (P * P) + i * P;
    for (int j = 0; j < P; ++j) dst[j] = from_f32<TO>(to_f32<TI>(src[j]));
  }
}

// ---- pos table: out[f*Tp + t, d] = time_pos[d, toff + t] + freq_pos[d, f]   (passt.py:503-519)
__global__ void posbias_kernel(const float* __restrict__ time_pos, const float* __restrict__ freq_pos, float* __restrict__ out, int D,
                               int F, int Tp, int Tp_table, int toff) {
  const long long total = (long long)F * Tp * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int p = (int)(idx / D);
    const int f = p / Tp, t = p % Tp;
    out[idx] = time_pos[(long long)d * Tp_table + toff + t] + freq_pos[(long long)d * F + f];
  }
}

// ---- x[b, 0, :] = cls + new_pos[0];  x[b, 1, :] = dist + new_pos[1]   (passt.py:560-569)
template <typename T>
__global__ void cls_dist_kernel(T* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ dist,
                                const float* __restrict__ new_pos, int B, long long bstride, int D) {
  const int total = B * 2 * D;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int d = idx % D, k = (idx / D) & 1, b = idx / (2 * D);
    x[(long long)b * bstride + (long long)k * D + d] = from_f32<T>((k ? dist[d] : cls[d]) + new_pos[k * D + d]);
  }
}

// ---- patch-embed small gradients, stage 1: tmp[p, d] = sum_b dx[b, p, d] for p in [0, n_tok)  (includes the 2 token rows)
template <typename T>
__global__ void batch_sum_kernel(const T* __restrict__ dx, float* __restrict__ tmp, int B, long long bstride, long long n) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += to_f32<T>(dx[(long long)b * bstride + idx]);
    tmp[idx] = acc;
  }
}
// stage 2: d_time[d, toff+t] = sum_f tmp[2+f*Tp+t, d]; d_freq[d, f] = sum_t ...; d_bias[d] = sum_p; d_cls/d_dist/d_newpos
__global__ void patch_small_grads_kernel(const float* __restrict__ tmp, float* __restrict__ d_time, float* __restrict__ d_freq,
                                         float* __restrict__ d_bias, float* __restrict__ d_cls, float* __restrict__ d_dist,
                                         float* __restrict__ d_newpos, int D, int F, int Tp, int Tp_table, int toff) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  if (d_cls) d_cls[d] = tmp[d];
  if (d_dist) d_dist[d] = tmp[D + d];
  if (d_newpos) {
    d_newpos[d] = tmp[d];
    d_newpos[D + d] = tmp[D + d];
  }
  float ball = 0.f;
  for (int f = 0; f < F; ++f) {
    float acc = 0.f;
    for (int t = 0; t < Tp; ++t) acc += tmp[(long long)(2 + f * Tp + t) * D + d];
    if (d_freq) d_freq[(long long)d * F + f] = acc;
    ball += acc;
  }
  if (d_bias) d_bias[d] = ball;
  if (d_time) {
    for (int t = 0; t < Tp_table; ++t) d_time[(long long)d * Tp_table + t] = 0.f;
    for (int t = 0; t < Tp; ++t) {
      float acc = 0.f;
      for (int f = 0; f < F; ++f) acc += tmp[(long long)(2 + f * Tp + t) * D + d];
      d_time[(long long)d * Tp_table + toff + t] = acc;
    }
  }
}

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
  return u;
}
// bf16 fast paths of the pooling / interpolation kernels below (C % 8 == 0, 16-byte aligned): a thread owns 8 adjacent channels of
// one output position; same fp32 arithmetic in the same order as the scalar kernels, 32-bit index math.
__global__ void __launch_bounds__(256) fpool_mean_fwd_bf16x8_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, int B, int F,
                                                                    int Tp, int c8n) {
  const int total = B * Tp * c8n;
  const float inv = 1.0f / (float)F;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / c8n, c = idx - r * c8n;
    const int b = r / Tp, t = r - b * Tp;
    const uint4* src = reinterpret_cast<const uint4*>(y) + ((long long)b * F * Tp + t) * c8n + c;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int f = 0; f < F; ++f) {
      float v[8];
      bf16x8_to_f32(src[(long long)f * Tp * c8n], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= inv;
    reinterpret_cast<uint4*>(out)[idx] = f32_to_bf16x8(acc);
  }
}
__global__ void __launch_bounds__(256) fpool_mean_bwd_bf16x8_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dy, int B, int F,
                                                                    int Tp, int c8n) {
  const int total = B * Tp * c8n;
  const float inv = 1.0f / (float)F;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / c8n, c = idx - r * c8n;
    const int b = r / Tp, t = r - b * Tp;
    float v[8];
    bf16x8_to_f32(reinterpret_cast<const uint4*>(dout)[idx], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= inv;
    const uint4 o = f32_to_bf16x8(v);
    uint4* dst = reinterpret_cast<uint4*>(dy) + ((long long)b * F * Tp + t) * c8n + c;
    for (int f = 0; f < F; ++f) dst[(long long)f * Tp * c8n] = o;
  }
}

// ---- frequency mean-pool: out[b, t, c] = mean_f y[b, f*Tp + t, c]   (passt_sed.py:206-208)
template <typename T>
__global__ void fpool_mean_fwd_kernel(const T* __restrict__ y, T* __restrict__ out, int B, int F, int Tp, int C) {
  const long long total = (long long)B * Tp * C;
  const float inv = 1.0f / (float)F;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long r = idx / C;
    const int t = (int)(r % Tp), b = (int)(r / Tp);
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc += to_f32<T>(y[((long long)b * F * Tp + (long long)f * Tp + t) * C + c]);
    out[idx] = from_f32<T>(acc * inv);
  }
}
template <typename T>
__global__ void fpool_mean_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dy, int B, int F, int Tp, int C) {
  const long long total = (long long)B * F * Tp * C;
  const float inv = 1.0f / (float)F;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int t = (int)(r % Tp);
    r /= Tp;
    const int b = (int)(r / F);
    dy[idx] = from_f32<T>(to_f32<T>(dout[((long long)b * Tp + t) * C + c]) * inv);
  }
}

// ---- pad (repeat last frame) + linear interpolation x ratio, align_corners=False  (passt_sed.py:258-259, 23-34)
__device__ __forceinline__ void interp_coeff(int o, int ratio, int L, int& i0, int& i1, float& w) {
  float src = ((float)o + 0.5f) / (float)ratio - 0.5f;
  src = fmaxf(src, 0.f);
  i0 = (int)src;
  i1 = min(i0 + 1, L - 1);
  w = src - (float)i0;
}
template <typename T>
__global__ void pad_interp_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int Tin, int ratio, int C, int pad) {
  const int L = Tin + pad, To = L * ratio;
  const long long total = (long long)B * To * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long r = idx / C;
    const int o = (int)(r % To), b = (int)(r / To);
    int i0, i1;
    float w;
    interp_coeff(o, ratio, L, i0, i1, w);
    const T* xb = x + (long long)b * Tin * C + c;
    const float a0 = to_f32<T>(xb[(long long)min(i0, Tin - 1) * C]), a1 = to_f32<T>(xb[(long long)min(i1, Tin - 1) * C]);
    out[idx] = from_f32<T>((1.0f - w) * a0 + w * a1);
  }
}
template <typename T>
__global__ void pad_interp_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dx, int B, int Tin, int ratio, int C, int pad) {
  const int L = Tin + pad, To = L * ratio;
  const long long total = (long long)B * Tin * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long r = idx / C;
    const int i = (int)(r % Tin), b = (int)(r / Tin);
    const T* db = dout + (long long)b * To * C + c;
    float acc = 0.f;
    // padded index j contributes to x[min(j, Tin-1)]; frame i gathers from j = i (and j = Tin when i == Tin-1)
    const int jhi = (pad && i == Tin - 1) ? Tin : i;
    for (int j = i; j <= jhi; ++j) {
      const int o_lo = max(0, (j - 1) * ratio), o_hi = min(To, (j + 2) * ratio);
      for (int o = o_lo; o < o_hi; ++o) {
        int i0, i1;
        float w;
        interp_coeff(o, ratio, L, i0, i1, w);
        float g = 0.f;
        if (i0 == j) g += 1.0f - w;
        if (i1 == j) g += w;
        if (g != 0.f) acc += g * to_f32<T>(db[(long long)o * C]);
      }
    }
    dx[idx] = from_f32<T>(acc);
  }
}

__global__ void __launch_bounds__(256) pad_interp_fwd_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int Tin,
                                                                    int ratio, int c8n, int pad) {
  const int L = Tin + pad, To = L * ratio, total = B * To * c8n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / c8n, c = idx - r * c8n;
    const int b = r / To, o = r - b * To;
    int i0, i1;
    float w;
    interp_coeff(o, ratio, L, i0, i1, w);
    const uint4* xb = reinterpret_cast<const uint4*>(x) + (long long)b * Tin * c8n + c;
    float a0[8], a1[8];
    bf16x8_to_f32(xb[(long long)min(i0, Tin - 1) * c8n], a0);
    bf16x8_to_f32(xb[(long long)min(i1, Tin - 1) * c8n], a1);
#pragma unroll
    for (int e = 0; e < 8; ++e) a0[e] = (1.0f - w) * a0[e] + w * a1[e];
    reinterpret_cast<uint4*>(out)[idx] = f32_to_bf16x8(a0);
  }
}
__global__ void __launch_bounds__(256) pad_interp_bwd_bf16x8_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dx, int B, int Tin,
                                                                    int ratio, int c8n, int pad) {
  const int L = Tin + pad, To = L * ratio, total = B * Tin * c8n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / c8n, c = idx - r * c8n;
    const int b = r / Tin, i = r - b * Tin;
    const uint4* db = reinterpret_cast<const uint4*>(dout) + (long long)b * To * c8n + c;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int jhi = (pad && i == Tin - 1) ? Tin : i;
    for (int j = i; j <= jhi; ++j) {
      const int o_lo = max(0, (j - 1) * ratio), o_hi = min(To, (j + 2) * ratio);
      for (int o = o_lo; o < o_hi; ++o) {
        int i0, i1;
        float w;
        interp_coeff(o, ratio, L, i0, i1, w);
        float g = 0.f;
        if (i0 == j) g += 1.0f - w;
        if (i1 == j) g += w;
        if (g != 0.f) {
          float v[8];
          bf16x8_to_f32(db[(long long)o * c8n], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += g * v[e];
        }
      }
    }
    reinterpret_cast<uint4*>(dx)[idx] = f32_to_bf16x8(acc);
  }
}

// ---- out[r, c] = scale * x[r*ld + c] + vec[c]
template <typename T>
__global__ void add_rowvec_kernel(const T* __restrict__ x, long long ld, const float* __restrict__ vec, T* __restrict__ out,
                                  long long rows, int cols, float scale) {
  const long long total = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    out[idx] = from_f32<T>(scale * to_f32<T>(x[r * ld + c]) + (vec ? vec[c] : 0.f));
  }
}

// bf16 fast path of the two strided element-wise kernels (cols % 8 == 0, pitches % 8 == 0, 16-byte aligned bases): a thread moves 8
// elements per 16-byte access and keeps two independent chunks in flight; 32-bit index arithmetic (rows * cols / 8 < 2^31).
// out[r*ldo + c] = alpha * x[r*ldx + c] + beta * y[r*ldy + c] + vec[c]   (y and / or vec may be NULL)
__global__ void __launch_bounds__(256) axpby_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ y,
                                                           long long ldy, const float* __restrict__ vec, __nv_bfloat16* __restrict__ out,
                                                           long long ldo, int total8, int c8n, float alpha, float beta) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += 2 * stride) {
    uint4 xv[2], yv[2];
    int row[2], c[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = i + u * stride;
      if (j < total8) {
        row[u] = (int)((unsigned)j / (unsigned)c8n);
        c[u] = 8 * (j - row[u] * c8n);
        xv[u] = *reinterpret_cast<const uint4*>(x + (long long)row[u] * ldx + c[u]);
        if (y) yv[u] = *reinterpret_cast<const uint4*>(y + (long long)row[u] * ldy + c[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (i + u * stride < total8) {
        float a[8], b[8];
        bf16x8_to_f32(xv[u], a);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] *= alpha;
        if (y) {
          bf16x8_to_f32(yv[u], b);
#pragma unroll
          for (int e = 0; e < 8; ++e) a[e] = fmaf(beta, b[e], a[e]);
        }
        if (vec) {
          const float4 v0 = *reinterpret_cast<const float4*>(vec + c[u]), v1 = *reinterpret_cast<const float4*>(vec + c[u] + 4);
          a[0] += v0.x; a[1] += v0.y; a[2] += v0.z; a[3] += v0.w; a[4] += v1.x; a[5] += v1.y; a[6] += v1.z; a[7] += v1.w;
        }
        *reinterpret_cast<uint4*>(out + (long long)row[u] * ldo + c[u]) = f32_to_bf16x8(a);
      }
    }
  }
}
static bool axpby_fast(const void* x, long long ldx, const void* y, long long ldy, const float* vec, void* out, long long ldo, long long rows, int cols,
                       int dtype) {
  if (dtype != T4S_BF16 || cols % 8 || ldx % 8 || ldo % 8 || (y && ldy % 8) || rows * (cols / 8) >= (1LL << 30)) return false;
  return !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(vec) | reinterpret_cast<uintptr_t>(out)) & 15);
}

// ---- out[r*ldo + c] = alpha * x[r*ldx + c] + beta * y[r*ldy + c]
template <typename T>
__global__ void add2_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ y, long long ldy, T* __restrict__ out, long long ldo,
                            long long rows, int cols, float alpha, float beta) {
  const long long total = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cols);
    const long long r = idx / cols;
    out[r * ldo + c] = from_f32<T>(alpha * to_f32<T>(x[r * ldx + c]) + beta * to_f32<T>(y[r * ldy + c]));
  }
}

template <typename TI, typename TO>
__global__ void convert_kernel(const TI* __restrict__ in, TO* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = from_f32<TO>(to_f32<TI>(in[i]));
}

// ---- 3xTF32 operand preparation: dst[z2][z1][row][3K] from a strided (optionally MN-major) fp32 operand.
// pattern 0 (A): [hi | lo | hi], pattern 1 (B): [hi | hi | lo]  =>  sum over 3K of A3*B3 = hi*hi + lo*hi + hi*lo.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int K, long long ld,
                                  int nb1, long long s1, int nb2, long long s2, int mn_major, int pattern, long long ld_dst) {
  const long long per_batch = rows * K;
  const long long total = per_batch * nb1 * nb2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long z = idx / per_batch, e = idx - z * per_batch;
    const int z1 = (int)(z % nb1), z2 = (int)(z / nb1);
    long long row, k;
    if (mn_major) { k = e / rows; row = e - k * rows; }   // consecutive threads walk the contiguous source dimension
    else          { row = e / K;  k = e - row * K; }
    const float v = src[(long long)z1 * s1 + (long long)z2 * s2 + (mn_major ? k * ld + row : row * ld + k)];
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    float* d = dst + (z * rows + row) * ld_dst + k;
    d[0] = hi;
    d[K] = pattern == 0 ? lo : hi;
    d[2LL * K] = pattern == 0 ? hi : lo;
  }
}

static int grid_for(long long n, int threads = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, (long long)sm_count() * 16));
}

}  // namespace misc
}  // namespace t4s

using namespace t4s::misc;

#define T4S_DISPATCH_DTYPE(dtype, ...)                                   \
  do {                                                                   \
    if ((dtype) == T4S_F32) { using T = float; __VA_ARGS__; }            \
    else if ((dtype) == T4S_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { t4s::set_error("bad dtype %d", (int)(dtype)); return T4S_ERR_ARG; } \
  } while (0)

extern "C" {

int t4s_patch_im2col(const void* img, int img_dtype, void* out, int out_dtype, int batch, int height, int width, int patch, int stride,
                     int f_dim, int t_dim, void* stream) {
  T4S_REQUIRE(img && out && batch > 0 && height >= patch && width >= patch && patch > 0 && stride > 0, "t4s_patch_im2col: bad arguments");
  const int F = f_dim, Tp = t_dim;
  T4S_REQUIRE(F > 0 && Tp > 0 && (F - 1) * stride + patch <= height && (Tp - 1) * stride + patch <= width,
              "t4s_patch_im2col: patch grid %dx%d does not fit a %dx%d image", F, Tp, height, width);
  const long long total = (long long)batch * F * Tp * patch;
  cudaStream_t st = t4s::as_stream(stream);
  const int grid = grid_for(total);
  if (img_dtype == T4S_F32 && out_dtype == T4S_F32) im2col_kernel<float, float><<<grid, 256, 0, st>>>((const float*)img, (float*)out, batch, height, width, F, Tp, patch, stride);
  else if (img_dtype == T4S_F32 && out_dtype == T4S_BF16) im2col_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)img, (__nv_bfloat16*)out, batch, height, width, F, Tp, patch, stride);
  else if (img_dtype == T4S_BF16 && out_dtype == T4S_BF16) im2col_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)img, (__nv_bfloat16*)out, batch, height, width, F, Tp, patch, stride);
  else if (img_dtype == T4S_BF16 && out_dtype == T4S_F32) im2col_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)img, (float*)out, batch, height, width, F, Tp, patch, stride);
  else { t4s::set_error("t4s_patch_im2col: bad dtypes"); return T4S_ERR_ARG; }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_patch_posbias(const float* time_pos, const float* freq_pos, float* out, int dim, int f_dim, int t_dim, int t_table, int t_offset,
                      void* stream) {
  T4S_REQUIRE(time_pos && freq_pos && out && t_offset >= 0 && t_offset + t_dim <= t_table, "t4s_patch_posbias: bad arguments");
  posbias_kernel<<<grid_for((long long)f_dim * t_dim * dim), 256, 0, t4s::as_stream(stream)>>>(time_pos, freq_pos, out, dim, f_dim, t_dim, t_table, t_offset);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_cls_dist_tokens(void* x, int dtype, const float* cls, const float* dist, const float* new_pos, int batch, int64_t batch_stride,
                        int dim, void* stream) {
  T4S_REQUIRE(x && cls && dist && new_pos, "t4s_cls_dist_tokens: null pointer");
  T4S_DISPATCH_DTYPE(dtype, (cls_dist_kernel<T><<<grid_for((long long)batch * 2 * dim), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<T*>(x), cls, dist, new_pos, batch, batch_stride, dim)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_patch_small_grads(const void* dx, int dtype, float* tmp, float* d_time, float* d_freq, float* d_bias, float* d_cls, float* d_dist,
                          float* d_newpos, int batch, int64_t batch_stride, int dim, int f_dim, int t_dim, int t_table, int t_offset,
                          void* stream) {
  T4S_REQUIRE(dx && tmp, "t4s_patch_small_grads: null pointer");
  const long long n = (long long)(2 + f_dim * t_dim) * dim;
  cudaStream_t st = t4s::as_stream(stream);
  T4S_DISPATCH_DTYPE(dtype, (batch_sum_kernel<T><<<grid_for(n), 256, 0, st>>>(static_cast<const T*>(dx), tmp, batch, batch_stride, n)));
  T4S_LAUNCH_CHECK();
  patch_small_grads_kernel<<<(dim + 127) / 128, 128, 0, st>>>(tmp, d_time, d_freq, d_bias, d_cls, d_dist, d_newpos, dim, f_dim, t_dim, t_table, t_offset);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_fpool_mean_fwd(const void* y, void* out, int dtype, int batch, int f_dim, int t_dim, int dim, void* stream) {
  T4S_REQUIRE(y && out, "t4s_fpool_mean_fwd: null pointer");
  if (dtype == T4S_BF16 && dim % 8 == 0 && !((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) && (long long)batch * f_dim * t_dim * (dim / 8) < (1LL << 30)) {
    fpool_mean_fwd_bf16x8_kernel<<<grid_for((long long)batch * t_dim * (dim / 8)), 256, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(out), batch, f_dim, t_dim, dim / 8);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (fpool_mean_fwd_kernel<T><<<grid_for((long long)batch * t_dim * dim), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(y), static_cast<T*>(out), batch, f_dim, t_dim, dim)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_fpool_mean_bwd(const void* dout, void* dy, int dtype, int batch, int f_dim, int t_dim, int dim, void* stream) {
  T4S_REQUIRE(dout && dy, "t4s_fpool_mean_bwd: null pointer");
  if (dtype == T4S_BF16 && dim % 8 == 0 && !((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dy)) & 15) && (long long)batch * f_dim * t_dim * (dim / 8) < (1LL << 30)) {
    fpool_mean_bwd_bf16x8_kernel<<<grid_for((long long)batch * t_dim * (dim / 8)), 256, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dy), batch, f_dim, t_dim, dim / 8);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (fpool_mean_bwd_kernel<T><<<grid_for((long long)batch * f_dim * t_dim * dim), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(dout), static_cast<T*>(dy), batch, f_dim, t_dim, dim)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_pad_interp_fwd(const void* x, void* out, int dtype, int batch, int t_in, int ratio, int dim, int pad, void* stream) {
  T4S_REQUIRE(x && out && t_in > 0 && ratio >= 1 && (pad == 0 || pad == 1), "t4s_pad_interp_fwd: bad arguments");
  if (dtype == T4S_BF16 && dim % 8 == 0 && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) && (long long)batch * (t_in + pad) * ratio * (dim / 8) < (1LL << 30)) {
    pad_interp_fwd_bf16x8_kernel<<<grid_for((long long)batch * (t_in + pad) * ratio * (dim / 8)), 256, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), batch, t_in, ratio, dim / 8, pad);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (pad_interp_fwd_kernel<T><<<grid_for((long long)batch * (t_in + pad) * ratio * dim), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(x), static_cast<T*>(out), batch, t_in, ratio, dim, pad)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_pad_interp_bwd(const void* dout, void* dx, int dtype, int batch, int t_in, int ratio, int dim, int pad, void* stream) {
  T4S_REQUIRE(dout && dx && t_in > 0 && ratio >= 1 && (pad == 0 || pad == 1), "t4s_pad_interp_bwd: bad arguments");
  if (dtype == T4S_BF16 && dim % 8 == 0 && !((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dx)) & 15) && (long long)batch * (t_in + pad) * ratio * (dim / 8) < (1LL << 30)) {
    pad_interp_bwd_bf16x8_kernel<<<grid_for((long long)batch * t_in * (dim / 8)), 256, 0, t4s::as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dx), batch, t_in, ratio, dim / 8, pad);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (pad_interp_bwd_kernel<T><<<grid_for((long long)batch * t_in * dim), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(dout), static_cast<T*>(dx), batch, t_in, ratio, dim, pad)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_add_rowvec(const void* x, int64_t ld, const float* vec, void* out, int64_t rows, int cols, float scale, int dtype, void* stream) {
  T4S_REQUIRE(x && out && rows > 0 && cols > 0, "t4s_add_rowvec: bad arguments");
  if (axpby_fast(x, ld, nullptr, 0, vec, out, cols, rows, cols, dtype)) {
    const int total8 = (int)(rows * (cols / 8));
    axpby_bf16x8_kernel<<<grid_for((total8 + 1) / 2), 256, 0, t4s::as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), ld, nullptr, 0, vec,
                                                                                      static_cast<__nv_bfloat16*>(out), cols, total8, cols / 8, scale, 0.f);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (add_rowvec_kernel<T><<<grid_for(rows * cols), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(x), ld, vec, static_cast<T*>(out), rows, cols, scale)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_add2(const void* x, int64_t ldx, const void* y, int64_t ldy, void* out, int64_t ldo, int64_t rows, int cols, float alpha, float beta,
             int dtype, void* stream) {
  T4S_REQUIRE(x && y && out && rows > 0 && cols > 0, "t4s_add2: bad arguments");
  if (axpby_fast(x, ldx, y, ldy, nullptr, out, ldo, rows, cols, dtype)) {
    const int total8 = (int)(rows * (cols / 8));
    axpby_bf16x8_kernel<<<grid_for((total8 + 1) / 2), 256, 0, t4s::as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                                                      static_cast<const __nv_bfloat16*>(y), ldy, nullptr,
                                                                                      static_cast<__nv_bfloat16*>(out), ldo, total8, cols / 8, alpha, beta);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (add2_kernel<T><<<grid_for(rows * cols), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(x), ldx, static_cast<const T*>(y), ldy, static_cast<T*>(out), ldo, rows, cols, alpha, beta)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_convert(const void* in, int in_dtype, void* out, int out_dtype, size_t n, void* stream) {
  T4S_REQUIRE(in && out, "t4s_convert: null pointer");
  if (n == 0) return T4S_OK;
  cudaStream_t st = t4s::as_stream(stream);
  const int grid = grid_for((long long)n);
  if (in_dtype == T4S_F32 && out_dtype == T4S_BF16) convert_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out, n);
  else if (in_dtype == T4S_BF16 && out_dtype == T4S_F32) convert_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (float*)out, n);
  else if (in_dtype == T4S_F32 && out_dtype == T4S_F32) convert_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, n);
  else if (in_dtype == T4S_BF16 && out_dtype == T4S_BF16) convert_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n);
  else { t4s::set_error("t4s_convert: bad dtypes"); return T4S_ERR_ARG; }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_split_tf32(const T4sOperand* src, int K, float* dst, int64_t ld_dst, int pattern, void* stream) {
  T4S_REQUIRE(src && src->ptr && dst && K > 0 && ld_dst >= 3LL * K && (pattern == 0 || pattern == 1), "t4s_split_tf32: bad arguments");
  const int nb1 = (int)std::max<int64_t>(1, src->nb1), nb2 = (int)std::max<int64_t>(1, src->nb2);
  const long long total = (long long)src->rows * K * nb1 * nb2;
  split_tf32_kernel<<<grid_for(total), 256, 0, t4s::as_stream(stream)>>>(static_cast<const float*>(src->ptr), dst, src->rows, K, src->ld, nb1,
                                                                        src->stride1, nb2, src->stride2, src->mn_major, pattern, ld_dst);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
