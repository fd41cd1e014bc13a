// PMAM / DASM CNN branch (reference src/models/cnn/base.py:33-113, cnn_transformer/passt_cnn.py:52-62) in channels-last layout.
//
// Activations are [B, H, W, C] with C contiguous ("pixels x channels"), so
//   * Conv2d(3x3, pad 1, stride 1) = im2col3x3 (this file) + the tcgen05 GEMM  [pixels, 9 C_in] x [C_out, 9 C_in]^T (+bias),
//   * ContextGating's Linear over channels is a GEMM on the same [pixels, C] matrix with no permutes,
//   * BatchNorm2d statistics are column statistics of that matrix.
// Everything here is HBM-bound glue around those GEMMs (the branch is ~1 % of the model's FLOPs): one read + one write per
// element, fp32 arithmetic, deterministic two-stage reductions (fixed partition, ordered second stage).
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace cnn {

static int grid_for(long long n, int threads = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, (long long)sm_count() * 16));
}

// ---- im2col for a 3x3 / pad 1 / stride 1 convolution ---------------------------------------------------------------------
// col[(b, h, w), (ky*3 + kx) * C + c] = in[b, h+ky-1, w+kx-1, c]  (0 outside the image); columns [9C, Kp) are zero padding so that
// rows are 16-byte multiples for TMA.  The input is addressed through element strides, so the first layer reads the mel image
// [B, F, T] as a [B, T, F, 1] tensor without a transposed copy (passt_cnn.py:53).
template <typename TI, typename TO>
__global__ void im2col3x3_kernel(const TI* __restrict__ in, long long sb, long long sh, long long sw, TO* __restrict__ col, int B, int H, int W,
                                 int C, int Kp) {
  const long long total = (long long)B * H * W * 9;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(idx % 9);
    long long p = idx / 9;
    const int w = (int)(p % W);
    const long long q = p / W;
    const int h = (int)(q % H), b = (int)(q / H);
    const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
    TO* dst = col + p * Kp + tap * C;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      const TI* src = in + b * sb + hh * sh + ww * sw;
      for (int c = 0; c < C; ++c) dst[c] = from_f32<TO>(to_f32<TI>(src[c]));
    } else {
      for (int c = 0; c < C; ++c) dst[c] = from_f32<TO>(0.f);
    }
    if (tap == 8)
      for (int c = 9 * C; c < Kp; ++c) col[p * Kp + c] = from_f32<TO>(0.f);
  }
}

// 16-byte variant for contiguous channels-last inputs with C % 8 == 0 (bf16) / C % 4 == 0 (fp32): thread = (pixel, tap, vector)
template <typename T>
__global__ void im2col3x3_vec_kernel(const T* __restrict__ in, T* __restrict__ col, int B, int H, int W, int C, int Kp) {
  constexpr int kV = 16 / sizeof(T);
  const int cv = C / kV;
  const long long total = (long long)B * H * W * 9 * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * kV;
    long long t = idx / cv;
    const int tap = (int)(t % 9);
    const long long p = t / 9;
    const int w = (int)(p % W);
    const long long q = p / W;
    const int h = (int)(q % H);
    const long long b = q / H;
    const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = *reinterpret_cast<const uint4*>(in + ((b * H + hh) * W + ww) * C + c);
    *reinterpret_cast<uint4*>(col + p * Kp + tap * C + c) = v;
  }
}
template <typename T>
__global__ void col2im3x3_vec_kernel(const T* __restrict__ dcol, T* __restrict__ din, int B, int H, int W, int C, int Kp) {
  constexpr int kV = 16 / sizeof(T);
  const int cv = C / kV;
  const long long total = (long long)B * H * W * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * kV;
    long long p = idx / cv;
    const int w = (int)(p % W);
    const long long q = p / W;
    const int h = (int)(q % H);
    const long long b = q / H;
    float acc[kV];
#pragma unroll
    for (int e = 0; e < kV; ++e) acc[e] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h - (tap / 3 - 1), ww = w - (tap % 3 - 1);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const uint4 u = *reinterpret_cast<const uint4*>(dcol + ((b * H + hh) * W + ww) * Kp + tap * C + c);
        const T* tv = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (int e = 0; e < kV; ++e) acc[e] += to_f32<T>(tv[e]);
      }
    }
    __align__(16) T o[kV];
#pragma unroll
    for (int e = 0; e < kV; ++e) o[e] = from_f32<T>(acc[e]);
    *reinterpret_cast<uint4*>(din + p * C + c) = *reinterpret_cast<const uint4*>(o);
  }
}

// gradient of the above: din[b, h, w, c] = sum over taps of dcol[(b, h-ky+1, w-kx+1), tap, c]   (contiguous channels-last output)
template <typename T>
__global__ void col2im3x3_kernel(const T* __restrict__ dcol, T* __restrict__ din, int B, int H, int W, int C, int Kp) {
  const long long total = (long long)B * H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long p = idx / C;
    const int w = (int)(p % W);
    const long long q = p / W;
    const int h = (int)(q % H);
    const long long b = q / H;
    float acc = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h - (tap / 3 - 1), ww = w - (tap % 3 - 1);  // the output pixel whose tap `tap` reads (h, w)
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) acc += to_f32<T>(dcol[((b * H + hh) * W + ww) * Kp + tap * C + c]);
    }
    din[idx] = from_f32<T>(acc);
  }
}

// ---- per-channel statistics of X [rows, C]: part[blk][0][c] = sum x (* y), part[blk][1][c] = sum x^2 (or sum of the 2nd product) ----
// mode 0: (sum x, sum x^2)                       -> BatchNorm forward statistics
// mode 1: (sum dy, sum dy * (x - mean) * rstd)   -> BatchNorm backward reductions (= dbeta, dgamma)
template <typename T, int kMode>
__global__ void chan_stats_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, float* __restrict__ part, long long rows, int C, long long rows_per_block) {
  extern __shared__ float sm[];   // [2][blockDim.x]
  const int rpi = blockDim.x / C;  // rows per iteration
  const int c = threadIdx.x % C, rr = threadIdx.x / C;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (rr < rpi) {
    const float mu = kMode == 1 ? mean[c] : 0.f, rs = kMode == 1 ? rstd[c] : 0.f;
    for (long long r = r0 + rr; r < r1; r += rpi) {
      const float v = to_f32<T>(x[r * C + c]);
      if (kMode == 0) {
        s0 += v;
        s1 += v * v;
      } else {
        const float g = to_f32<T>(dy[r * C + c]);
        s0 += g;
        s1 += g * (v - mu) * rs;
      }
    }
  }
  sm[threadIdx.x] = s0;
  sm[blockDim.x + threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < rpi; ++k) {
      a += sm[k * C + threadIdx.x];
      b += sm[blockDim.x + k * C + threadIdx.x];
    }
    part[((long long)blockIdx.x * 2) * C + threadIdx.x] = a;
    part[((long long)blockIdx.x * 2 + 1) * C + threadIdx.x] = b;
  }
}

// BatchNorm2d training statistics (cnn/base.py:75: eps 1e-3, momentum 0.99): mean, rstd = 1/sqrt(biased var + eps) and the
// running-statistics update running = (1 - m) running + m batch (running_var with the unbiased batch variance, like torch).
__global__ void bn_finalize_kernel(const float* __restrict__ part, int blocks, int C, long long rows, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, ss = 0.0;
  for (int b = 0; b < blocks; ++b) {
    s += part[((long long)b * 2) * C + c];
    ss += part[((long long)b * 2 + 1) * C + c];
  }
  const double mu = s / (double)rows;
  const double var = fmax(ss / (double)rows - mu * mu, 0.0);
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(rows > 1 ? var * (double)rows / (double)(rows - 1) : var);
}
// eval mode: statistics from the running buffers
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, float eps, int C,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = running_mean[c];
  rstd[c] = rsqrtf(running_var[c] + eps);
}
__global__ void sum_parts2_kernel(const float* __restrict__ part, int blocks, int C, float* __restrict__ out0, float* __restrict__ out1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < blocks; ++k) {
    a += part[((long long)k * 2) * C + c];
    b += part[((long long)k * 2 + 1) * C + c];
  }
  out0[c] = a;
  if (out1) out1[c] = b;
}

// y = (x - mean) * rstd * gamma + beta
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y, long long n, int C) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    y[i] = from_f32<T>((to_f32<T>(x[i]) - mean[c]) * rstd[c] * gamma[c] + beta[c]);
  }
}
// training: dx = gamma rstd (dy - sum_dy / P - xhat * sum_dy_xhat / P);  eval (sums == NULL): dx = gamma rstd dy
template <typename T>
__global__ void bn_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                              const float* __restrict__ gamma, const float* __restrict__ sum_dy, const float* __restrict__ sum_dy_xhat,
                              T* __restrict__ dx, long long n, int C, float inv_rows) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float g = to_f32<T>(dy[i]);
    float v = g;
    if (sum_dy) {
      const float xh = (to_f32<T>(x[i]) - mean[c]) * rstd[c];
      v = g - sum_dy[c] * inv_rows - xh * sum_dy_xhat[c] * inv_rows;
    }
    dx[i] = from_f32<T>(gamma[c] * rstd[c] * v);
  }
}

// ---- ContextGating product (cnn/base.py:19-30): out = y * sigmoid(lin), optional inverted dropout (keep-scale 1/(1-p)) ----------
__device__ __forceinline__ uint32_t hash32(uint64_t k) {  // splitmix64 finaliser: counter-based, reproducible from (seed, index)
  k += 0x9E3779B97F4A7C15ull;
  k = (k ^ (k >> 30)) * 0xBF58476D1CE4E5B9ull;
  k = (k ^ (k >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((k ^ (k >> 31)) >> 32);
}
__device__ __forceinline__ float keep_scale(uint64_t seed, long long i, float p) {
  if (p <= 0.f) return 1.f;
  const float u = (float)(hash32(seed * 0x100000001B3ull + (uint64_t)i) >> 8) * (1.0f / 16777216.0f);
  return u >= p ? 1.0f / (1.0f - p) : 0.f;
}
template <typename T>
__global__ void gate_fwd_kernel(const T* __restrict__ y, const T* __restrict__ lin, T* __restrict__ out, long long n, float p, uint64_t seed) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = 1.0f / (1.0f + __expf(-to_f32<T>(lin[i])));
    out[i] = from_f32<T>(to_f32<T>(y[i]) * s * keep_scale(seed, i, p));
  }
}
template <typename T>
__global__ void gate_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ y, const T* __restrict__ lin, T* __restrict__ dy,
                                T* __restrict__ dlin, long long n, float p, uint64_t seed) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = 1.0f / (1.0f + __expf(-to_f32<T>(lin[i])));
    const float g = to_f32<T>(dout[i]) * keep_scale(seed, i, p), yv = to_f32<T>(y[i]);
    dy[i] = from_f32<T>(g * s);
    dlin[i] = from_f32<T>(g * yv * s * (1.0f - s));
  }
}

// plain inverted dropout (the DASM tagging decoder's nn.TransformerDecoderLayer dropouts); its own backward with the same seed
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ out, long long n, float p, uint64_t seed) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f32<T>(to_f32<T>(x[i]) * keep_scale(seed, i, p));
}

// ---- AvgPool2d((ph, pw)) on [B, H, W, C] -> [B, H/ph, W/pw, C] -----------------------------------------------------------------
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int ph, int pw) {
  const int Ho = H / ph, Wo = W / pw;
  const long long total = (long long)B * Ho * Wo * C;
  const float inv = 1.0f / (float)(ph * pw);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long p = idx / C;
    const int wo = (int)(p % Wo);
    const long long q = p / Wo;
    const int ho = (int)(q % Ho);
    const long long b = q / Ho;
    float acc = 0.f;
    for (int i = 0; i < ph; ++i)
      for (int j = 0; j < pw; ++j) acc += to_f32<T>(x[((b * H + ho * ph + i) * W + wo * pw + j) * C + c]);
    out[idx] = from_f32<T>(acc * inv);
  }
}
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dx, int B, int H, int W, int C, int ph, int pw) {
  const int Ho = H / ph, Wo = W / pw;
  const long long total = (long long)B * H * W * C;
  const float inv = 1.0f / (float)(ph * pw);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long p = idx / C;
    const int w = (int)(p % W);
    const long long q = p / W;
    const int h = (int)(q % H);
    const long long b = q / H;
    const int ho = h / ph, wo = w / pw;
    dx[idx] = (ho < Ho && wo < Wo) ? from_f32<T>(to_f32<T>(dout[((b * Ho + ho) * Wo + wo) * C + c]) * inv) : from_f32<T>(0.f);
  }
}

// ---- out = a + w * b with w a learnable scalar in device memory (merge_weight, passt_cnn.py:60-61) ---------------------------
template <typename T>
__global__ void scale_add_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ w, T* __restrict__ out,
                                     long long n) {
  const float wv = *w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f32<T>(to_f32<T>(a[i]) + wv * to_f32<T>(b[i]));
}
// db = w * dout (written when db != NULL);  part[blk] = sum over the block's fixed slice of dout * b  (dw, second stage below)
template <typename T>
__global__ void scale_add_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ b, const float* __restrict__ w, T* __restrict__ db,
                                     float* __restrict__ part, long long n, long long per_block) {
  __shared__ float s_red[32];
  const float wv = *w;
  const long long i0 = (long long)blockIdx.x * per_block, i1 = min(n, i0 + per_block);
  float acc = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const float g = to_f32<T>(dout[i]);
    acc += g * to_f32<T>(b[i]);
    if (db) db[i] = from_f32<T>(wv * g);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_red[k];
    part[blockIdx.x] = t;
  }
}
__global__ void sum_scalar_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < n; ++i) t += part[i];
    *out = t;
  }
}

// ---- prototype head (recipes/desed/pmam/train.py:82-87): z = x / max(||x||, 1e-12) row-wise ------------------------------------
template <typename T>
__global__ void l2norm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ inv_norm, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = to_f32<T>(x[r * C + c]); ss += v * v; }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < C; c += 32) y[r * C + c] = from_f32<T>(to_f32<T>(x[r * C + c]) * inv);
    if (lane == 0) inv_norm[r] = inv;
  }
}
// dx = inv (dy - y <dy, y>)
template <typename T>
__global__ void l2norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, const float* __restrict__ inv_norm, T* __restrict__ dx,
                                  long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot += to_f32<T>(dy[r * C + c]) * to_f32<T>(y[r * C + c]);
    dot = warp_sum(dot);
    const float inv = inv_norm[r];
    for (int c = lane; c < C; c += 32) dx[r * C + c] = from_f32<T>(inv * (to_f32<T>(dy[r * C + c]) - to_f32<T>(y[r * C + c]) * dot));
  }
}
// p = sigmoid((leaky_relu(s, slope) * 2 - 1) / temp)   on fp32 similarities; backward multiplies through
__global__ void proto_act_fwd_kernel(const float* __restrict__ s, float* __restrict__ p, long long n, float slope, float inv_temp) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = s[i], l = v >= 0.f ? v : slope * v;
    p[i] = 1.0f / (1.0f + __expf(-(2.f * l - 1.f) * inv_temp));
  }
}
__global__ void proto_act_bwd_kernel(const float* __restrict__ s, const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ ds,
                                     long long n, float slope, float inv_temp) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    ds[i] = dp[i] * pv * (1.f - pv) * inv_temp * 2.f * (s[i] >= 0.f ? 1.f : slope);
  }
}

}  // namespace cnn
}  // namespace t4s

using namespace t4s::cnn;

#define T4S_DISPATCH_DTYPE(dtype, ...)                                   \
  do {                                                                   \
    if ((dtype) == T4S_F32) { using T = float; __VA_ARGS__; }            \
    else if ((dtype) == T4S_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { t4s::set_error("bad dtype %d", (int)(dtype)); return T4S_ERR_ARG; } \
  } while (0)

static int stats_geometry(long long rows, int C, int* threads, int* blocks, long long* rows_per_block) {
  if (C <= 0 || C > 1024) return 0;
  const int rpi = std::max(1, 256 / C);
  *threads = ((rpi * C + 31) / 32) * 32;
  long long want = std::min<long long>((rows + 255) / 256, (long long)t4s::sm_count() * 8);
  want = std::max<long long>(1, std::min<long long>(want, 4096));
  *rows_per_block = (rows + want - 1) / want;
  *blocks = (int)((rows + *rows_per_block - 1) / *rows_per_block);
  return 1;
}

extern "C" {

int t4s_im2col3x3(const void* in, int in_dtype, int64_t stride_b, int64_t stride_h, int64_t stride_w, void* col, int col_dtype, int batch,
                  int height, int width, int channels, int k_padded, void* stream) {
  T4S_REQUIRE(in && col && batch > 0 && height > 0 && width > 0 && channels > 0 && k_padded >= 9 * channels, "t4s_im2col3x3: bad arguments");
  const long long total = (long long)batch * height * width * 9;
  cudaStream_t st = t4s::as_stream(stream);
  const int esz = in_dtype == T4S_BF16 ? 2 : 4, kv = 16 / esz;
  if (in_dtype == col_dtype && channels % kv == 0 && k_padded == 9 * channels && stride_w == channels && stride_h == (int64_t)width * channels &&
      stride_b == (int64_t)height * width * channels && !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(col)) & 15)) {
    const int vgrid = grid_for(total * (channels / kv));
    if (in_dtype == T4S_BF16)
      im2col3x3_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)col, batch, height, width, channels, k_padded);
    else
      im2col3x3_vec_kernel<float><<<vgrid, 256, 0, st>>>((const float*)in, (float*)col, batch, height, width, channels, k_padded);
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  const int grid = grid_for(total);
#define T4S_I2C(TI, TO) im2col3x3_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)in, stride_b, stride_h, stride_w, (TO*)col, batch, height, width, channels, k_padded)
  if (in_dtype == T4S_F32 && col_dtype == T4S_F32) T4S_I2C(float, float);
  else if (in_dtype == T4S_F32 && col_dtype == T4S_BF16) T4S_I2C(float, __nv_bfloat16);
  else if (in_dtype == T4S_BF16 && col_dtype == T4S_BF16) T4S_I2C(__nv_bfloat16, __nv_bfloat16);
  else { t4s::set_error("t4s_im2col3x3: bad dtypes"); return T4S_ERR_ARG; }
#undef T4S_I2C
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_col2im3x3(const void* dcol, void* din, int dtype, int batch, int height, int width, int channels, int k_padded, void* stream) {
  T4S_REQUIRE(dcol && din && batch > 0 && k_padded >= 9 * channels, "t4s_col2im3x3: bad arguments");
  const long long total = (long long)batch * height * width * channels;
  const int kv = dtype == T4S_BF16 ? 8 : 4;
  if (channels % kv == 0 && k_padded % kv == 0 && !((reinterpret_cast<uintptr_t>(dcol) | reinterpret_cast<uintptr_t>(din)) & 15)) {
    T4S_DISPATCH_DTYPE(dtype, (col2im3x3_vec_kernel<T><<<grid_for(total / kv), 256, 0, t4s::as_stream(stream)>>>(
                                  static_cast<const T*>(dcol), static_cast<T*>(din), batch, height, width, channels, k_padded)));
    T4S_LAUNCH_CHECK();
    return T4S_OK;
  }
  T4S_DISPATCH_DTYPE(dtype, (col2im3x3_kernel<T><<<grid_for(total), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(dcol), static_cast<T*>(din), batch, height, width, channels, k_padded)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

size_t t4s_chan_stats_workspace(int64_t rows, int channels) {
  int th, bl;
  long long rpb;
  if (!stats_geometry(rows, channels, &th, &bl, &rpb)) return 0;
  return (size_t)bl * 2 * channels * sizeof(float);
}

int t4s_batchnorm_fwd(const void* x, void* y, int dtype, int64_t rows, int channels, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float eps, float momentum, int training, float* mean, float* rstd, float* ws, size_t ws_bytes,
                      void* stream) {
  T4S_REQUIRE(x && y && gamma && beta && mean && rstd && rows > 0, "t4s_batchnorm_fwd: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  if (training) {
    int th, bl;
    long long rpb;
    T4S_REQUIRE(stats_geometry(rows, channels, &th, &bl, &rpb), "t4s_batchnorm_fwd: unsupported channel count %d", channels);
    T4S_REQUIRE(ws && ws_bytes >= (size_t)bl * 2 * channels * sizeof(float), "t4s_batchnorm_fwd: workspace too small");
    T4S_DISPATCH_DTYPE(dtype, (chan_stats_kernel<T, 0><<<bl, th, 2 * th * sizeof(float), st>>>(static_cast<const T*>(x), nullptr, nullptr, nullptr, ws,
                                                                                             rows, channels, rpb)));
    T4S_LAUNCH_CHECK();
    bn_finalize_kernel<<<(channels + 127) / 128, 128, 0, st>>>(ws, bl, channels, rows, eps, momentum, mean, rstd, running_mean, running_var);
    T4S_LAUNCH_CHECK();
  } else {
    T4S_REQUIRE(running_mean && running_var, "t4s_batchnorm_fwd: eval mode needs the running statistics");
    bn_eval_stats_kernel<<<(channels + 127) / 128, 128, 0, st>>>(running_mean, running_var, eps, channels, mean, rstd);
    T4S_LAUNCH_CHECK();
  }
  const long long n = rows * channels;
  T4S_DISPATCH_DTYPE(dtype, (bn_apply_kernel<T><<<grid_for(n), 256, 0, st>>>(static_cast<const T*>(x), mean, rstd, gamma, beta, static_cast<T*>(y), n, channels)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_batchnorm_bwd(const void* dy, const void* x, int dtype, int64_t rows, int channels, const float* gamma, const float* mean, const float* rstd,
                      int training, void* dx, float* dgamma, float* dbeta, float* ws, size_t ws_bytes, void* stream) {
  T4S_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && rows > 0, "t4s_batchnorm_bwd: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  int th, bl;
  long long rpb;
  T4S_REQUIRE(stats_geometry(rows, channels, &th, &bl, &rpb), "t4s_batchnorm_bwd: unsupported channel count %d", channels);
  T4S_REQUIRE(ws && ws_bytes >= (size_t)bl * 2 * channels * sizeof(float), "t4s_batchnorm_bwd: workspace too small");
  T4S_DISPATCH_DTYPE(dtype, (chan_stats_kernel<T, 1><<<bl, th, 2 * th * sizeof(float), st>>>(static_cast<const T*>(x), static_cast<const T*>(dy), mean,
                                                                                           rstd, ws, rows, channels, rpb)));
  T4S_LAUNCH_CHECK();
  sum_parts2_kernel<<<(channels + 127) / 128, 128, 0, st>>>(ws, bl, channels, dbeta, dgamma);
  T4S_LAUNCH_CHECK();
  const long long n = rows * channels;
  T4S_DISPATCH_DTYPE(dtype, (bn_bwd_kernel<T><<<grid_for(n), 256, 0, st>>>(static_cast<const T*>(dy), static_cast<const T*>(x), mean, rstd, gamma,
                                                                         training ? dbeta : nullptr, training ? dgamma : nullptr,
                                                                         static_cast<T*>(dx), n, channels, 1.0f / (float)rows)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_gate_fwd(const void* y, const void* lin, void* out, size_t n, float dropout_p, uint64_t seed, int dtype, void* stream) {
  T4S_REQUIRE(y && lin && out && dropout_p >= 0.f && dropout_p < 1.f, "t4s_gate_fwd: bad arguments");
  T4S_DISPATCH_DTYPE(dtype, (gate_fwd_kernel<T><<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(y), static_cast<const T*>(lin), static_cast<T*>(out), (long long)n, dropout_p, seed)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
int t4s_gate_bwd(const void* dout, const void* y, const void* lin, void* dy, void* dlin, size_t n, float dropout_p, uint64_t seed, int dtype,
                 void* stream) {
  T4S_REQUIRE(dout && y && lin && dy && dlin, "t4s_gate_bwd: bad arguments");
  T4S_DISPATCH_DTYPE(dtype, (gate_bwd_kernel<T><<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(dout), static_cast<const T*>(y), static_cast<const T*>(lin), static_cast<T*>(dy),
                                static_cast<T*>(dlin), (long long)n, dropout_p, seed)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_dropout(const void* x, void* out, size_t n, float dropout_p, uint64_t seed, int dtype, void* stream) {
  T4S_REQUIRE(x && out && dropout_p >= 0.f && dropout_p < 1.f, "t4s_dropout: bad arguments");
  T4S_DISPATCH_DTYPE(dtype, (dropout_kernel<T><<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(static_cast<const T*>(x), static_cast<T*>(out),
                                                                                                          (long long)n, dropout_p, seed)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_avgpool_fwd(const void* x, void* out, int dtype, int batch, int height, int width, int channels, int pool_h, int pool_w, void* stream) {
  T4S_REQUIRE(x && out && pool_h > 0 && pool_w > 0 && height >= pool_h && width >= pool_w, "t4s_avgpool_fwd: bad arguments");
  const long long total = (long long)batch * (height / pool_h) * (width / pool_w) * channels;
  T4S_DISPATCH_DTYPE(dtype, (avgpool_fwd_kernel<T><<<grid_for(total), 256, 0, t4s::as_stream(stream)>>>(static_cast<const T*>(x), static_cast<T*>(out), batch,
                                                                                                       height, width, channels, pool_h, pool_w)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
int t4s_avgpool_bwd(const void* dout, void* dx, int dtype, int batch, int height, int width, int channels, int pool_h, int pool_w, void* stream) {
  T4S_REQUIRE(dout && dx && pool_h > 0 && pool_w > 0, "t4s_avgpool_bwd: bad arguments");
  const long long total = (long long)batch * height * width * channels;
  T4S_DISPATCH_DTYPE(dtype, (avgpool_bwd_kernel<T><<<grid_for(total), 256, 0, t4s::as_stream(stream)>>>(static_cast<const T*>(dout), static_cast<T*>(dx), batch,
                                                                                                       height, width, channels, pool_h, pool_w)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_scale_add_fwd(const void* a, const void* b, const float* w, void* out, size_t n, int dtype, void* stream) {
  T4S_REQUIRE(a && b && w && out, "t4s_scale_add_fwd: bad arguments");
  T4S_DISPATCH_DTYPE(dtype, (scale_add_fwd_kernel<T><<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(a), static_cast<const T*>(b), w, static_cast<T*>(out), (long long)n)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
/* ws: T4S_SCALE_ADD_BLOCKS floats */
int t4s_scale_add_bwd(const void* dout, const void* b, const float* w, void* db, float* dw, float* ws, size_t n, int dtype, void* stream) {
  T4S_REQUIRE(dout && b && w && dw && ws, "t4s_scale_add_bwd: bad arguments");
  const int blocks = T4S_SCALE_ADD_BLOCKS;
  const long long per = ((long long)n + blocks - 1) / blocks;
  cudaStream_t st = t4s::as_stream(stream);
  T4S_DISPATCH_DTYPE(dtype, (scale_add_bwd_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(dout), static_cast<const T*>(b), w, static_cast<T*>(db), ws,
                                                                           (long long)n, per)));
  T4S_LAUNCH_CHECK();
  sum_scalar_kernel<<<1, 32, 0, st>>>(ws, blocks, dw);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_l2norm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int cols, int dtype, void* stream) {
  T4S_REQUIRE(x && y && inv_norm && rows > 0 && cols > 0, "t4s_l2norm_fwd: bad arguments");
  T4S_DISPATCH_DTYPE(dtype, (l2norm_fwd_kernel<T><<<grid_for(rows * 32), 256, 0, t4s::as_stream(stream)>>>(static_cast<const T*>(x), static_cast<T*>(y), inv_norm,
                                                                                                          rows, cols)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
int t4s_l2norm_bwd(const void* dy, const void* y, const float* inv_norm, void* dx, int64_t rows, int cols, int dtype, void* stream) {
  T4S_REQUIRE(dy && y && inv_norm && dx && rows > 0 && cols > 0, "t4s_l2norm_bwd: bad arguments");
  T4S_DISPATCH_DTYPE(dtype, (l2norm_bwd_kernel<T><<<grid_for(rows * 32), 256, 0, t4s::as_stream(stream)>>>(static_cast<const T*>(dy), static_cast<const T*>(y),
                                                                                                          inv_norm, static_cast<T*>(dx), rows, cols)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
int t4s_proto_act_fwd(const float* s, float* p, size_t n, float slope, float temperature, void* stream) {
  T4S_REQUIRE(s && p && temperature != 0.f, "t4s_proto_act_fwd: bad arguments");
  proto_act_fwd_kernel<<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(s, p, (long long)n, slope, 1.0f / temperature);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
int t4s_proto_act_bwd(const float* s, const float* p, const float* dp, float* ds, size_t n, float slope, float temperature, void* stream) {
  T4S_REQUIRE(s && p && dp && ds && temperature != 0.f, "t4s_proto_act_bwd: bad arguments");
  proto_act_bwd_kernel<<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(s, p, dp, ds, (long long)n, slope, 1.0f / temperature);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
