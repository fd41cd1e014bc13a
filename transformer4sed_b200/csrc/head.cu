// K6/K7: heads and losses of the SED path (fp32 statistics, deterministic two-stage reductions):
//   sed_pool   : sigmoid(logit / temp), padded frames -> 0, linear-softmax pooling   (reference passt_sed.py:285-296)
//   attnpool   : single learned-query multi-head attention pooling                    (reference pooling.py:37-51)
//   bce / mse  : torch.nn.BCELoss / MSELoss(mean) with optional row mask              (finetune/train.py:166-178,
//                                                                                      mlm/mlm_passt/train.py:36-38)
//   sigmoid    : audio-tagging head                                                    (passt_sed.py:236-240)
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace head {

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.f;
  if (warp == 0) t = warp_sum(t);
  if (threadIdx.x == 0) s_red[0] = t;
  __syncthreads();
  return s_red[0];
}

// ---- sed_pool: logits [B, T, K] fp32 -> strong [B, K, T], weak [B, K].  One block per (b, k).
__global__ void sed_pool_fwd_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ pad_mask, float inv_temp,
                                    float* __restrict__ strong, float* __restrict__ weak, int T, int K) {
  __shared__ float s_red[32];
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  float s1 = 0.f, s2 = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float p = 1.0f / (1.0f + expf(-logits[((long long)b * T + t) * K + k] * inv_temp));
    if (pad_mask && pad_mask[(long long)b * T + t]) p = 0.f;
    strong[((long long)b * K + k) * T + t] = p;
    s1 += p;
    s2 += p * p;
  }
  s1 = block_sum(s1, s_red);
  s2 = block_sum(s2, s_red);
  if (threadIdx.x == 0) weak[blockIdx.x] = fminf(fmaxf(s2 / s1, 1e-7f), 1.0f);
}

// dlogit[b,t,k] = (dstrong + dweak * d(weak)/dp) * p (1 - p) / temp
__global__ void sed_pool_bwd_kernel(const float* __restrict__ strong, const float* __restrict__ dstrong, const float* __restrict__ dweak,
                                    const unsigned char* __restrict__ pad_mask, float inv_temp, float* __restrict__ dlogits, int T, int K) {
  __shared__ float s_red[32];
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  const float* p_row = strong + ((long long)b * K + k) * T;
  float s1 = 0.f, s2 = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float p = p_row[t];
    s1 += p;
    s2 += p * p;
  }
  s1 = block_sum(s1, s_red);
  s2 = block_sum(s2, s_red);
  const float w = s2 / s1;
  float gw = dweak ? dweak[blockIdx.x] : 0.f;
  if (!(w >= 1e-7f && w <= 1.0f)) gw = 0.f;  // clamp passes gradient only inside [min, max]
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float p = p_row[t];
    float dp = dstrong ? dstrong[((long long)b * K + k) * T + t] : 0.f;
    dp += gw * (2.0f * p * s1 - s2) / (s1 * s1);
    float dl = dp * p * (1.0f - p) * inv_temp;
    if (pad_mask && pad_mask[(long long)b * T + t]) dl = 0.f;
    dlogits[((long long)b * T + t) * K + k] = dl;
  }
}

// ---- elementwise sigmoid
__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = 1.0f / (1.0f + expf(-x[i]));
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dx[i] = dy[i] * y[i] * (1.0f - y[i]);
}

// ---- BCE (torch semantics: log clamped at -100; backward divides by max(p(1-p), 1e-12))
__global__ void bce_partial_kernel(const float* __restrict__ p, const float* __restrict__ y, size_t n, float* __restrict__ part) {
  __shared__ float s_red[32];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i], yi = y[i];
    acc -= yi * fmaxf(logf(pi), -100.f) + (1.0f - yi) * fmaxf(logf(1.0f - pi), -100.f);
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void finish_mean_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ count_part, float denom,
                                   float* __restrict__ out) {
  // single warp: out[0] = sum(part) / (count_part ? sum(count_part) * denom : denom); out[1] = that divisor
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) {
    a += part[i];
    if (count_part) c += count_part[i];
  }
  a = warp_sum(a);
  c = warp_sum(c);
  if (threadIdx.x == 0) {
    const float div = count_part ? c * denom : denom;
    out[0] = a / div;
    out[1] = div;
  }
}
__global__ void bce_bwd_kernel(const float* __restrict__ p, const float* __restrict__ y, const float* __restrict__ gout, float inv_n,
                               float* __restrict__ dp, size_t n) {
  const float g = gout[0] * inv_n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    dp[i] = g * (pi - y[i]) / fmaxf(pi * (1.0f - pi), 1e-12f);
  }
}

// ---- (masked) MSE over rows of width C: mean over selected rows x C of (a - b)^2
template <typename T>
__global__ void mse_partial_kernel(const T* __restrict__ a, const T* __restrict__ b, const unsigned char* __restrict__ mask, long long rows,
                                   int C, float* __restrict__ part, float* __restrict__ count_part) {
  __shared__ float s_red[32];
  float acc = 0.f, cnt = 0.f;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    if (mask && !mask[r]) continue;
    if (threadIdx.x == 0) cnt += 1.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float d = to_f32<T>(a[r * C + c]) - to_f32<T>(b[r * C + c]);
      acc += d * d;
    }
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = acc;
    count_part[blockIdx.x] = cnt;
  }
}
template <typename T>
__global__ void mse_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const unsigned char* __restrict__ mask, long long rows, int C,
                               const float* __restrict__ gout, const float* __restrict__ fwd_out /*[1] = divisor*/, T* __restrict__ da,
                               T* __restrict__ db) {
  const float g = 2.0f * gout[0] / fwd_out[1];
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    float d = 0.f;
    if (!mask || mask[r]) d = g * (to_f32<T>(a[i]) - to_f32<T>(b[i]));
    if (da) da[i] = from_f32<T>(d);
    if (db) db[i] = from_f32<T>(-d);
  }
}

// ---- attention pooling with one query.  kv [Bp, K, 2C] (k then v per key row), q [C] fp32 already scaled by hd^-1/2.
// One block per pooled item; dynamic smem: probs [H][K] + q [C] (+ dctx [C] in backward).
template <typename T>
__global__ void __launch_bounds__(256) attnpool_fwd_kernel(const T* __restrict__ kv, const float* __restrict__ q, T* __restrict__ ctx,
                                                           float* __restrict__ probs, int K, int C, int H, long long item_stride) {
  extern __shared__ float sm[];
  float* s_p = sm;           // [H][K]
  float* s_q = sm + H * K;   // [C]
  const int bp = blockIdx.x, hd = C / H;
  const T* kvb = kv + (long long)bp * item_stride;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_q[c] = q[c];
  __syncthreads();
  for (int i = threadIdx.x; i < H * K; i += blockDim.x) {
    const int h = i / K, j = i % K;
    const T* kr = kvb + (long long)j * 2 * C + h * hd;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc += s_q[h * hd + d] * to_f32<T>(kr[d]);
    s_p[i] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < H; h += (blockDim.x >> 5)) {
    float mx = -INFINITY;
    for (int j = lane; j < K; j += 32) mx = fmaxf(mx, s_p[h * K + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < K; j += 32) {
      const float e = __expf(s_p[h * K + j] - mx);
      s_p[h * K + j] = e;
      sum += e;
    }
    const float inv = 1.0f / warp_sum(sum);
    for (int j = lane; j < K; j += 32) {
      const float p = s_p[h * K + j] * inv;
      s_p[h * K + j] = p;
      if (probs) probs[((long long)bp * H + h) * K + j] = p;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int h = c / hd;
    float acc = 0.f;
    for (int j = 0; j < K; ++j) acc += s_p[h * K + j] * to_f32<T>(kvb[(long long)j * 2 * C + C + c]);
    ctx[(long long)bp * C + c] = from_f32<T>(acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attnpool_bwd_kernel(const T* __restrict__ kv, const float* __restrict__ q, const float* __restrict__ probs,
                                                           const T* __restrict__ dctx, T* __restrict__ dkv, float* __restrict__ dq_part,
                                                           int K, int C, int H, long long item_stride) {
  extern __shared__ float sm[];
  float* s_ds = sm;              // [H][K]: dp then ds
  float* s_q = sm + H * K;       // [C]
  float* s_dc = s_q + C;         // [C]
  const int bp = blockIdx.x, hd = C / H;
  const T* kvb = kv + (long long)bp * item_stride;
  T* dkvb = dkv + (long long)bp * item_stride;
  const float* pb = probs + (long long)bp * H * K;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_q[c] = q[c];
    s_dc[c] = to_f32<T>(dctx[(long long)bp * C + c]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * K; i += blockDim.x) {  // dp[h][j] = dctx_h . v[j, h]
    const int h = i / K, j = i % K;
    const T* vr = kvb + (long long)j * 2 * C + C + h * hd;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc += s_dc[h * hd + d] * to_f32<T>(vr[d]);
    s_ds[i] = acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < H; h += (blockDim.x >> 5)) {
    float dot = 0.f;
    for (int j = lane; j < K; j += 32) dot += pb[h * K + j] * s_ds[h * K + j];
    dot = warp_sum(dot);
    for (int j = lane; j < K; j += 32) s_ds[h * K + j] = pb[h * K + j] * (s_ds[h * K + j] - dot);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int h = c / hd;
    const float qc = s_q[c], dc = s_dc[c];
    float dq = 0.f;
    for (int j = 0; j < K; ++j) {
      const float ds = s_ds[h * K + j];
      dq += ds * to_f32<T>(kvb[(long long)j * 2 * C + c]);
      dkvb[(long long)j * 2 * C + c] = from_f32<T>(ds * qc);
      dkvb[(long long)j * 2 * C + C + c] = from_f32<T>(pb[h * K + j] * dc);
    }
    dq_part[(long long)bp * C + c] = dq;
  }
}


// ---- vectorised, head-split variant (the AT branch: 64 items x 1188 keys would otherwise occupy 64 of the 148 SMs) -------------
// Block = (item, head group); heads are independent, so no cross-block merge is needed.  16-byte loads along the channel
// dimension; lanes that share a head reduce by shuffles; keys are strided over the warps and the per-warp partial sums are
// combined through shared memory in warp order (deterministic).
template <typename T> struct PoolVec;
template <> struct PoolVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 v = __bfloat1622float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};
template <> struct PoolVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};
constexpr int kPoolMaxChunks = 6;   // 16-byte chunks per lane: head-group width <= 32 * 6 chunks

// s_out[h_local][j] = vec_h . m[j, h] for the block's heads; `off` selects k (0) or v (C) inside the [k | v] rows
template <typename T>
__device__ __forceinline__ void pool_scores(const T* __restrict__ kvb, int off, int c0, int Cb, int hd, int K, int C, const float* s_vec,
                                            float* s_out) {
  using V = PoolVec<T>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int nch = Cb / V::N, group = hd / V::N;   // lanes per head (power of two <= 32)
  for (int j = warp; j < K; j += nw) {
    const T* row = kvb + (long long)j * 2 * C + off + c0;
    for (int v0 = 0; v0 < nch; v0 += 32) {
      const int v = v0 + lane;
      float acc = 0.f;
      if (v < nch) {
        float f[V::N];
        V::load(row + v * V::N, f);
#pragma unroll
        for (int e = 0; e < V::N; ++e) acc += f[e] * s_vec[v * V::N + e];
      }
      for (int o = group >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (v < nch && (lane & (group - 1)) == 0) s_out[(v / group) * K + j] = acc;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attnpool_fwd_vec_kernel(const T* __restrict__ kv, const float* __restrict__ q, T* __restrict__ ctx,
                                                               float* __restrict__ probs, int K, int C, int H, long long item_stride, int hsplit) {
  using V = PoolVec<T>;
  extern __shared__ float sm[];
  const int bp = blockIdx.x / hsplit, hs = blockIdx.x % hsplit;
  const int hd = C / H, Hb = H / hsplit, Cb = Hb * hd, c0 = hs * Cb, nw = blockDim.x >> 5;
  float* s_p = sm;                 // [Hb][K]
  float* s_q = s_p + Hb * K;       // [Cb]
  float* s_part = s_q + Cb;        // [nw][Cb]
  const T* kvb = kv + (long long)bp * item_stride;
  for (int c = threadIdx.x; c < Cb; c += blockDim.x) s_q[c] = q[c0 + c];
  __syncthreads();
  pool_scores<T>(kvb, 0, c0, Cb, hd, K, C, s_q, s_p);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < Hb; h += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < K; j += 32) mx = fmaxf(mx, s_p[h * K + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < K; j += 32) {
      const float e = __expf(s_p[h * K + j] - mx);
      s_p[h * K + j] = e;
      sum += e;
    }
    const float inv = 1.0f / warp_sum(sum);
    for (int j = lane; j < K; j += 32) {
      const float p = s_p[h * K + j] * inv;
      s_p[h * K + j] = p;
      if (probs) probs[((long long)bp * H + hs * Hb + h) * K + j] = p;
    }
  }
  __syncthreads();
  const int nch = Cb / V::N;
  float acc[kPoolMaxChunks][V::N];
#pragma unroll
  for (int i = 0; i < kPoolMaxChunks; ++i)
#pragma unroll
    for (int e = 0; e < V::N; ++e) acc[i][e] = 0.f;
  for (int j = warp; j < K; j += nw) {
    const T* row = kvb + (long long)j * 2 * C + C + c0;
#pragma unroll
    for (int i = 0; i < kPoolMaxChunks; ++i) {
      const int v = i * 32 + lane;
      if (v < nch) {
        float f[V::N];
        V::load(row + v * V::N, f);
        const float p = s_p[((v * V::N) / hd) * K + j];
#pragma unroll
        for (int e = 0; e < V::N; ++e) acc[i][e] += p * f[e];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kPoolMaxChunks; ++i) {
    const int v = i * 32 + lane;
    if (v < nch)
#pragma unroll
      for (int e = 0; e < V::N; ++e) s_part[warp * Cb + v * V::N + e] = acc[i][e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cb; c += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += s_part[w * Cb + c];
    ctx[(long long)bp * C + c0 + c] = from_f32<T>(t);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attnpool_bwd_vec_kernel(const T* __restrict__ kv, const float* __restrict__ q, const float* __restrict__ probs,
                                                               const T* __restrict__ dctx, T* __restrict__ dkv, float* __restrict__ dq_part,
                                                               int K, int C, int H, long long item_stride, int hsplit) {
  using V = PoolVec<T>;
  extern __shared__ float sm[];
  const int bp = blockIdx.x / hsplit, hs = blockIdx.x % hsplit;
  const int hd = C / H, Hb = H / hsplit, Cb = Hb * hd, c0 = hs * Cb, nw = blockDim.x >> 5;
  float* s_ds = sm;                // [Hb][K]: dp, then ds
  float* s_pr = s_ds + Hb * K;     // [Hb][K]: probabilities of this head group
  float* s_q = s_pr + Hb * K;      // [Cb]
  float* s_dc = s_q + Cb;          // [Cb]
  float* s_part = s_dc + Cb;       // [nw][Cb]
  const T* kvb = kv + (long long)bp * item_stride;
  T* dkvb = dkv + (long long)bp * item_stride;
  const float* pb = probs + ((long long)bp * H + hs * Hb) * K;
  for (int c = threadIdx.x; c < Cb; c += blockDim.x) {
    s_q[c] = q[c0 + c];
    s_dc[c] = to_f32<T>(dctx[(long long)bp * C + c0 + c]);
  }
  for (int i = threadIdx.x; i < Hb * K; i += blockDim.x) s_pr[i] = pb[i];
  __syncthreads();
  pool_scores<T>(kvb, C, c0, Cb, hd, K, C, s_dc, s_ds);   // dp[h][j] = dctx_h . v[j, h]
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < Hb; h += nw) {
    float dot = 0.f;
    for (int j = lane; j < K; j += 32) dot += s_pr[h * K + j] * s_ds[h * K + j];
    dot = warp_sum(dot);
    for (int j = lane; j < K; j += 32) s_ds[h * K + j] = s_pr[h * K + j] * (s_ds[h * K + j] - dot);
  }
  __syncthreads();
  const int nch = Cb / V::N;
  float acc[kPoolMaxChunks][V::N];
#pragma unroll
  for (int i = 0; i < kPoolMaxChunks; ++i)
#pragma unroll
    for (int e = 0; e < V::N; ++e) acc[i][e] = 0.f;
  for (int j = warp; j < K; j += nw) {
    const long long ro = (long long)j * 2 * C + c0;
#pragma unroll
    for (int i = 0; i < kPoolMaxChunks; ++i) {
      const int v = i * 32 + lane;
      if (v < nch) {
        const int c = v * V::N, h = c / hd;
        const float ds = s_ds[h * K + j], p = s_pr[h * K + j];
        float f[V::N], dk[V::N], dv[V::N];
        V::load(kvb + ro + c, f);
#pragma unroll
        for (int e = 0; e < V::N; ++e) {
          acc[i][e] += ds * f[e];
          dk[e] = ds * s_q[c + e];
          dv[e] = p * s_dc[c + e];
        }
        V::store(dkvb + ro + c, dk);
        V::store(dkvb + ro + C + c, dv);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kPoolMaxChunks; ++i) {
    const int v = i * 32 + lane;
    if (v < nch)
#pragma unroll
      for (int e = 0; e < V::N; ++e) s_part[warp * Cb + v * V::N + e] = acc[i][e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cb; c += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += s_part[w * Cb + c];
    dq_part[(long long)bp * C + c0 + c] = t;
  }
}

// Head groups per item for the vectorised kernels, or 0 when the shape needs the generic kernel.
static int pool_hsplit(int items, int K, int C, int H, int dtype, const void* kv, const void* other, long long item_stride) {
  const int vec = dtype == T4S_BF16 ? 8 : 4, hd = C / H, group = hd / vec;
  if (hd % vec || group < 1 || group > 32 || (group & (group - 1))) return 0;
  if (((uintptr_t)kv | (uintptr_t)other) % 16 || item_stride % vec || (2LL * C) % vec) return 0;
  int best = 0;
  for (int s = 1; s <= H; ++s) {
    if (H % s) continue;
    const int Cb = (H / s) * hd;
    if (Cb / vec > 32 * kPoolMaxChunks) continue;
    if (best && Cb / vec < 24) break;   // keep at least 24 of the 32 lanes busy
    best = s;
    if ((long long)items * s >= 2LL * sm_count()) break;
  }
  return best;
}


// ---- DASM (detect_any_sound.py:376-388): p = clamp(sigmoid(score / temp) * at_out, 1e-7, 1) with padded frames forced to 0 first,
// weak = clamp(sum p^2 / sum p, 1e-7, 1).  score [B, T, K] fp32 (query x frame GEMM), at_out [B, K]; strong [B, K, T].  Block = (b, k).
__global__ void query_pool_fwd_kernel(const float* __restrict__ score, const float* __restrict__ at_out, const unsigned char* __restrict__ pad_mask,
                                      float inv_temp, float* __restrict__ strong, float* __restrict__ weak, int T, int K) {
  __shared__ float s_red[32];
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  const float at = at_out[blockIdx.x];
  float s1 = 0.f, s2 = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float v = at / (1.0f + expf(-score[((long long)b * T + t) * K + k] * inv_temp));
    if (pad_mask && pad_mask[(long long)b * T + t]) v = 0.f;
    const float p = fminf(fmaxf(v, 1e-7f), 1.0f);
    strong[((long long)b * K + k) * T + t] = p;
    s1 += p;
    s2 += p * p;
  }
  s1 = block_sum(s1, s_red);
  s2 = block_sum(s2, s_red);
  if (threadIdx.x == 0) weak[blockIdx.x] = fminf(fmaxf(s2 / s1, 1e-7f), 1.0f);
}
// dscore[b,t,k] and the fixed-order partial of dat[b,k]; clamp passes gradient only inside [1e-7, 1]
__global__ void query_pool_bwd_kernel(const float* __restrict__ score, const float* __restrict__ at_out, const float* __restrict__ strong,
                                      const float* __restrict__ dstrong, const float* __restrict__ dweak, const unsigned char* __restrict__ pad_mask,
                                      float inv_temp, float* __restrict__ dscore, float* __restrict__ dat, int T, int K) {
  __shared__ float s_red[32];
  const int b = blockIdx.x / K, k = blockIdx.x % K;
  const float at = at_out[blockIdx.x];
  const float* p_row = strong + ((long long)b * K + k) * T;
  float s1 = 0.f, s2 = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float p = p_row[t];
    s1 += p;
    s2 += p * p;
  }
  s1 = block_sum(s1, s_red);
  s2 = block_sum(s2, s_red);
  const float w = s2 / s1;
  float gw = dweak ? dweak[blockIdx.x] : 0.f;
  if (!(w >= 1e-7f && w <= 1.0f)) gw = 0.f;
  float acc = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const long long si = ((long long)b * T + t) * K + k;
    const float sig = 1.0f / (1.0f + expf(-score[si] * inv_temp));
    const float v = (pad_mask && pad_mask[(long long)b * T + t]) ? 0.f : at * sig;
    const float p = p_row[t];
    float g = (dstrong ? dstrong[((long long)b * K + k) * T + t] : 0.f) + gw * (2.f * p * s1 - s2) / (s1 * s1);
    if (!(v >= 1e-7f && v <= 1.0f)) g = 0.f;
    dscore[si] = g * at * sig * (1.f - sig) * inv_temp;
    acc += g * sig;
  }
  acc = block_sum(acc, s_red);
  if (threadIdx.x == 0) dat[blockIdx.x] = acc;
}

// attention scores s [rows, ld] (rows = (b, h, q) with q fastest): entries whose mask[q, c] is set become -inf (boolean attn_mask of
// nn.MultiheadAttention: True = not allowed to attend)
template <typename T>
__global__ void mask_scores_kernel(T* __restrict__ s, const unsigned char* __restrict__ mask, long long rows, int cols, long long ld, int nq) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    if (mask[(r % nq) * cols + c]) s[r * ld + c] = from_f32<T>(-INFINITY);
  }
}

static int grid_for(long long n, int threads = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, (long long)sm_count() * 8));
}

}  // namespace head
}  // namespace t4s

using namespace t4s::head;

extern "C" {

int t4s_sed_pool_fwd(const float* logits, const unsigned char* pad_mask, float temp, float* strong, float* weak, int batch, int frames,
                     int classes, void* stream) {
  T4S_REQUIRE(logits && strong && weak && batch > 0 && frames > 0 && classes > 0 && temp != 0.f, "t4s_sed_pool_fwd: bad arguments");
  sed_pool_fwd_kernel<<<batch * classes, 256, 0, t4s::as_stream(stream)>>>(logits, pad_mask, 1.0f / temp, strong, weak, frames, classes);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_sed_pool_bwd(const float* strong, const float* dstrong, const float* dweak, const unsigned char* pad_mask, float temp, float* dlogits,
                     int batch, int frames, int classes, void* stream) {
  T4S_REQUIRE(strong && dlogits && batch > 0 && frames > 0 && classes > 0 && temp != 0.f, "t4s_sed_pool_bwd: bad arguments");
  sed_pool_bwd_kernel<<<batch * classes, 256, 0, t4s::as_stream(stream)>>>(strong, dstrong, dweak, pad_mask, 1.0f / temp, dlogits, frames, classes);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_sigmoid_fwd(const float* x, float* y, size_t n, void* stream) {
  T4S_REQUIRE(x && y, "t4s_sigmoid_fwd: null pointer");
  if (n) sigmoid_fwd_kernel<<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(x, y, n);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_sigmoid_bwd(const float* y, const float* dy, float* dx, size_t n, void* stream) {
  T4S_REQUIRE(y && dy && dx, "t4s_sigmoid_bwd: null pointer");
  if (n) sigmoid_bwd_kernel<<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(y, dy, dx, n);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

#define T4S_LOSS_PARTS 256

/* out[0] = mean BCE, out[1] = n.  ws: >= 256 floats. */
int t4s_bce_fwd(const float* p, const float* y, size_t n, float* ws, float* out, void* stream) {
  T4S_REQUIRE(p && y && ws && out && n > 0, "t4s_bce_fwd: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  const int parts = (int)std::min<size_t>(T4S_LOSS_PARTS, (n + 255) / 256);
  bce_partial_kernel<<<parts, 256, 0, st>>>(p, y, n, ws);
  T4S_LAUNCH_CHECK();
  finish_mean_kernel<<<1, 32, 0, st>>>(ws, parts, nullptr, (float)n, out);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_bce_bwd(const float* p, const float* y, const float* grad_out, size_t n, float* dp, void* stream) {
  T4S_REQUIRE(p && y && grad_out && dp && n > 0, "t4s_bce_bwd: bad arguments");
  bce_bwd_kernel<<<grid_for((long long)n), 256, 0, t4s::as_stream(stream)>>>(p, y, grad_out, 1.0f / (float)n, dp, n);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

/* out[0] = mean over (selected rows x cols) of (a-b)^2, out[1] = divisor.  ws: >= 512 floats. */
int t4s_mse_fwd(const void* a, const void* b, const unsigned char* row_mask, int64_t rows, int cols, int dtype, float* ws, float* out,
                void* stream) {
  T4S_REQUIRE(a && b && ws && out && rows > 0 && cols > 0, "t4s_mse_fwd: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  const int parts = (int)std::min<long long>(T4S_LOSS_PARTS, rows);
  if (dtype == T4S_F32) mse_partial_kernel<float><<<parts, 256, 0, st>>>((const float*)a, (const float*)b, row_mask, rows, cols, ws, ws + T4S_LOSS_PARTS);
  else if (dtype == T4S_BF16) mse_partial_kernel<__nv_bfloat16><<<parts, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, row_mask, rows, cols, ws, ws + T4S_LOSS_PARTS);
  else { t4s::set_error("t4s_mse_fwd: bad dtype"); return T4S_ERR_ARG; }
  T4S_LAUNCH_CHECK();
  finish_mean_kernel<<<1, 32, 0, st>>>(ws, parts, ws + T4S_LOSS_PARTS, (float)cols, out);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_mse_bwd(const void* a, const void* b, const unsigned char* row_mask, int64_t rows, int cols, int dtype, const float* grad_out,
                const float* fwd_out, void* da, void* db, void* stream) {
  T4S_REQUIRE(a && b && grad_out && fwd_out && (da || db), "t4s_mse_bwd: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  const int grid = grid_for(rows * cols);
  if (dtype == T4S_F32) mse_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, row_mask, rows, cols, grad_out, fwd_out, (float*)da, (float*)db);
  else if (dtype == T4S_BF16) mse_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, row_mask, rows, cols, grad_out, fwd_out, (__nv_bfloat16*)da, (__nv_bfloat16*)db);
  else { t4s::set_error("t4s_mse_bwd: bad dtype"); return T4S_ERR_ARG; }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_query_pool_fwd(const float* score, const float* at_out, const unsigned char* pad_mask, float temp, float* strong, float* weak, int batch,
                        int frames, int queries, void* stream) {
  T4S_REQUIRE(score && at_out && strong && weak && batch > 0 && frames > 0 && queries > 0 && temp != 0.f, "t4s_query_pool_fwd: bad arguments");
  query_pool_fwd_kernel<<<batch * queries, 256, 0, t4s::as_stream(stream)>>>(score, at_out, pad_mask, 1.0f / temp, strong, weak, frames, queries);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_query_pool_bwd(const float* score, const float* at_out, const float* strong, const float* dstrong, const float* dweak,
                       const unsigned char* pad_mask, float temp, float* dscore, float* dat, int batch, int frames, int queries, void* stream) {
  T4S_REQUIRE(score && at_out && strong && dscore && dat && temp != 0.f, "t4s_query_pool_bwd: bad arguments");
  query_pool_bwd_kernel<<<batch * queries, 256, 0, t4s::as_stream(stream)>>>(score, at_out, strong, dstrong, dweak, pad_mask, 1.0f / temp, dscore, dat,
                                                                            frames, queries);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_mask_scores(void* s, const unsigned char* mask, int64_t rows, int cols, int64_t ld, int n_queries, int dtype, void* stream) {
  T4S_REQUIRE(s && mask && rows > 0 && cols > 0 && n_queries > 0, "t4s_mask_scores: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  const int grid = grid_for((long long)rows * cols);
  if (dtype == T4S_F32) mask_scores_kernel<float><<<grid, 256, 0, st>>>((float*)s, mask, rows, cols, ld, n_queries);
  else if (dtype == T4S_BF16) mask_scores_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)s, mask, rows, cols, ld, n_queries);
  else { t4s::set_error("t4s_mask_scores: bad dtype"); return T4S_ERR_ARG; }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_attnpool_fwd(const void* kv, const float* q, void* ctx, float* probs, int items, int keys, int dim, int heads, int64_t item_stride,
                     int dtype, void* stream) {
  if (item_stride <= 0) item_stride = (int64_t)keys * 2 * dim;
  T4S_REQUIRE(kv && q && ctx && items > 0 && keys > 0 && heads > 0 && dim % heads == 0, "t4s_attnpool_fwd: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  if (const int hsplit = pool_hsplit(items, keys, dim, heads, dtype, kv, ctx, item_stride)) {
    const int Hb = heads / hsplit, Cb = Hb * (dim / heads);
    const size_t vsmem = ((size_t)Hb * keys + Cb + 8 * (size_t)Cb) * sizeof(float);
    if (vsmem <= 200 * 1024) {
      if (dtype == T4S_F32) {
        T4S_CUDA(cudaFuncSetAttribute(attnpool_fwd_vec_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
        attnpool_fwd_vec_kernel<float><<<items * hsplit, 256, vsmem, st>>>((const float*)kv, q, (float*)ctx, probs, keys, dim, heads, item_stride, hsplit);
      } else {
        T4S_CUDA(cudaFuncSetAttribute(attnpool_fwd_vec_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
        attnpool_fwd_vec_kernel<__nv_bfloat16><<<items * hsplit, 256, vsmem, st>>>((const __nv_bfloat16*)kv, q, (__nv_bfloat16*)ctx, probs, keys, dim,
                                                                                  heads, item_stride, hsplit);
      }
      T4S_LAUNCH_CHECK();
      return T4S_OK;
    }
  }
  const size_t smem = ((size_t)heads * keys + dim) * sizeof(float);
  T4S_REQUIRE(smem <= 200 * 1024, "t4s_attnpool_fwd: heads*keys too large for shared memory");
  if (dtype == T4S_F32) {
    T4S_CUDA(cudaFuncSetAttribute(attnpool_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attnpool_fwd_kernel<float><<<items, 256, smem, st>>>((const float*)kv, q, (float*)ctx, probs, keys, dim, heads, item_stride);
  } else if (dtype == T4S_BF16) {
    T4S_CUDA(cudaFuncSetAttribute(attnpool_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attnpool_fwd_kernel<__nv_bfloat16><<<items, 256, smem, st>>>((const __nv_bfloat16*)kv, q, (__nv_bfloat16*)ctx, probs, keys, dim, heads, item_stride);
  } else { t4s::set_error("t4s_attnpool_fwd: bad dtype"); return T4S_ERR_ARG; }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_attnpool_bwd(const void* kv, const float* q, const float* probs, const void* dctx, void* dkv, float* dq_part, int items, int keys,
                     int dim, int heads, int64_t item_stride, int dtype, void* stream) {
  if (item_stride <= 0) item_stride = (int64_t)keys * 2 * dim;
  T4S_REQUIRE(kv && q && probs && dctx && dkv && dq_part && items > 0 && keys > 0 && heads > 0 && dim % heads == 0, "t4s_attnpool_bwd: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  if (const int hsplit = pool_hsplit(items, keys, dim, heads, dtype, kv, dkv, item_stride)) {
    const int Hb = heads / hsplit, Cb = Hb * (dim / heads);
    const size_t vsmem = (2 * (size_t)Hb * keys + 2 * Cb + 8 * (size_t)Cb) * sizeof(float);
    if (vsmem <= 200 * 1024 && (uintptr_t)dctx % 16 == 0) {
      if (dtype == T4S_F32) {
        T4S_CUDA(cudaFuncSetAttribute(attnpool_bwd_vec_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
        attnpool_bwd_vec_kernel<float><<<items * hsplit, 256, vsmem, st>>>((const float*)kv, q, probs, (const float*)dctx, (float*)dkv, dq_part, keys,
                                                                          dim, heads, item_stride, hsplit);
      } else {
        T4S_CUDA(cudaFuncSetAttribute(attnpool_bwd_vec_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem));
        attnpool_bwd_vec_kernel<__nv_bfloat16><<<items * hsplit, 256, vsmem, st>>>((const __nv_bfloat16*)kv, q, probs, (const __nv_bfloat16*)dctx,
                                                                                  (__nv_bfloat16*)dkv, dq_part, keys, dim, heads, item_stride, hsplit);
      }
      T4S_LAUNCH_CHECK();
      return T4S_OK;
    }
  }
  const size_t smem = ((size_t)heads * keys + 2 * dim) * sizeof(float);
  T4S_REQUIRE(smem <= 200 * 1024, "t4s_attnpool_bwd: heads*keys too large for shared memory");
  if (dtype == T4S_F32) {
    T4S_CUDA(cudaFuncSetAttribute(attnpool_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attnpool_bwd_kernel<float><<<items, 256, smem, st>>>((const float*)kv, q, probs, (const float*)dctx, (float*)dkv, dq_part, keys, dim, heads, item_stride);
  } else if (dtype == T4S_BF16) {
    T4S_CUDA(cudaFuncSetAttribute(attnpool_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attnpool_bwd_kernel<__nv_bfloat16><<<items, 256, smem, st>>>((const __nv_bfloat16*)kv, q, probs, (const __nv_bfloat16*)dctx, (__nv_bfloat16*)dkv, dq_part, keys, dim, heads, item_stride);
  } else { t4s::set_error("t4s_attnpool_bwd: bad dtype"); return T4S_ERR_ARG; }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
