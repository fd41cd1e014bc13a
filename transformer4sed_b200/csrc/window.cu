// Sliding-window global-local fusion and MLM frame masking (sm_100a):
//   * windowed patch im2col: every time window of the mel image becomes one more sequence in the batch, so the 11 / 17
//     window passes of the reference's Python loop (encoder_slide_window.py:29-33) are ONE backbone pass;
//   * overlap-add mean of the per-window frame embeddings (encoder_slide_window.py:24-36), forward and backward;
//   * masked-frame replacement (transformer/mask.py:62-82): mask token / random other frame / keep, forward and backward.
// All three are HBM-bound gathers: one read of the sources and one write of the destination, 16-byte vectors along C.
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace window {

struct Starts { int v[T4S_MAX_WINDOWS]; };

struct Segments {
  const void* ptr[T4S_MAX_WINDOWS];      // window w, clip 0, frame 0
  long long bstride[T4S_MAX_WINDOWS];    // elements between clips of window w
  int start[T4S_MAX_WINDOWS];            // first output frame of window w
  int len[T4S_MAX_WINDOWS];              // frames of window w that land inside the output
  int n;
};
struct SegmentsOut {
  void* ptr[T4S_MAX_WINDOWS];
  long long bstride[T4S_MAX_WINDOWS];
  int start[T4S_MAX_WINDOWS];
  int len[T4S_MAX_WINDOWS];
  int full[T4S_MAX_WINDOWS];             // frames window w really has (>= len; the tail beyond the output gets zero gradient)
  int n;
};

static int grid_for(long long n, int threads = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, (long long)sm_count() * 16));
}

// A[(w, b, f, t), i*P + j] = img[b, f*S + i, start_w + t*S + j]      (passt.py:302-315 applied to input[:, :, w_left:w_right])
template <typename TI, typename TO>
__global__ void im2col_windows_kernel(const TI* __restrict__ img, TO* __restrict__ out, int B, int H, int W, Starts starts, int nW,
                                      int F, int Tp, int P, int S) {
  const long long total = (long long)nW * B * F * Tp * P;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % P);
    long long r = idx / P;
    const int t = (int)(r % Tp);
    r /= Tp;
    const int f = (int)(r % F);
    r /= F;
    const int b = (int)(r % B);
    const int w = (int)(r / B);
    const TI* src = img + ((long long)b * H + f * S + i) * W + starts.v[w] + t * S;
    TO* dst = out + idx * P;
    for (int j = 0; j < P; ++j) dst[j] = from_f32<TO>(to_f32<TI>(src[j]));
  }
}

template <typename T> struct Vec8;  // 8 elements for bf16 (16 bytes), 4 for fp32 (16 bytes)
template <> struct Vec8<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(&raw)[i]); }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<__nv_bfloat16*>(&raw)[i] = __float2bfloat16_rn(v); }
};
template <> struct Vec8<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return reinterpret_cast<const float*>(&raw)[i]; }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<float*>(&raw)[i] = v; }
};

// out[b, t, :] = (sum over windows w covering t of local_w[b, t - start_w, :]) / #covering windows; 0 where none covers
// (embedding /= accumlator; nan -> 0: encoder_slide_window.py:35-36)
template <typename T>
__global__ void overlap_add_fwd_kernel(Segments seg, T* __restrict__ out, int B, int frames, int C) {
  using V = Vec8<T>;
  const int cv = C / V::N;
  const long long total = (long long)B * frames * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V::N;
    const long long r = idx / cv;
    const int t = (int)(r % frames), b = (int)(r / frames);
    float acc[V::N];
#pragma unroll
    for (int k = 0; k < V::N; ++k) acc[k] = 0.f;
    int cnt = 0;
    for (int w = 0; w < seg.n; ++w) {
      const int rel = t - seg.start[w];
      if (rel < 0 || rel >= seg.len[w]) continue;
      V v;
      v.load(static_cast<const T*>(seg.ptr[w]) + (long long)b * seg.bstride[w] + (long long)rel * C + c);
#pragma unroll
      for (int k = 0; k < V::N; ++k) acc[k] += v.get(k);
      ++cnt;
    }
    const float inv = cnt ? 1.0f / (float)cnt : 0.f;
    V o;
#pragma unroll
    for (int k = 0; k < V::N; ++k) o.set(k, acc[k] * inv);
    o.store(out + ((long long)b * frames + t) * C + c);
  }
}

// dlocal_w[b, rel, :] = dout[b, start_w + rel, :] / count(start_w + rel)   (zero for frames cut off at the end of the clip)
template <typename T>
__global__ void overlap_add_bwd_kernel(const T* __restrict__ dout, SegmentsOut seg, int w, int B, int frames, int C) {
  using V = Vec8<T>;
  const int cv = C / V::N;
  const int full = seg.full[w];
  const long long total = (long long)B * full * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V::N;
    const long long r = idx / cv;
    const int rel = (int)(r % full), b = (int)(r / full);
    V o;
    if (rel < seg.len[w]) {
      const int t = seg.start[w] + rel;
      int cnt = 0;
      for (int u = 0; u < seg.n; ++u) cnt += (t >= seg.start[u] && t < seg.start[u] + seg.len[u]) ? 1 : 0;
      const float inv = 1.0f / (float)cnt;  // cnt >= 1: window w itself covers t
      V g;
      g.load(dout + ((long long)b * frames + t) * C + c);
#pragma unroll
      for (int k = 0; k < V::N; ++k) o.set(k, g.get(k) * inv);
    } else {
#pragma unroll
      for (int k = 0; k < V::N; ++k) o.set(k, 0.f);
    }
    o.store(static_cast<T*>(seg.ptr[w]) + (long long)b * seg.bstride[w] + (long long)rel * C + c);
  }
}

// ---- masked-frame replacement: kind[r] = 0 keep, 1 mask token, 2 copy row src[r] of the ORIGINAL sequence ----------------
template <typename T>
__global__ void mask_rows_fwd_kernel(const T* __restrict__ x, const float* __restrict__ token, const unsigned char* __restrict__ kind,
                                     const long long* __restrict__ src, T* __restrict__ out, long long rows, int C) {
  using V = Vec8<T>;
  const int cv = C / V::N;
  const long long total = rows * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V::N;
    const long long r = idx / cv;
    const int k = kind[r];
    V v;
    if (k == 1) {
#pragma unroll
      for (int e = 0; e < V::N; ++e) v.set(e, token[c + e]);
    } else {
      v.load(x + (k == 2 ? src[r] : r) * C + c);
    }
    v.store(out + r * C + c);
  }
}

// dx[r] = (kind[r] == 0 ? dout[r] : 0) + sum over rows i with kind[i] == 2 and src[i] == r of dout[i], summed in ascending i
// (deterministic).  `list` holds the n_list row indices with kind == 2 in ascending order.
template <typename T>
__global__ void mask_rows_bwd_kernel(const T* __restrict__ dout, const unsigned char* __restrict__ kind, const long long* __restrict__ src,
                                     const long long* __restrict__ list, int n_list, T* __restrict__ dx, long long rows, int C) {
  using V = Vec8<T>;
  __shared__ int s_cnt;
  __shared__ long long s_hit[64];
  const int cv = C / V::N;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_list; i += blockDim.x) {
      const long long row = list[i];
      if (src[row] == r) {
        const int slot = atomicAdd(&s_cnt, 1);
        if (slot < 64) s_hit[slot] = row;
      }
    }
    __syncthreads();
    const bool overflow = s_cnt > 64;  // (practically never) more than 64 copies of one frame: ordered scan of the whole list instead
    const int n = overflow ? 0 : s_cnt;
    if (threadIdx.x == 0 && n > 1) {  // ascending order -> summation order independent of the atomics
      for (int a = 1; a < n; ++a) {
        const long long key = s_hit[a];
        int b = a - 1;
        while (b >= 0 && s_hit[b] > key) { s_hit[b + 1] = s_hit[b]; --b; }
        s_hit[b + 1] = key;
      }
    }
    __syncthreads();
    const bool keep = kind[r] == 0;
    for (int v = threadIdx.x; v < cv; v += blockDim.x) {
      const int c = v * V::N;
      float acc[V::N];
      V g;
      if (keep) {
        g.load(dout + r * C + c);
#pragma unroll
        for (int e = 0; e < V::N; ++e) acc[e] = g.get(e);
      } else {
#pragma unroll
        for (int e = 0; e < V::N; ++e) acc[e] = 0.f;
      }
      for (int h = 0; h < n; ++h) {
        g.load(dout + s_hit[h] * C + c);
#pragma unroll
        for (int e = 0; e < V::N; ++e) acc[e] += g.get(e);
      }
      if (overflow) {
        for (int i = 0; i < n_list; ++i) {
          if (src[list[i]] != r) continue;
          g.load(dout + list[i] * C + c);
#pragma unroll
          for (int e = 0; e < V::N; ++e) acc[e] += g.get(e);
        }
      }
      V o;
#pragma unroll
      for (int e = 0; e < V::N; ++e) o.set(e, acc[e]);
      o.store(dx + r * C + c);
    }
    __syncthreads();
  }
}

// d_token[c] = sum over rows with kind == 1 of dout[r, c]  (fixed block partition + ordered second stage: deterministic)
template <typename T>
__global__ void mask_token_grad_kernel(const T* __restrict__ dout, const unsigned char* __restrict__ kind, float* __restrict__ part,
                                       long long rows, int C, long long rows_per_block) {
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r)
      if (kind[r] == 1) acc += to_f32<T>(dout[r * C + c]);
    part[(long long)blockIdx.x * C + c] = acc;
  }
}
__global__ void mask_token_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int blocks, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int b = 0; b < blocks; ++b) acc += part[(long long)b * C + c];
  out[c] = acc;
}

}  // namespace window
}  // namespace t4s

using namespace t4s::window;

#define T4S_DISPATCH_DTYPE(dtype, ...)                                   \
  do {                                                                   \
    if ((dtype) == T4S_F32) { using T = float; __VA_ARGS__; }            \
    else if ((dtype) == T4S_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { t4s::set_error("bad dtype %d", (int)(dtype)); return T4S_ERR_ARG; } \
  } while (0)

extern "C" {

int t4s_patch_im2col_windows(const void* img, int img_dtype, void* out, int out_dtype, int batch, int height, int width, const int* starts,
                             int n_windows, int patch, int stride, int f_dim, int t_dim, void* stream) {
  T4S_REQUIRE(img && out && starts && batch > 0 && patch > 0 && stride > 0 && f_dim > 0 && t_dim > 0, "t4s_patch_im2col_windows: bad arguments");
  T4S_REQUIRE(n_windows > 0 && n_windows <= T4S_MAX_WINDOWS, "t4s_patch_im2col_windows: %d windows (max %d)", n_windows, T4S_MAX_WINDOWS);
  T4S_REQUIRE((f_dim - 1) * stride + patch <= height, "t4s_patch_im2col_windows: %d frequency patches do not fit %d mel bins", f_dim, height);
  Starts s;
  for (int w = 0; w < n_windows; ++w) {
    T4S_REQUIRE(starts[w] >= 0 && starts[w] + (t_dim - 1) * stride + patch <= width,
                "t4s_patch_im2col_windows: window %d (start %d, %d patches) leaves the %d-frame image", w, starts[w], t_dim, width);
    s.v[w] = starts[w];
  }
  const long long total = (long long)n_windows * batch * f_dim * t_dim * patch;
  cudaStream_t st = t4s::as_stream(stream);
  const int grid = grid_for(total);
#define T4S_IM2COL_W(TI, TO) \
  im2col_windows_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)img, (TO*)out, batch, height, width, s, n_windows, f_dim, t_dim, patch, stride)
  if (img_dtype == T4S_F32 && out_dtype == T4S_F32) T4S_IM2COL_W(float, float);
  else if (img_dtype == T4S_F32 && out_dtype == T4S_BF16) T4S_IM2COL_W(float, __nv_bfloat16);
  else if (img_dtype == T4S_BF16 && out_dtype == T4S_BF16) T4S_IM2COL_W(__nv_bfloat16, __nv_bfloat16);
  else if (img_dtype == T4S_BF16 && out_dtype == T4S_F32) T4S_IM2COL_W(__nv_bfloat16, float);
  else { t4s::set_error("t4s_patch_im2col_windows: bad dtypes"); return T4S_ERR_ARG; }
#undef T4S_IM2COL_W
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

static int check_segments(const T4sWindowSegment* segs, int n, int frames, int dim, int dtype, const char* who) {
  T4S_REQUIRE(segs && n > 0 && n <= T4S_MAX_WINDOWS, "%s: %d windows (max %d)", who, n, T4S_MAX_WINDOWS);
  const int vec = dtype == T4S_BF16 ? 8 : 4;
  T4S_REQUIRE(dim % vec == 0, "%s: dim %d must be a multiple of %d", who, dim, vec);
  for (int w = 0; w < n; ++w)
    T4S_REQUIRE(segs[w].ptr && segs[w].out_start >= 0 && segs[w].frames > 0 && segs[w].out_start < frames && ((uintptr_t)segs[w].ptr % 16) == 0 &&
                    segs[w].batch_stride % vec == 0,
                "%s: bad segment %d", who, w);
  return T4S_OK;
}

int t4s_window_overlap_add_fwd(const T4sWindowSegment* segs, int n_windows, void* out, int dtype, int batch, int frames, int dim, void* stream) {
  T4S_REQUIRE(out && batch > 0 && frames > 0, "t4s_window_overlap_add_fwd: bad arguments");
  if (int rc = check_segments(segs, n_windows, frames, dim, dtype, "t4s_window_overlap_add_fwd")) return rc;
  Segments s;
  s.n = n_windows;
  for (int w = 0; w < n_windows; ++w) {
    s.ptr[w] = segs[w].ptr;
    s.bstride[w] = segs[w].batch_stride;
    s.start[w] = segs[w].out_start;
    s.len[w] = std::min(segs[w].frames, frames - segs[w].out_start);
  }
  const int vec = dtype == T4S_BF16 ? 8 : 4;
  const long long total = (long long)batch * frames * (dim / vec);
  T4S_DISPATCH_DTYPE(dtype, (overlap_add_fwd_kernel<T><<<grid_for(total), 256, 0, t4s::as_stream(stream)>>>(s, static_cast<T*>(out), batch, frames, dim)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_window_overlap_add_bwd(const void* dout, const T4sWindowSegment* segs, int n_windows, int dtype, int batch, int frames, int dim,
                               void* stream) {
  T4S_REQUIRE(dout && batch > 0 && frames > 0, "t4s_window_overlap_add_bwd: bad arguments");
  if (int rc = check_segments(segs, n_windows, frames, dim, dtype, "t4s_window_overlap_add_bwd")) return rc;
  SegmentsOut s;
  s.n = n_windows;
  for (int w = 0; w < n_windows; ++w) {
    s.ptr[w] = segs[w].ptr;
    s.bstride[w] = segs[w].batch_stride;
    s.start[w] = segs[w].out_start;
    s.full[w] = segs[w].frames;
    s.len[w] = std::min(segs[w].frames, frames - segs[w].out_start);
  }
  const int vec = dtype == T4S_BF16 ? 8 : 4;
  cudaStream_t st = t4s::as_stream(stream);
  for (int w = 0; w < n_windows; ++w) {
    const long long total = (long long)batch * s.full[w] * (dim / vec);
    T4S_DISPATCH_DTYPE(dtype, (overlap_add_bwd_kernel<T><<<grid_for(total), 256, 0, st>>>(static_cast<const T*>(dout), s, w, batch, frames, dim)));
    T4S_LAUNCH_CHECK();
  }
  return T4S_OK;
}

int t4s_mask_rows_fwd(const void* x, const float* token, const unsigned char* kind, const int64_t* src, void* out, int64_t rows, int dim,
                      int dtype, void* stream) {
  T4S_REQUIRE(x && token && kind && src && out && rows > 0, "t4s_mask_rows_fwd: bad arguments");
  const int vec = dtype == T4S_BF16 ? 8 : 4;
  T4S_REQUIRE(dim % vec == 0, "t4s_mask_rows_fwd: dim %d must be a multiple of %d", dim, vec);
  const long long total = (long long)rows * (dim / vec);
  T4S_DISPATCH_DTYPE(dtype, (mask_rows_fwd_kernel<T><<<grid_for(total), 256, 0, t4s::as_stream(stream)>>>(
                                static_cast<const T*>(x), token, kind, (const long long*)src, static_cast<T*>(out), rows, dim)));
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

size_t t4s_mask_rows_bwd_workspace(int64_t rows, int dim) {
  (void)rows;
  return (size_t)T4S_MASK_GRAD_BLOCKS * dim * sizeof(float);
}

int t4s_mask_rows_bwd(const void* dout, const unsigned char* kind, const int64_t* src, const int64_t* copy_rows, int n_copy_rows, void* dx,
                      float* d_token, float* ws, size_t ws_bytes, int64_t rows, int dim, int dtype, void* stream) {
  T4S_REQUIRE(dout && kind && src && rows > 0 && (n_copy_rows == 0 || copy_rows), "t4s_mask_rows_bwd: bad arguments");
  const int vec = dtype == T4S_BF16 ? 8 : 4;
  T4S_REQUIRE(dim % vec == 0, "t4s_mask_rows_bwd: dim %d must be a multiple of %d", dim, vec);
  cudaStream_t st = t4s::as_stream(stream);
  if (dx) {
    const int grid = (int)std::min<long long>(rows, (long long)t4s::sm_count() * 8);
    T4S_DISPATCH_DTYPE(dtype, (mask_rows_bwd_kernel<T><<<grid, 128, 0, st>>>(static_cast<const T*>(dout), kind, (const long long*)src,
                                                                           (const long long*)copy_rows, n_copy_rows, static_cast<T*>(dx), rows, dim)));
    T4S_LAUNCH_CHECK();
  }
  if (d_token) {
    T4S_REQUIRE(ws && ws_bytes >= t4s_mask_rows_bwd_workspace(rows, dim), "t4s_mask_rows_bwd: workspace too small");
    const int blocks = T4S_MASK_GRAD_BLOCKS;
    const long long per = (rows + blocks - 1) / blocks;
    T4S_DISPATCH_DTYPE(dtype, (mask_token_grad_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(dout), kind, ws, rows, dim, per)));
    T4S_LAUNCH_CHECK();
    mask_token_reduce_kernel<<<(dim + 127) / 128, 128, 0, st>>>(ws, d_token, blocks, dim);
    T4S_LAUNCH_CHECK();
  }
  return T4S_OK;
}

}  // extern "C"
