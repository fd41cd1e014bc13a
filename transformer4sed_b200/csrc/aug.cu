// Train-loop glue next to the hot path: per-clip circular shift (frame_shift), mixup and the class-wise median
// filter of the post-processing, each one read + one write of fp32 data.
//   frame_shift  : src/preprocess/data_aug.py:12-31   (torch.roll per sample; the shifts are drawn on the host with the reference's RNG)
//   mixup        : src/preprocess/data_aug.py:34-91   (c x + (1-c) x[perm], labels clamped to [0, 1])
//   median filter: src/postprocess/filter.py:4-36     (per-class odd window, replicate padding, exact median)
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace aug {

static int grid_for(long long n, int threads = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, (long long)sm_count() * 16));
}

// out[b, r, (t + shift_b) mod len] = x[b, r, t]   (torch.roll(x[b], shift_b, dims=-1))
__global__ void roll_kernel(const float* __restrict__ x, float* __restrict__ out, const int* __restrict__ shifts, int B, int rows, int len) {
  const long long total = (long long)B * rows * len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % len);
    const long long br = i / len;
    const int b = (int)(br / rows);
    int src = (t - shifts[b]) % len;
    if (src < 0) src += len;
    out[i] = x[br * len + src];
  }
}

// out[b, :] = wa x[b, :] + wb x[perm[b], :], optionally clamped to [0, 1]
__global__ void mixup_kernel(const float* __restrict__ x, const long long* __restrict__ perm, float* __restrict__ out, int B, long long inner,
                             float wa, float wb, int clamp01) {
  const long long total = (long long)B * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / inner, j = i - b * inner;
    float v = wa * x[i] + wb * x[perm[b] * inner + j];
    if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = v;
  }
}

// freq_nonlinear (data_aug.py:239-254): every (b, t) column of the mel image is re-sampled along frequency with np.interp over the
// SAME warped knots, so the source bin j0[k] and weight w[k] of every output bin are precomputed once on the host (float64, like
// numpy) and the kernel is a two-row gather + lerp.  out[b, k, t] = x[b, j0[k], t] + w[k] (x[b, j0[k]+1, t] - x[b, j0[k], t])
__global__ void freq_warp_kernel(const float* __restrict__ x, float* __restrict__ out, const int* __restrict__ j0, const float* __restrict__ w,
                                 int B, int F, int T) {
  const long long total = (long long)B * F * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long bf = i / T;
    const int k = (int)(bf % F);
    const long long b = bf / F;
    const int j = j0[k];
    const float a = x[(b * F + j) * T + t];
    const float wk = w[k];
    out[i] = wk == 0.f ? a : fmaf(wk, x[(b * F + min(j + 1, F - 1)) * T + t] - a, a);
  }
}

// out[b, f, t] = x[b, f, t] + bias[b, f]   (FilterAugment on log-mel features: + log(filter + 1e-5) / norm_std, data_aug.py:188-190)
__global__ void add_rowbias_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ out, long long rows, int T) {
  const long long total = rows * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    out[i] = x[i] + bias[i / T];
}

struct Sizes { int v[T4S_MEDIAN_MAX_CLASSES]; };

// in / out [B, L, C]; class c uses the odd window sizes.v[c] (<= T4S_MEDIAN_MAX_WINDOW) with replicate padding.  Block = 256
// consecutive frames of one (b, c): the frames and their halo are staged in shared memory, every thread ranks its window.
__global__ void __launch_bounds__(256) median_kernel(const float* __restrict__ in, float* __restrict__ out, Sizes sizes, int B, int L, int C) {
  __shared__ float s[256 + T4S_MEDIAN_MAX_WINDOW];
  const int c = blockIdx.y, b = blockIdx.z, l0 = blockIdx.x * 256;
  const int k = sizes.v[c], half = k / 2;
  for (int i = threadIdx.x; i < 256 + 2 * half; i += 256) {
    const int l = min(max(l0 + i - half, 0), L - 1);          // replicate padding
    s[i] = in[((long long)b * L + l) * C + c];
  }
  __syncthreads();
  const int l = l0 + threadIdx.x;
  if (l >= L) return;
  const float* w = s + threadIdx.x;   // window = w[0 .. k)
  float med = w[half];
  for (int i = 0; i < k; ++i) {
    const float vi = w[i];
    int rank = 0;
    for (int j = 0; j < k; ++j) rank += (w[j] < vi || (w[j] == vi && j < i)) ? 1 : 0;
    if (rank == half) med = vi;
  }
  out[((long long)b * L + l) * C + c] = med;
}

}  // namespace aug
}  // namespace t4s

extern "C" {

int t4s_roll_rows(const float* x, float* out, const int* shifts_dev, int batch, int rows, int len, void* stream) {
  T4S_REQUIRE(x && out && shifts_dev && batch > 0 && rows > 0 && len > 0 && x != out, "t4s_roll_rows: bad arguments");
  const long long total = (long long)batch * rows * len;
  t4s::aug::roll_kernel<<<t4s::aug::grid_for(total), 256, 0, t4s::as_stream(stream)>>>(x, out, shifts_dev, batch, rows, len);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_mixup(const float* x, const int64_t* perm_dev, float* out, int batch, int64_t inner, float w_self, float w_other, int clamp01, void* stream) {
  T4S_REQUIRE(x && perm_dev && out && batch > 0 && inner > 0 && x != out, "t4s_mixup: bad arguments");
  const long long total = (long long)batch * inner;
  t4s::aug::mixup_kernel<<<t4s::aug::grid_for(total), 256, 0, t4s::as_stream(stream)>>>(x, (const long long*)perm_dev, out, batch, inner, w_self, w_other,
                                                                                   clamp01);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_freq_warp(const float* x, float* out, const int* src_bin_dev, const float* weight_dev, int batch, int n_freq, int n_frames, void* stream) {
  T4S_REQUIRE(x && out && src_bin_dev && weight_dev && batch > 0 && n_freq > 0 && n_frames > 0 && x != out, "t4s_freq_warp: bad arguments");
  const long long total = (long long)batch * n_freq * n_frames;
  t4s::aug::freq_warp_kernel<<<t4s::aug::grid_for(total), 256, 0, t4s::as_stream(stream)>>>(x, out, src_bin_dev, weight_dev, batch, n_freq, n_frames);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_add_rowbias(const float* x, const float* bias_dev, float* out, int64_t rows, int cols, void* stream) {
  T4S_REQUIRE(x && bias_dev && out && rows > 0 && cols > 0, "t4s_add_rowbias: bad arguments");
  t4s::aug::add_rowbias_kernel<<<t4s::aug::grid_for(rows * cols), 256, 0, t4s::as_stream(stream)>>>(x, bias_dev, out, rows, cols);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_median_filter(const float* in, float* out, const int* window_sizes, int batch, int length, int classes, void* stream) {
  T4S_REQUIRE(in && out && window_sizes && batch > 0 && length > 0 && classes > 0 && classes <= T4S_MEDIAN_MAX_CLASSES && in != out,
              "t4s_median_filter: bad arguments (classes <= %d)", T4S_MEDIAN_MAX_CLASSES);
  t4s::aug::Sizes s;
  for (int c = 0; c < classes; ++c) {
    T4S_REQUIRE(window_sizes[c] >= 1 && (window_sizes[c] & 1) && window_sizes[c] <= T4S_MEDIAN_MAX_WINDOW,
                "t4s_median_filter: window %d of class %d must be odd and <= %d", window_sizes[c], c, T4S_MEDIAN_MAX_WINDOW);
    s.v[c] = window_sizes[c];
  }
  dim3 grid((length + 255) / 256, classes, batch);
  t4s::aug::median_kernel<<<grid, 256, 0, t4s::as_stream(stream)>>>(in, out, s, batch, length, classes);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
