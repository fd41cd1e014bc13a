// Callers either side of the hot path (SURVEY §8 a1', f3, f4):
//   scaler       : src/preprocess/scaler.py:91-121       TorchScaler.forward, statistics over dims (1, 2) of a [B, F, T] tensor
//   rank filter  : src/codec/decoder.py:86-92            scipy.ndimage median_filter / maximum_filter per class (any window, 'reflect')
//   event sweep  : src/codec/decoder.py:15-35, src/codec/encoder.py:51-84   threshold sweep + run-length (onset, offset) decoding
//   SED losses   : recipes/desed/finetune/train.py:166-188   the six mean-teacher losses and their weighted total, one pass
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace post {

__device__ __forceinline__ float block_sum256(float v, float* s_red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) s_red[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  if (threadIdx.x == 0) s_red[0] = v;
  __syncthreads();
  v = s_red[0];
  return v;
}
__device__ __forceinline__ float block_max256(float v, float* s_red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) s_red[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : -INFINITY;
  if (w == 0) v = warp_max(v);
  if (threadIdx.x == 0) s_red[0] = v;
  __syncthreads();
  v = s_red[0];
  return v;
}

// ---- TorchScaler ----------------------------------------------------------------------------------------------------------
// One CTA per instance (clip): mode 0 'mean', 1 'standard' (unbiased std, like torch.std), 2 'minmax'.  Two reads, one write.
__global__ void __launch_bounds__(256) scaler_instance_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int mode, float eps) {
  __shared__ float s_red[32];
  const float* xi = x + (long long)blockIdx.x * n;
  float* oi = out + (long long)blockIdx.x * n;
  if (mode == 2) {
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = threadIdx.x; i < n; i += 256) {
      const float v = xi[i];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    hi = block_max256(hi, s_red);
    lo = -block_max256(-lo, s_red);
    const float den = hi - lo + eps;
    for (long long i = threadIdx.x; i < n; i += 256) oi[i] = (xi[i] - lo) / den;
    return;
  }
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) s += xi[i];
  const float mean = block_sum256(s, s_red) / (float)n;
  float den = 1.f;
  if (mode == 1) {
    float q = 0.f;
    for (long long i = threadIdx.x; i < n; i += 256) {
      const float d = xi[i] - mean;
      q = fmaf(d, d, q);
    }
    q = block_sum256(q, s_red);
    den = sqrtf(q / (float)(n - 1)) + eps;
  }
  for (long long i = threadIdx.x; i < n; i += 256) oi[i] = (xi[i] - mean) / den;
}
// dataset statistics: out = (x - mean) / den with scalars read from device memory (den = 1 for normtype 'mean')
__global__ void scaler_affine_kernel(const float* __restrict__ x, float* __restrict__ out, long long total, const float* __restrict__ mean,
                                     const float* __restrict__ mean_sq, int standard, float eps) {
  const float m = mean[0];
  const float den = standard ? sqrtf(mean_sq[0] - m * m) + eps : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) out[i] = (x[i] - m) / den;
}

// ---- per-class rank filter along time ------------------------------------------------------------------------------------------
struct Sizes {
  int v[T4S_MEDIAN_MAX_CLASSES];
};
// in / out [B, L, C].  Window of class c: k = sizes.v[c] samples starting at l - k / 2 (scipy.ndimage's centring, any k >= 1),
// borders by 'reflect' in scipy's sense (d c b a | a b c d | d c b a).  op 0: element of rank k / 2 (median), op 1: maximum.
__global__ void __launch_bounds__(256) rank_filter_kernel(const float* __restrict__ in, float* __restrict__ out, Sizes sizes, int B, int L, int C, int op) {
  __shared__ float s[256 + T4S_MEDIAN_MAX_WINDOW];
  const int c = blockIdx.y, b = blockIdx.z, l0 = blockIdx.x * 256;
  const int k = sizes.v[c], left = k / 2;
  for (int i = threadIdx.x; i < 256 + k - 1; i += 256) {
    int l = l0 + i - left;
    // symmetric reflection, repeated for windows longer than the signal
    const int period = 2 * L;
    l %= period;
    if (l < 0) l += period;
    if (l >= L) l = period - 1 - l;
    s[i] = in[((long long)b * L + l) * C + c];
  }
  __syncthreads();
  const int l = l0 + threadIdx.x;
  if (l >= L) return;
  const float* w = s + threadIdx.x;   // window = w[0 .. k)
  float res = w[0];
  if (op == 1) {
    for (int i = 1; i < k; ++i) res = fmaxf(res, w[i]);
  } else {
    const int want = k / 2;
    for (int i = 0; i < k; ++i) {
      const float vi = w[i];
      int rank = 0;
      for (int j = 0; j < k; ++j) rank += (w[j] < vi || (w[j] == vi && j < i)) ? 1 : 0;
      if (rank == want) res = vi;
    }
  }
  out[((long long)b * L + l) * C + c] = res;
}

// ---- threshold sweep + event decoding --------------------------------------------------------------------------------------------
// scores [B, L, C] (already filtered), weak [B, C] or NULL, thresholds [n_th].  Column (th, b, c) is active where
// weak[b, c] >= th (decode_pred_batch_fast zeroes the others before filtering: a zero column stays zero) and score > th.
// Events are the maximal runs of active frames, (onset frame, offset frame) = (first, last + 1) as find_contiguous_regions returns.
// Pass 1 (events == NULL) writes counts[th, b, c]; pass 2 writes the events of column q at events[offsets[q] ...] as
// (th, b, c, onset, offset).
__global__ void event_sweep_kernel(const float* __restrict__ scores, const float* __restrict__ weak, const float* __restrict__ th, int n_th, int B, int L,
                                   int C, int* __restrict__ counts, const long long* __restrict__ offsets, int* __restrict__ events) {
  const long long total = (long long)n_th * B * C;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(q % C);
    const int b = (int)((q / C) % B);
    const int t = (int)(q / ((long long)C * B));
    const float thr = th[t];
    int n = 0;
    if (!weak || !(weak[(long long)b * C + c] < thr)) {
      const float* col = scores + (long long)b * L * C + c;
      int* ev = events ? events + offsets[q] * 5 : nullptr;
      bool prev = false;
      int onset = 0;
      for (int l = 0; l < L; ++l) {
        const bool cur = col[(long long)l * C] > thr;
        if (cur && !prev) onset = l;
        if (!cur && prev) {
          if (ev) { ev[5 * n] = t; ev[5 * n + 1] = b; ev[5 * n + 2] = c; ev[5 * n + 3] = onset; ev[5 * n + 4] = l; }
          ++n;
        }
        prev = cur;
      }
      if (prev) {
        if (ev) { ev[5 * n] = t; ev[5 * n + 1] = b; ev[5 * n + 2] = c; ev[5 * n + 3] = onset; ev[5 * n + 4] = L; }
        ++n;
      }
    }
    if (!events) counts[q] = n;
  }
}

// ---- the six mean-teacher losses ---------------------------------------------------------------------------------------------------
// strong [B, n_s] (n_s = classes x frames, any fixed layout shared by student / teacher / labels), weak / at [B, C].
// Rows [s0, s1) carry strong labels, rows [w0, w1) weak labels (get_mask, train.py:80-91).
//   part[0] BCE(strong[s0:s1], y)   part[1] BCE(weak[w0:w1], yw)   part[2] BCE(at[w0:w1], yw)
//   part[3] MSE(strong, t_strong)   part[4] MSE(weak, t_at)         part[5] MSE(at, t_at)
struct LossArgs {
  const float *strong, *weak, *at, *t_strong, *t_at, *y, *yw;
  int B, C;
  long long n_s;
  int s0, s1, w0, w1;
};
#define T4S_SEDLOSS_PARTS 128
__device__ __forceinline__ float bce_term(float p, float y) { return -(y * fmaxf(logf(p), -100.f) + (1.0f - y) * fmaxf(logf(1.0f - p), -100.f)); }

__global__ void __launch_bounds__(256) sed_losses_partial_kernel(LossArgs a, float* __restrict__ part) {
  __shared__ float s_red[32];
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long total = (long long)a.B * a.n_s;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int b = (int)(i / a.n_s);
    const float p = a.strong[i];
    if (b >= a.s0 && b < a.s1) acc[0] += bce_term(p, a.y[i]);
    const float d = p - a.t_strong[i];
    acc[3] = fmaf(d, d, acc[3]);
  }
  const int small = a.B * a.C;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < small; i += gridDim.x * 256) {
    const int b = i / a.C;
    const float pw = a.weak[i], pa = a.at[i], ta = a.t_at[i];
    if (b >= a.w0 && b < a.w1) {
      const float yw = a.yw[i];
      acc[1] += bce_term(pw, yw);
      acc[2] += bce_term(pa, yw);
    }
    acc[4] = fmaf(pw - ta, pw - ta, acc[4]);
    acc[5] = fmaf(pa - ta, pa - ta, acc[5]);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float v = block_sum256(acc[k], s_red);
    if (threadIdx.x == 0) part[k * T4S_SEDLOSS_PARTS + blockIdx.x] = v;
  }
}
// out[0..5] = the six means, out[6] = total = l0 + w_weak l1 + w_at l2 + w_cons (l3 + w_weak_cons l4 + w_at l5)
__global__ void sed_losses_finish_kernel(const float* __restrict__ part, int nparts, LossArgs a, float w_weak, float w_at, float w_cons, float w_weak_cons,
                                         float* __restrict__ out) {
  float v[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float s = 0.f;
    for (int i = threadIdx.x; i < nparts; i += 32) s += part[k * T4S_SEDLOSS_PARTS + i];
    v[k] = warp_sum(s);
  }
  if (threadIdx.x == 0) {
    const float n_strong = (float)(a.s1 - a.s0) * (float)a.n_s, n_weak = (float)(a.w1 - a.w0) * (float)a.C;
    v[0] /= n_strong;
    v[1] /= n_weak;
    v[2] /= n_weak;
    v[3] /= (float)a.B * (float)a.n_s;
    v[4] /= (float)a.B * (float)a.C;
    v[5] /= (float)a.B * (float)a.C;
#pragma unroll
    for (int k = 0; k < 6; ++k) out[k] = v[k];
    out[6] = v[0] + w_weak * v[1] + w_at * v[2] + w_cons * (v[3] + w_weak_cons * v[4] + w_at * v[5]);
  }
}
__device__ __forceinline__ float bce_grad(float p, float y) { return (p - y) / fmaxf(p * (1.0f - p), 1e-12f); }
// d total / d strong, d weak, d at (the teacher side is detached upstream), scaled by gout[0]
__global__ void sed_losses_bwd_kernel(LossArgs a, float w_weak, float w_at, float w_cons, float w_weak_cons, const float* __restrict__ gout,
                                      float* __restrict__ d_strong, float* __restrict__ d_weak, float* __restrict__ d_at) {
  const float g = gout[0];
  const float c_bs = g / ((float)(a.s1 - a.s0) * (float)a.n_s), c_ms = g * w_cons * 2.f / ((float)a.B * (float)a.n_s);
  const long long total = (long long)a.B * a.n_s;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int b = (int)(i / a.n_s);
    const float p = a.strong[i];
    float d = c_ms * (p - a.t_strong[i]);
    if (b >= a.s0 && b < a.s1) d += c_bs * bce_grad(p, a.y[i]);
    d_strong[i] = d;
  }
  const float n_weak = (float)(a.w1 - a.w0) * (float)a.C, n_all = (float)a.B * (float)a.C;
  const int small = a.B * a.C;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < small; i += gridDim.x * 256) {
    const int b = i / a.C;
    const float pw = a.weak[i], pa = a.at[i], ta = a.t_at[i];
    float dw = g * w_cons * w_weak_cons * 2.f * (pw - ta) / n_all;
    float da = g * w_cons * w_at * 2.f * (pa - ta) / n_all;
    if (b >= a.w0 && b < a.w1) {
      const float yw = a.yw[i];
      dw += g * w_weak * bce_grad(pw, yw) / n_weak;
      da += g * w_at * bce_grad(pa, yw) / n_weak;
    }
    d_weak[i] = dw;
    d_at[i] = da;
  }
}

static int grid_for(long long n, int threads = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + threads - 1) / threads, (long long)sm_count() * 8));
}

}  // namespace post
}  // namespace t4s

extern "C" {

int t4s_scaler_instance(const float* x, float* out, int batch, int64_t inner, int mode, float eps, void* stream) {
  T4S_REQUIRE(x && out && batch > 0 && inner > 0 && mode >= 0 && mode <= 2, "t4s_scaler_instance: bad arguments");
  T4S_REQUIRE(mode != 1 || inner > 1, "t4s_scaler_instance: 'standard' needs more than one element per instance");
  t4s::post::scaler_instance_kernel<<<batch, 256, 0, t4s::as_stream(stream)>>>(x, out, inner, mode, eps);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_scaler_dataset(const float* x, float* out, int64_t total, const float* mean_dev, const float* mean_sq_dev, int standard, float eps, void* stream) {
  T4S_REQUIRE(x && out && total > 0 && mean_dev && (!standard || mean_sq_dev), "t4s_scaler_dataset: bad arguments");
  t4s::post::scaler_affine_kernel<<<t4s::post::grid_for(total), 256, 0, t4s::as_stream(stream)>>>(x, out, total, mean_dev, mean_sq_dev, standard, eps);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_rank_filter(const float* in, float* out, const int* window_sizes, int batch, int length, int classes, int op, void* stream) {
  T4S_REQUIRE(in && out && window_sizes && batch > 0 && length > 0 && classes > 0 && classes <= T4S_MEDIAN_MAX_CLASSES && in != out && (op == 0 || op == 1),
              "t4s_rank_filter: bad arguments (classes <= %d, op 0 = median, 1 = max)", T4S_MEDIAN_MAX_CLASSES);
  t4s::post::Sizes s;
  for (int c = 0; c < classes; ++c) {
    T4S_REQUIRE(window_sizes[c] >= 1 && window_sizes[c] <= T4S_MEDIAN_MAX_WINDOW, "t4s_rank_filter: window %d of class %d must be in [1, %d]", window_sizes[c],
                c, T4S_MEDIAN_MAX_WINDOW);
    s.v[c] = window_sizes[c];
  }
  dim3 grid((length + 255) / 256, classes, batch);
  t4s::post::rank_filter_kernel<<<grid, 256, 0, t4s::as_stream(stream)>>>(in, out, s, batch, length, classes, op);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_event_sweep(const float* scores, const float* weak, const float* thresholds_dev, int n_thresholds, int batch, int length, int classes, int* counts,
                    const int64_t* offsets, int* events, void* stream) {
  T4S_REQUIRE(scores && thresholds_dev && n_thresholds > 0 && batch > 0 && length > 0 && classes > 0, "t4s_event_sweep: bad arguments");
  T4S_REQUIRE((events == nullptr) ? counts != nullptr : offsets != nullptr, "t4s_event_sweep: pass 1 needs counts, pass 2 needs offsets + events");
  const long long total = (long long)n_thresholds * batch * classes;
  t4s::post::event_sweep_kernel<<<t4s::post::grid_for(total, 128), 128, 0, t4s::as_stream(stream)>>>(scores, weak, thresholds_dev, n_thresholds, batch, length,
                                                                                                 classes, counts, (const long long*)offsets, events);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

static int fill_loss_args(t4s::post::LossArgs& a, const T4sSedLosses* p) {
  T4S_REQUIRE(p && p->strong && p->weak && p->at && p->t_strong && p->t_at && p->y && p->yw, "t4s_sed_losses: null pointer");
  T4S_REQUIRE(p->batch > 0 && p->classes > 0 && p->strong_inner > 0, "t4s_sed_losses: bad sizes");
  T4S_REQUIRE(0 <= p->s0 && p->s0 < p->s1 && p->s1 <= p->batch && 0 <= p->w0 && p->w0 < p->w1 && p->w1 <= p->batch,
              "t4s_sed_losses: the strong / weak row ranges must be non-empty and inside the batch");
  a.strong = p->strong; a.weak = p->weak; a.at = p->at; a.t_strong = p->t_strong; a.t_at = p->t_at; a.y = p->y; a.yw = p->yw;
  a.B = p->batch; a.C = p->classes; a.n_s = p->strong_inner;
  a.s0 = p->s0; a.s1 = p->s1; a.w0 = p->w0; a.w1 = p->w1;
  return T4S_OK;
}

int t4s_sed_losses_fwd(const T4sSedLosses* p, float* ws, float* out, void* stream) {
  t4s::post::LossArgs a;
  int rc = fill_loss_args(a, p);
  if (rc) return rc;
  T4S_REQUIRE(ws && out, "t4s_sed_losses_fwd: ws (6 x 128 floats) and out (7 floats) are required");
  cudaStream_t st = t4s::as_stream(stream);
  const long long total = (long long)a.B * a.n_s;
  const int parts = (int)std::min<long long>(T4S_SEDLOSS_PARTS, (total + 255) / 256);
  t4s::post::sed_losses_partial_kernel<<<parts, 256, 0, st>>>(a, ws);
  T4S_LAUNCH_CHECK();
  t4s::post::sed_losses_finish_kernel<<<1, 32, 0, st>>>(ws, parts, a, p->w_weak, p->w_at, p->w_cons, p->w_weak_cons, out);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_sed_losses_bwd(const T4sSedLosses* p, const float* grad_total, float* d_strong, float* d_weak, float* d_at, void* stream) {
  t4s::post::LossArgs a;
  int rc = fill_loss_args(a, p);
  if (rc) return rc;
  T4S_REQUIRE(grad_total && d_strong && d_weak && d_at, "t4s_sed_losses_bwd: null pointer");
  const long long total = (long long)a.B * a.n_s;
  t4s::post::sed_losses_bwd_kernel<<<t4s::post::grid_for(total), 256, 0, t4s::as_stream(stream)>>>(a, p->w_weak, p->w_at, p->w_cons, p->w_weak_cons, grad_total,
                                                                                               d_strong, d_weak, d_at);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
