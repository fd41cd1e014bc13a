// K1': the front end for any power-of-two n_fft other than the PaSST 1024 (256 ... 4096): the DCASE-style 16 kHz parametrisation of
// reference src/preprocess/feats_extraction.py:41-57 (setmelspectrogram + take_log: hamming window = n_fft = 2048, hop 256,
// magnitude spectrum, HTK mel, 20 log10 clamp) -- dead code upstream, SURVEY §8 a1'.
//
// One persistent CTA (256 threads) walks tiles of 8 consecutive frames of one clip.  Per frame: the n_fft windowed samples are
// gathered straight from the waveform (reflect padding / pre-emphasis / peak scale are index arithmetic), a radix-2 Stockham
// FFT runs in shared memory (log2(n_fft) barrier-separated stages, twiddles staged once per CTA), then |X| or |X|^2 and the
// sparse mel rows.  The [n_mels x 8] tile leaves with the log fused, 32 contiguous bytes per mel row.  HBM traffic = wav once
// (frame overlap is served by L1/L2) + mel once.
#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace melg {

constexpr int kThreads = 256;
constexpr int kTile = 8;
constexpr int kMaxMels = 128;

struct Params {
  const float* wav;
  const float* peak;
  const float2* tw;      // W_N^k, k < N/2
  const float* window;   // n_fft floats (win_length window centred, zero outside)
  const int* bin_start;
  const int* bin_count;
  const int* w_offset;
  const float* weights;
  int n_weights;
  void* out;
  int batch, n_samples, n_frames, tiles_per_clip;
  int n_fft, log2n, hop, n_mels, preemphasis, wav_norm, magnitude, out_mode;
};

__global__ void tables_kernel(float2* tw, int n_fft) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_fft / 2) {
    float s, c;
    sincospif(-2.0f * (float)k / (float)n_fft, &s, &c);
    tw[k] = make_float2(c, s);
  }
}

__device__ __forceinline__ float sample(const Params& p, const float* w, int n, int Ly, float scale) {
  // reflect padding of the (pre-emphasised) signal of length Ly
  if (n < 0) n = -n;
  if (n >= Ly) n = 2 * (Ly - 1) - n;
  n = max(0, min(n, Ly - 1));
  const float v = p.preemphasis ? (w[n + 1] - 0.97f * w[n]) : w[n];
  return v * scale;
}

template <typename OutT>
__global__ void __launch_bounds__(kThreads) mel_generic_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.n_fft, H = N / 2;
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);          // H
  float2* s_a = s_tw + H;                                      // N
  float2* s_b = s_a + N;                                       // N
  float* s_win = reinterpret_cast<float*>(s_b + N);            // N
  float* s_spec = s_win + N;                                   // H + 1 (padded to H + 4)
  float* s_w = s_spec + H + 4;                                 // n_weights (padded to 4)
  float* s_tile = s_w + ((p.n_weights + 3) & ~3);              // n_mels * (kTile + 1)
  int* s_idx = reinterpret_cast<int*>(s_tile + kMaxMels * (kTile + 1));  // 3 * n_mels
  for (int i = threadIdx.x; i < H; i += kThreads) s_tw[i] = p.tw[i];
  for (int i = threadIdx.x; i < N; i += kThreads) s_win[i] = p.window[i];
  for (int i = threadIdx.x; i < p.n_weights; i += kThreads) s_w[i] = p.weights[i];
  for (int i = threadIdx.x; i < p.n_mels; i += kThreads) {
    s_idx[i] = p.bin_start[i];
    s_idx[p.n_mels + i] = p.bin_count[i];
    s_idx[2 * p.n_mels + i] = p.w_offset[i];
  }
  __syncthreads();
  const int Ly = p.preemphasis ? p.n_samples - 1 : p.n_samples;
  const long total_tiles = (long)p.batch * p.tiles_per_clip;
  for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = (int)(tile / p.tiles_per_clip), t0 = (int)(tile % p.tiles_per_clip) * kTile;
    const float* w = p.wav + (long)b * p.n_samples;
    const float scale = p.wav_norm ? 1.0f / (p.peak[b] + 1e-10f) : 1.0f;
    const int nf = min(kTile, p.n_frames - t0);
    for (int f = 0; f < nf; ++f) {
      const int start = (t0 + f) * p.hop - H;
      for (int i = threadIdx.x; i < N; i += kThreads) s_a[i] = make_float2(sample(p, w, start + i, Ly, scale) * s_win[i], 0.f);
      __syncthreads();
      float2 *src = s_a, *dst = s_b;
      for (int ns = 1, sh = p.log2n - 1; ns < N; ns <<= 1, --sh) {   // Stockham radix-2: natural-order output
        for (int j = threadIdx.x; j < H; j += kThreads) {
          const int k = j & (ns - 1);
          const float2 tw = s_tw[k << sh];
          const float2 a = src[j], bb = src[j + H];
          const float2 bw = make_float2(bb.x * tw.x - bb.y * tw.y, bb.x * tw.y + bb.y * tw.x);
          const int o = ((j - k) << 1) + k;
          dst[o] = make_float2(a.x + bw.x, a.y + bw.y);
          dst[o + ns] = make_float2(a.x - bw.x, a.y - bw.y);
        }
        __syncthreads();
        float2* tmp = src; src = dst; dst = tmp;
      }
      for (int i = threadIdx.x; i <= H; i += kThreads) {
        const float2 v = src[i];
        const float pw = v.x * v.x + v.y * v.y;
        s_spec[i] = p.magnitude ? sqrtf(pw) : pw;
      }
      __syncthreads();
      for (int m = threadIdx.x; m < p.n_mels; m += kThreads) {
        const int bs = s_idx[m], bc = s_idx[p.n_mels + m], wo = s_idx[2 * p.n_mels + m];
        float acc = 0.f;
        for (int i = 0; i < bc; ++i) acc += s_w[wo + i] * s_spec[bs + i];
        float o = acc;
        if (p.out_mode == 1) o = (logf(acc + 1e-5f) + 4.5f) / 5.0f;
        else if (p.out_mode == 2) o = fminf(fmaxf((p.magnitude ? 20.0f : 10.0f) * log10f(fmaxf(acc, 1e-5f)), -50.0f), 80.0f);
        s_tile[m * (kTile + 1) + f] = o;
      }
      __syncthreads();
    }
    OutT* out = static_cast<OutT*>(p.out) + (long)b * p.n_mels * p.n_frames + t0;
    for (int i = threadIdx.x; i < p.n_mels * nf; i += kThreads) {
      const int m = i / nf, f = i - m * nf;
      out[(long)m * p.n_frames + f] = from_f32<OutT>(s_tile[m * (kTile + 1) + f]);
    }
    __syncthreads();
  }
}

__global__ void amp_to_db_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, float multiplier, float amin, float lo, float hi) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = fminf(fmaxf(multiplier * log10f(fmaxf(in[i], amin)), lo), hi);
}

size_t smem_bytes(int n_fft, int n_weights) {
  return (size_t)(n_fft / 2 + 2 * n_fft) * 8 + (size_t)(n_fft + n_fft / 2 + 4 + ((n_weights + 3) & ~3) + kMaxMels * (kTile + 1) + 3 * kMaxMels) * 4;
}

bool supported(int n_fft) { return n_fft >= 256 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0; }

size_t tables_bytes(int n_fft) { return (size_t)(n_fft / 2) * 8 + (size_t)n_fft * 4; }

int tables_init(void* tables, const float* window_host, int n_fft, int win_length, cudaStream_t st) {
  float2* tw = reinterpret_cast<float2*>(tables);
  float* win = reinterpret_cast<float*>(tw + n_fft / 2);
  tables_kernel<<<(n_fft / 2 + 255) / 256, 256, 0, st>>>(tw, n_fft);
  T4S_LAUNCH_CHECK();
  T4S_CUDA(cudaMemsetAsync(win, 0, sizeof(float) * n_fft, st));
  const int left = (n_fft - win_length) / 2;   // torch.stft centres a short window inside the frame
  T4S_CUDA(cudaMemcpyAsync(win + left, window_host, sizeof(float) * win_length, cudaMemcpyHostToDevice, st));
  T4S_CUDA(cudaStreamSynchronize(st));
  return T4S_OK;
}

int forward(const float* wav, const float* peak, const void* tables, const int* bin_start, const int* bin_count, const int* w_offset,
            const float* weights, int n_weights, void* out, int batch, int n_samples, int n_frames, const T4sMelParams* mp, cudaStream_t st) {
  T4S_REQUIRE(mp->n_mels > 0 && mp->n_mels <= kMaxMels, "t4s_mel_forward: n_mels must be in 1..%d", kMaxMels);
  T4S_REQUIRE(mp->hop > 0 && n_weights > 0, "t4s_mel_forward: bad hop / basis");
  const int Ly = mp->preemphasis ? n_samples - 1 : n_samples;
  T4S_REQUIRE(batch > 0 && Ly > mp->n_fft / 2, "t4s_mel_forward: clip too short for reflect padding (need > %d samples)", mp->n_fft / 2 + 1);
  T4S_REQUIRE(n_frames == 1 + Ly / mp->hop, "t4s_mel_forward: n_frames must be 1 + %d / hop", Ly);
  T4S_REQUIRE(!mp->wav_norm || peak, "t4s_mel_forward: wav_norm needs the peak buffer");
  Params p;
  p.wav = wav; p.peak = peak;
  p.tw = reinterpret_cast<const float2*>(tables);
  p.window = reinterpret_cast<const float*>(p.tw + mp->n_fft / 2);
  p.bin_start = bin_start; p.bin_count = bin_count; p.w_offset = w_offset; p.weights = weights; p.n_weights = n_weights;
  p.out = out; p.batch = batch; p.n_samples = n_samples; p.n_frames = n_frames;
  p.tiles_per_clip = (n_frames + kTile - 1) / kTile;
  p.n_fft = mp->n_fft; p.log2n = 0;
  while ((1 << p.log2n) < mp->n_fft) ++p.log2n;
  p.hop = mp->hop; p.n_mels = mp->n_mels; p.preemphasis = mp->preemphasis; p.wav_norm = mp->wav_norm; p.magnitude = mp->magnitude;
  p.out_mode = mp->out_mode;
  const size_t smem = smem_bytes(mp->n_fft, n_weights);
  T4S_REQUIRE(smem <= 227 * 1024, "t4s_mel_forward: n_fft=%d needs %zu B shared memory", mp->n_fft, smem);
  const long total_tiles = (long)batch * p.tiles_per_clip;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / smem));
  const int grid = (int)std::min<long>(total_tiles, (long)per_sm * sm_count());
  if (mp->out_dtype == T4S_F32) {
    T4S_CUDA(cudaFuncSetAttribute(mel_generic_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mel_generic_kernel<float><<<grid, kThreads, smem, st>>>(p);
  } else if (mp->out_dtype == T4S_BF16) {
    T4S_CUDA(cudaFuncSetAttribute(mel_generic_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mel_generic_kernel<__nv_bfloat16><<<grid, kThreads, smem, st>>>(p);
  } else {
    set_error("t4s_mel_forward: bad out_dtype %d", mp->out_dtype);
    return T4S_ERR_ARG;
  }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // namespace melg
}  // namespace t4s

/* take_log (src/preprocess/feats_extraction.py:41-44): clamp(multiplier * log10(max(x, amin)), lo, hi) */
extern "C" int t4s_amp_to_db(const float* in, float* out, size_t n, float multiplier, float amin, float lo, float hi, void* stream) {
  T4S_REQUIRE(in && out, "t4s_amp_to_db: null pointer");
  if (n == 0) return T4S_OK;
  const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)t4s::sm_count() * 8);
  t4s::melg::amp_to_db_kernel<<<grid, 256, 0, t4s::as_stream(stream)>>>(in, out, n, multiplier, amin, lo, hi);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}
