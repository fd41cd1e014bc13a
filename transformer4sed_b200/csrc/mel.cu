// K1: fused wav -> (peak-norm, pre-emphasis, reflect pad) -> windowed 1024-point rFFT -> power -> sparse mel -> log.
//
// Replaces reference src/models/passt/passt_feature_extraction.py:46-94.  One persistent CTA (8 warps) walks tiles of
// 16 consecutive frames of one clip:
//   1. all threads stage the tile's pre-emphasised samples (5.6 K floats) in shared memory with coalesced loads
//      (reflect padding is index arithmetic, the padded signal is never materialised);
//   2. each warp owns a frame: 512-point complex FFT of the even/odd-packed windowed frame as three radix-8 passes
//      held in registers (2 butterflies per lane per pass) with warp-private shared-memory exchanges (no CTA barrier),
//      real-FFT split, |X|^2, then the <=32-tap rows of the mel basis;
//   3. the [n_mels x 16] tile is written with the log fused, 64 B contiguous per mel row.
// Window, twiddles and the sparse basis are staged in shared memory once per CTA.
// HBM traffic = wav once (frame overlap is served by L1/L2) + mel once: 1.79 MB/clip algorithmic.
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"

namespace t4s {
namespace mel {

constexpr int kNfft = 1024;
constexpr int kHalf = 512;           // complex FFT length
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTileFrames = 16;
constexpr int kMaxMels = 128;
constexpr int kMaxWeights = 4096;
constexpr int kBufStride1 = 72;      // pass-1 output layout [q0][72]
constexpr int kBufStride2 = 66;      // pass-2 output layout [c][66]
constexpr int kBufLen = 8 * kBufStride1;  // float2 per warp
constexpr int kPowLen = 520;         // floats per warp (513 used)

// Table layout (float2 units unless noted), built by tables_init_kernel:
//   tw1[7][64]  : W64^(b*q0),  q0 = 1..7, butterfly u = 8b + c
//   tw2[8][64]  : W512^(c*(q0+8*q1)), q1 = 0..7, butterfly u = 8*q0 + c
//   tws[516]    : W1024^k, k = 0..512
//   window[win] : floats
constexpr int kTw1Off = 0;
constexpr int kTw2Off = 7 * 64;
constexpr int kTwsOff = kTw2Off + 8 * 64;
constexpr int kWinOff = kTwsOff + 516;  // in float2 units; window floats start at 2*kWinOff

__host__ __device__ inline size_t tables_bytes(int win_length) { return (size_t)kWinOff * 8 + (size_t)((win_length + 3) & ~3) * 4; }

__global__ void tables_init_kernel(float2* tab) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 7 * 64) {
    int q0 = i / 64 + 1, u = i % 64, b = u >> 3;
    float s, c;
    sincospif(-2.0f * (float)(b * q0) / 64.0f, &s, &c);
    tab[kTw1Off + i] = make_float2(c, s);
  } else if (i < 7 * 64 + 8 * 64) {
    int j = i - 7 * 64;
    int q1 = j / 64, u = j % 64, q0 = u >> 3, c = u & 7;
    float s, co;
    sincospif(-2.0f * (float)((c * (q0 + 8 * q1)) & 511) / 512.0f, &s, &co);
    tab[kTw2Off + j] = make_float2(co, s);
  } else if (i < 7 * 64 + 8 * 64 + 516) {
    int k = i - (7 * 64 + 8 * 64);
    float s, c;
    sincospif(-2.0f * (float)k / 1024.0f, &s, &c);
    tab[kTwsOff + k] = make_float2(c, s);
  }
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// In-place forward 8-point DFT: v[q] <- sum_a v[a] exp(-2 pi i a q / 8).
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
  const float r = 0.70710678118654752440f;
  float2 e0 = cadd(v[0], v[4]), o0 = csub(v[0], v[4]);
  float2 e1 = cadd(v[1], v[5]), o1 = csub(v[1], v[5]);
  float2 e2 = cadd(v[2], v[6]), o2 = csub(v[2], v[6]);
  float2 e3 = cadd(v[3], v[7]), o3 = csub(v[3], v[7]);
  // odd outputs: twiddle o_a by W8^a
  o1 = make_float2(r * (o1.x + o1.y), r * (o1.y - o1.x));    // * (1 - i)/sqrt2
  o2 = mul_mi(o2);                                            // * (-i)
  o3 = make_float2(r * (o3.y - o3.x), -r * (o3.x + o3.y));   // * (-1 - i)/sqrt2
  // 4-point DFTs
  float2 s0 = cadd(e0, e2), d0 = csub(e0, e2), s1 = cadd(e1, e3), d1 = mul_mi(csub(e1, e3));
  v[0] = cadd(s0, s1);
  v[4] = csub(s0, s1);
  v[2] = cadd(d0, d1);
  v[6] = csub(d0, d1);
  float2 t0 = cadd(o0, o2), u0 = csub(o0, o2), t1 = cadd(o1, o3), u1 = mul_mi(csub(o1, o3));
  v[1] = cadd(t0, t1);
  v[5] = csub(t0, t1);
  v[3] = cadd(u0, u1);
  v[7] = csub(u0, u1);
}

struct Params {
  const float* wav;
  const float* peak;
  const float2* tables;
  const int* bin_start;
  const int* bin_count;
  const int* w_offset;
  const float* weights;
  int n_weights;
  void* out;
  int batch, n_samples, n_frames, tiles_per_clip;
  int win_length, hop, n_mels, preemphasis, wav_norm, magnitude, out_mode;
};

template <typename OutT>
__global__ void __launch_bounds__(kThreads, 2) mel_kernel(const Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // carve shared memory
  float2* s_tw1 = reinterpret_cast<float2*>(smem_raw);               // 7*64
  float2* s_tw2 = s_tw1 + 7 * 64;                                    // 8*64
  float2* s_tws = s_tw2 + 8 * 64;                                    // 516
  float2* s_buf = s_tws + 516;                                       // kWarps * kBufLen
  float* s_pow = reinterpret_cast<float*>(s_buf + kWarps * kBufLen); // kWarps * kPowLen
  float* s_win = s_pow + kWarps * kPowLen;                           // win_length (padded to 4)
  const int win_pad = (p.win_length + 3) & ~3;
  float* s_wts = s_win + win_pad;                                    // n_weights (padded to 4)
  const int wts_pad = (p.n_weights + 3) & ~3;
  int* s_bs = reinterpret_cast<int*>(s_wts + wts_pad);               // kMaxMels
  int* s_bc = s_bs + kMaxMels;
  int* s_wo = s_bc + kMaxMels;
  float* s_tile = reinterpret_cast<float*>(s_wo + kMaxMels);         // kMaxMels * (kTileFrames+1)
  float* s_y = s_tile + kMaxMels * (kTileFrames + 1);                // span samples
  const int span = (kTileFrames - 1) * p.hop + p.win_length;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // one-time staging of tables, window and sparse basis
  for (int i = tid; i < kWinOff; i += kThreads) s_tw1[i] = p.tables[i];
  {
    const float* gwin = reinterpret_cast<const float*>(p.tables + kWinOff);
    for (int i = tid; i < p.win_length; i += kThreads) s_win[i] = gwin[i];
  }
  for (int i = tid; i < p.n_weights; i += kThreads) s_wts[i] = p.weights[i];
  for (int i = tid; i < kMaxMels; i += kThreads) {
    bool ok = i < p.n_mels;
    s_bs[i] = ok ? p.bin_start[i] : 0;
    s_bc[i] = ok ? p.bin_count[i] : 0;
    s_wo[i] = ok ? p.w_offset[i] : 0;
  }
  __syncthreads();

  const int Ly = p.preemphasis ? p.n_samples - 1 : p.n_samples;  // length of the signal that is framed
  const int win_left = (kNfft - p.win_length) / 2;               // torch.stft centres a short window
  float2* buf = s_buf + warp * kBufLen;
  float* pw = s_pow + warp * kPowLen;
  const int total_tiles = p.batch * p.tiles_per_clip;

  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_clip;
    const int t0 = (tile - b * p.tiles_per_clip) * kTileFrames;
    const float* x = p.wav + (size_t)b * p.n_samples;
    const float denom = p.wav_norm ? (p.peak[b] + 1e-10f) : 1.0f;
    // ---- 1. stage pre-emphasised, normalised samples y[ybase .. ybase+span) with reflect indexing
    const int ybase = t0 * p.hop - kNfft / 2 + win_left;
    for (int i = tid; i < span; i += kThreads) {
      int n = ybase + i;
      if (n < 0) n = -n;
      if (n >= Ly) n = 2 * (Ly - 1) - n;
      float v = 0.f;
      if (n >= 0 && n < Ly) {
        if (p.preemphasis) {
          float x0 = __fdiv_rn(__ldg(x + n), denom), x1 = __fdiv_rn(__ldg(x + n + 1), denom);
          v = x1 - 0.97f * x0;
        } else {
          v = __fdiv_rn(__ldg(x + n), denom);
        }
      }
      s_y[i] = v;
    }
    __syncthreads();

    // ---- 2. one frame per warp iteration
    for (int f = warp; f < kTileFrames; f += kWarps) {
      const int t = t0 + f;
      if (t >= p.n_frames) break;  // warp-uniform
      const float* yf = s_y + f * p.hop;
      float2 v0[8], v1[8];
      // pass 1: butterflies u = lane, lane+32 over a; inputs z[64a+u] = (f[2k], f[2k+1]) windowed
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        int w0 = 128 * a + 2 * lane - win_left, w1 = w0 + 64;  // window index of the pair's first element
        float2 z0 = make_float2(0.f, 0.f), z1 = z0;
        if (w0 >= 0 && w0 + 1 < p.win_length) {
          float2 yy = *reinterpret_cast<const float2*>(yf + w0), ww = *reinterpret_cast<const float2*>(s_win + w0);
          z0 = make_float2(yy.x * ww.x, yy.y * ww.y);
        } else if (w0 >= 0 && w0 < p.win_length) {
          z0.x = yf[w0] * s_win[w0];
        }
        if (w1 >= 0 && w1 + 1 < p.win_length) {
          float2 yy = *reinterpret_cast<const float2*>(yf + w1), ww = *reinterpret_cast<const float2*>(s_win + w1);
          z1 = make_float2(yy.x * ww.x, yy.y * ww.y);
        } else if (w1 >= 0 && w1 < p.win_length) {
          z1.x = yf[w1] * s_win[w1];
        }
        v0[a] = z0;
        v1[a] = z1;
      }
      fft8(v0);
      fft8(v1);
      buf[lane] = v0[0];
      buf[lane + 32] = v1[0];
#pragma unroll
      for (int q = 1; q < 8; ++q) {
        buf[q * kBufStride1 + lane] = cmul(v0[q], s_tw1[(q - 1) * 64 + lane]);
        buf[q * kBufStride1 + lane + 32] = cmul(v1[q], s_tw1[(q - 1) * 64 + lane + 32]);
      }
      __syncwarp();
      // pass 2: butterfly u = 8*q0 + c over b
      {
        const int u0 = lane, u1 = lane + 32;
        const int q00 = u0 >> 3, c0 = u0 & 7, q01 = u1 >> 3, c1 = u1 & 7;
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
          v0[bb] = buf[q00 * kBufStride1 + 8 * bb + c0];
          v1[bb] = buf[q01 * kBufStride1 + 8 * bb + c1];
        }
        __syncwarp();
        fft8(v0);
        fft8(v1);
#pragma unroll
        for (int q1 = 0; q1 < 8; ++q1) {
          buf[c0 * kBufStride2 + q1 * 8 + q00] = cmul(v0[q1], s_tw2[q1 * 64 + u0]);
          buf[c1 * kBufStride2 + q1 * 8 + q01] = cmul(v1[q1], s_tw2[q1 * 64 + u1]);
        }
      }
      __syncwarp();
      // pass 3: butterfly r = 8*q1 + q0 over c; output Z[r + 64*q2]
      {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          v0[c] = buf[c * kBufStride2 + lane];
          v1[c] = buf[c * kBufStride2 + lane + 32];
        }
        __syncwarp();
        fft8(v0);
        fft8(v1);
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) {
          buf[64 * q2 + lane] = v0[q2];
          buf[64 * q2 + lane + 32] = v1[q2];
        }
      }
      __syncwarp();
      // real-FFT split + power: X[k] = E + W1024^k * O, E = (Z[k]+conj Z[512-k])/2, O = (Z[k]-conj Z[512-k])/(2i)
#pragma unroll 4
      for (int k = lane; k < kHalf; k += 32) {
        float2 zk = buf[k], zn = buf[(kHalf - k) & (kHalf - 1)];
        float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
        float2 xk = cadd(e, cmul(s_tws[k], o));
        float pwr = xk.x * xk.x + xk.y * xk.y;
        pw[k] = p.magnitude ? sqrtf(pwr) : pwr;
      }
      if (lane == 0) {
        float2 z0 = buf[0];
        float xn = z0.x - z0.y;  // Nyquist bin
        pw[kHalf] = p.magnitude ? fabsf(xn) : xn * xn;
      }
      __syncwarp();
      // sparse mel rows
      for (int m = lane; m < p.n_mels; m += 32) {
        const int bs = s_bs[m], bc = s_bc[m], wo = s_wo[m];
        float acc = 0.f;
        for (int i = 0; i < bc; ++i) acc = fmaf(s_wts[wo + i], pw[bs + i], acc);
        s_tile[m * (kTileFrames + 1) + f] = acc;
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- 3. coalesced tile store with the log fused
    OutT* out = reinterpret_cast<OutT*>(p.out) + (size_t)b * p.n_mels * p.n_frames;
    for (int i = tid; i < p.n_mels * kTileFrames; i += kThreads) {
      int m = i / kTileFrames, f = i % kTileFrames, t = t0 + f;
      if (t < p.n_frames) {
        float v = s_tile[m * (kTileFrames + 1) + f];
        if (p.out_mode == 1) v = (logf(v + 1e-5f) + 4.5f) / 5.0f;
        else if (p.out_mode == 2) v = fminf(fmaxf(20.0f * log10f(fmaxf(v, 1e-5f)), -50.0f), 80.0f);
        out[(size_t)m * p.n_frames + t] = from_f32<OutT>(v);
      }
    }
    // s_tile / s_y are rewritten only after the next tile's staging barrier
    __syncthreads();
  }
}

__global__ void peak_kernel(const float* __restrict__ wav, float* __restrict__ peak, int n_samples, int chunk) {
  const int b = blockIdx.y;
  const float* x = wav + (size_t)b * n_samples;
  int lo = blockIdx.x * chunk, hi = min(n_samples, lo + chunk);
  float m = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) m = fmaxf(m, fabsf(__ldg(x + i)));
  m = warp_max(m);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(peak + b), __float_as_int(m));  // non-negative floats order as ints
  }
}

__global__ void normalize_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (logf(in[i] + 1e-5f) + 4.5f) / 5.0f;
}

static size_t smem_bytes(const T4sMelParams& mp, int n_weights) {
  size_t fl2 = 7 * 64 + 8 * 64 + 516 + (size_t)kWarps * kBufLen;
  size_t fl = (size_t)kWarps * kPowLen + ((mp.win_length + 3) & ~3) + ((n_weights + 3) & ~3) + 3 * kMaxMels +
              (size_t)kMaxMels * (kTileFrames + 1) + (size_t)(kTileFrames - 1) * mp.hop + mp.win_length + 4;
  return fl2 * 8 + fl * 4;
}

}  // namespace mel
namespace melg {   // csrc/mel_generic.cu: every other power-of-two n_fft
bool supported(int n_fft);
size_t tables_bytes(int n_fft);
int tables_init(void* tables, const float* window_host, int n_fft, int win_length, cudaStream_t st);
int forward(const float* wav, const float* peak, const void* tables, const int* bin_start, const int* bin_count, const int* w_offset,
            const float* weights, int n_weights, void* out, int batch, int n_samples, int n_frames, const T4sMelParams* mp, cudaStream_t st);
}  // namespace melg
}  // namespace t4s

extern "C" {

int t4s_wav_peak(const float* wav, float* peak, int batch, int n_samples, void* stream) {
  T4S_REQUIRE(wav && peak && batch > 0 && n_samples > 0, "t4s_wav_peak: bad arguments");
  cudaStream_t st = t4s::as_stream(stream);
  T4S_CUDA(cudaMemsetAsync(peak, 0, sizeof(float) * batch, st));
  const int chunk = 16384;
  dim3 grid((n_samples + chunk - 1) / chunk, batch);
  t4s::mel::peak_kernel<<<grid, 256, 0, st>>>(wav, peak, n_samples, chunk);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

size_t t4s_mel_tables_bytes(int n_fft, int win_length) {
  if (win_length <= 0 || win_length > n_fft) return 0;
  if (n_fft != t4s::mel::kNfft) return t4s::melg::supported(n_fft) ? t4s::melg::tables_bytes(n_fft) : 0;
  return t4s::mel::tables_bytes(win_length);
}

int t4s_mel_tables_init(void* tables, const float* window_host, int n_fft, int win_length, void* stream) {
  T4S_REQUIRE(tables && window_host, "t4s_mel_tables_init: null pointer");
  if (n_fft != t4s::mel::kNfft) {
    if (!t4s::melg::supported(n_fft) || win_length <= 0 || win_length > n_fft) {
      t4s::set_error("t4s_mel_tables_init: n_fft=%d unsupported (powers of two in 256..4096)", n_fft);
      return T4S_ERR_UNSUPPORTED;
    }
    return t4s::melg::tables_init(tables, window_host, n_fft, win_length, t4s::as_stream(stream));
  }
  T4S_REQUIRE(win_length > 0 && win_length <= n_fft && (win_length % 4) == 0, "t4s_mel_tables_init: win_length must be a multiple of 4 and <= n_fft");
  cudaStream_t st = t4s::as_stream(stream);
  const int n = 7 * 64 + 8 * 64 + 516;
  t4s::mel::tables_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<float2*>(tables));
  T4S_LAUNCH_CHECK();
  T4S_CUDA(cudaMemcpyAsync(reinterpret_cast<float2*>(tables) + t4s::mel::kWinOff, window_host, sizeof(float) * win_length,
                           cudaMemcpyHostToDevice, st));
  T4S_CUDA(cudaStreamSynchronize(st));  // window_host may be a temporary
  return T4S_OK;
}

int t4s_mel_forward(const float* wav, const float* peak, const void* tables, const int* bin_start, const int* bin_count,
                    const int* w_offset, const float* weights, int n_weights, void* out, int batch, int n_samples,
                    int n_frames, const T4sMelParams* mp, void* stream) {
  using namespace t4s::mel;
  T4S_REQUIRE(wav && tables && bin_start && bin_count && w_offset && weights && out && mp, "t4s_mel_forward: null pointer");
  if (mp->n_fft != kNfft) {
    if (!t4s::melg::supported(mp->n_fft)) {
      t4s::set_error("t4s_mel_forward: n_fft=%d unsupported (powers of two in 256..4096)", mp->n_fft);
      return T4S_ERR_UNSUPPORTED;
    }
    return t4s::melg::forward(wav, peak, tables, bin_start, bin_count, w_offset, weights, n_weights, out, batch, n_samples, n_frames, mp,
                              t4s::as_stream(stream));
  }
  T4S_REQUIRE(mp->win_length > 0 && mp->win_length <= kNfft && mp->win_length % 4 == 0, "t4s_mel_forward: win_length must be a multiple of 4 and <= n_fft");
  T4S_REQUIRE(mp->hop > 0 && mp->hop % 2 == 0, "t4s_mel_forward: hop must be positive and even");
  T4S_REQUIRE(mp->n_mels > 0 && mp->n_mels <= kMaxMels, "t4s_mel_forward: n_mels must be in 1..%d", kMaxMels);
  T4S_REQUIRE(n_weights > 0 && n_weights <= kMaxWeights, "t4s_mel_forward: n_weights must be in 1..%d", kMaxWeights);
  T4S_REQUIRE(!mp->wav_norm || peak, "t4s_mel_forward: wav_norm needs the peak buffer");
  const int Ly = mp->preemphasis ? n_samples - 1 : n_samples;
  T4S_REQUIRE(batch > 0 && Ly > kNfft / 2, "t4s_mel_forward: clip too short for reflect padding (need > %d samples)", kNfft / 2 + 1);
  T4S_REQUIRE(n_frames == 1 + Ly / mp->hop, "t4s_mel_forward: n_frames must be 1 + %d / hop", Ly);
  Params p;
  p.wav = wav; p.peak = peak; p.tables = reinterpret_cast<const float2*>(tables);
  p.bin_start = bin_start; p.bin_count = bin_count; p.w_offset = w_offset; p.weights = weights; p.n_weights = n_weights;
  p.out = out; p.batch = batch; p.n_samples = n_samples; p.n_frames = n_frames;
  p.tiles_per_clip = (n_frames + kTileFrames - 1) / kTileFrames;
  p.win_length = mp->win_length; p.hop = mp->hop; p.n_mels = mp->n_mels; p.preemphasis = mp->preemphasis;
  p.wav_norm = mp->wav_norm; p.magnitude = mp->magnitude; p.out_mode = mp->out_mode;
  const size_t smem = smem_bytes(*mp, n_weights);
  T4S_REQUIRE(smem <= 227 * 1024, "t4s_mel_forward: hop/win_length need %zu B shared memory", smem);
  cudaStream_t st = t4s::as_stream(stream);
  const long total_tiles = (long)batch * p.tiles_per_clip;
  const int grid = (int)std::min<long>(total_tiles, 2L * t4s::sm_count());
  if (mp->out_dtype == T4S_F32) {
    T4S_CUDA(cudaFuncSetAttribute(mel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mel_kernel<float><<<grid, kThreads, smem, st>>>(p);
  } else if (mp->out_dtype == T4S_BF16) {
    T4S_CUDA(cudaFuncSetAttribute(mel_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mel_kernel<__nv_bfloat16><<<grid, kThreads, smem, st>>>(p);
  } else {
    t4s::set_error("t4s_mel_forward: bad out_dtype %d", mp->out_dtype);
    return T4S_ERR_ARG;
  }
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

int t4s_mel_normalize(const float* in, float* out, size_t n, void* stream) {
  T4S_REQUIRE(in && out, "t4s_mel_normalize: null pointer");
  if (n == 0) return T4S_OK;
  int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)t4s::sm_count() * 8);
  t4s::mel::normalize_kernel<<<grid, 256, 0, t4s::as_stream(stream)>>>(in, out, n);
  T4S_LAUNCH_CHECK();
  return T4S_OK;
}

}  // extern "C"
