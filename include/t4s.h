/*
 * t4s.h — C ABI of libt4s.so: hand-written sm_100a kernels for the Transformer4SED frame-level SED hot path.
 *
 * Boundary rules:
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (including workspaces); the library never allocates device memory;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on that stream and re-entrant;
 *   - return value 0 = ok, negative = error; t4s_last_error() returns a thread-local message.
 *
 * The reference (cai525/Transformer4SED) has no FFI layer: its hot path is PyTorch module code.  Each entry point
 * below cites the reference lines whose arithmetic it replaces (paths relative to the reference root).
 */
#ifndef T4S_H_
#define T4S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T4S_OK 0
#define T4S_ERR_ARG (-1)
#define T4S_ERR_CUDA (-2)
#define T4S_ERR_DEVICE (-3)
#define T4S_ERR_UNSUPPORTED (-4)

/* dtype tags used by the type-generic entry points */
#define T4S_F32 0
#define T4S_BF16 1

/* ---- library ------------------------------------------------------------------------------------------- */
int t4s_version(void);
const char* t4s_last_error(void);
/* 0 iff the current device is compute capability 10.x (sm_100a SASS is loadable). */
int t4s_device_check(void);
int t4s_sm_count(void);

/* ---- K1: fused STFT -> power -> mel -> log front end ----------------------------------------------------
 * Replaces src/models/passt/passt_feature_extraction.py:46-94 (PasstFeatureExtractor.forward + .normalize):
 * peak-normalise, pre-emphasis [-.97, 1], reflect-pad n_fft/2, frames of n_fft with a win_length window centred
 * in the frame, rFFT power, sparse mel basis, optional (ln(x+1e-5)+4.5)/5.
 * and, with T4S_MEL_DCASE flags, src/preprocess/feats_extraction.py:41-57 (setmelspectrogram + take_log).
 */
typedef struct {
  int n_fft;        /* 1024 or 2048 */
  int win_length;   /* <= n_fft */
  int hop;
  int n_mels;       /* <= 128 */
  int preemphasis;  /* 1: y[n] = x[n+1] - 0.97 x[n] (PaSST), 0: none */
  int wav_norm;     /* 1: divide by per-clip peak + 1e-10 (needs peak buffer filled by t4s_wav_peak) */
  int magnitude;    /* 0: power |X|^2 (PaSST), 1: magnitude |X| (DCASE, power=1) */
  int out_mode;     /* 0: linear mel, 1: (ln(x+1e-5)+4.5)/5, 2: clamp(20 log10(max(x,1e-5)), -50, 80) */
  int out_dtype;    /* T4S_F32 or T4S_BF16 */
} T4sMelParams;

/* peak[b] = max |wav[b,:]|  (passt_feature_extraction.py:46-51).  wav [B, L] f32 row-major. */
int t4s_wav_peak(const float* wav, float* peak, int batch, int n_samples, void* stream);

/* Precompute twiddle/window tables.  tables: device buffer of t4s_mel_tables_bytes(n_fft, win_length) bytes.
 * window_host: win_length fp32 host values (hann / hamming, periodic=False), copied with cudaMemcpyAsync. */
size_t t4s_mel_tables_bytes(int n_fft, int win_length);
int t4s_mel_tables_init(void* tables, const float* window_host, int n_fft, int win_length, void* stream);

/* Sparse mel basis in CSR-by-row form: row m covers bins [bin_start[m], bin_start[m]+bin_count[m]) with weights
 * weights[w_offset[m] ...].  Every row has <= 32 taps. */
int t4s_mel_forward(const float* wav, const float* peak, const void* tables,
                    const int* bin_start, const int* bin_count, const int* w_offset, const float* weights, int n_weights,
                    void* out /* [B, n_mels, T] */, int batch, int n_samples, int n_frames,
                    const T4sMelParams* p, void* stream);

/* normalize only: out = (ln(in + 1e-5) + 4.5) / 5  (passt_feature_extraction.py:91-94). */
int t4s_mel_normalize(const float* in, float* out, size_t n, void* stream);

/* ---- K3: tcgen05 / TMA GEMM with fused epilogue -----------------------------------------------------------
 * C[z][M,N] = act(alpha * A[z][M,K] . B[z][N,K]^T + bias[N]) + residual[z][M,N]
 * Both operands are K-major ("x @ W^T", exactly nn.Linear's layout), bf16 (tcgen05 kind::f16) or fp32 (kind::tf32),
 * fp32 accumulation in TMEM.  Replaces every nn.Linear / torch.matmul / bmm on the path:
 *   src/models/passt/passt.py:270-276 (Mlp fc1+GELU, fc2), :333,:342 (qkv, proj), :336,:341 (q k^T, attn v),
 *   :302-315 (patch-embed conv as GEMM), src/models/transformer/transformerXL.py:372-374 (in_proj), :487 (linear_pos),
 *   :510-513 (matrix_ac / matrix_bd), :566-576 (bmm, out_proj), src/models/passt/passt_sed.py:194-196 (mlm_mlp).
 * z = (z1, z2) is a two-level batch index (e.g. clip, head); an operand with nb == 1 at a level is broadcast.
 */
typedef struct {
  const void* ptr;  /* element (row 0, k 0) of batch (0,0) */
  int64_t rows;     /* M for A, N for B */
  int64_t ld;       /* K-major: elements between consecutive rows (K contiguous);
                       MN-major: elements between consecutive K indices (rows contiguous) */
  int64_t nb1, stride1, nb2, stride2; /* batch extents and strides in elements */
  int mn_major;     /* 0: stored [rows][K]; 1: stored [K][rows] (transposed operand, no copy needed for dgrad/wgrad) */
} T4sOperand;

typedef struct {
  void* ptr;        /* NULL = absent */
  int dtype;        /* T4S_F32 or T4S_BF16 */
  int64_t ld, stride1, stride2;
} T4sMatrix;

#define T4S_ACT_NONE 0
#define T4S_ACT_GELU 1

typedef struct {
  int M, N, K;
  int in_dtype;       /* T4S_BF16 or T4S_F32 (tf32 tensor-core math) */
  int nb1, nb2;       /* batch grid; total batches = nb1 * nb2 */
  int split_k;        /* >1: K is cut into split_k ranges; range s writes its partial product to C + s*c_split_stride
                         (epilogue extras apply to every partial; use t4s_reduce_splits afterwards) */
  int64_t c_split_stride;
  T4sOperand A, B;
  T4sMatrix C;        /* output (required) */
  T4sMatrix aux;      /* optional second output: alpha*acc + bias, i.e. the pre-activation (saved for backward) */
  T4sMatrix residual; /* optional, added after the activation; may alias C (accumulate) */
  const float* bias;  /* optional [N] fp32 */
  float alpha;
  int act;
} T4sGemm;

int t4s_gemm(const T4sGemm* g, void* stream);

/* out[i] = (accumulate ? out[i] : 0) + sum_s ws[s*n + i]   (fp32; finishes a split-K weight-gradient GEMM) */
int t4s_reduce_splits(const float* ws, int splits, size_t n, float* out, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T4S_H_ */
