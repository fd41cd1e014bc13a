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
  int n_fft;        /* 1024 (PaSST, register FFT) or any other power of two in 256..4096 (shared-memory FFT) */
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
 * weights[w_offset[m] ...].  n_fft = 1024: every row has <= 32 taps; other n_fft: any span. */
int t4s_mel_forward(const float* wav, const float* peak, const void* tables,
                    const int* bin_start, const int* bin_count, const int* w_offset, const float* weights, int n_weights,
                    void* out /* [B, n_mels, T] */, int batch, int n_samples, int n_frames,
                    const T4sMelParams* p, void* stream);

/* take_log (src/preprocess/feats_extraction.py:41-44): out = clamp(multiplier * log10(max(in, amin)), lo, hi) */
int t4s_amp_to_db(const float* in, float* out, size_t n, float multiplier, float amin, float lo, float hi, void* stream);
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
/* backward of GELU fused into a dgrad GEMM: C = (alpha * acc + bias) * gelu'(residual), `residual` holding the saved
 * pre-activation (it is multiplied in, not added) */
#define T4S_ACT_GELU_GRAD 2

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
  float* colsum;      /* optional [N] fp32 (zeroed by the caller): colsum[n] += sum_m C[m, n], the un-rounded epilogue values added with
                         atomics by the epilogue warps -- the bias gradient of the layer whose output gradient this GEMM produces
                         (fc1: the fc2-dgrad GEMM with T4S_ACT_GELU_GRAD) without re-reading C.  bf16 un-split outputs only. */
  int band_lo, band_hi; /* optional (band_hi > 0): the caller guarantees A[z][m, k] == 0 unless band_lo <= m + k < band_hi (an anti-diagonal
                         band, e.g. the un-shifted position-score gradient of csrc/attn_rel.cu); k-blocks wholly outside the band of
                         an M tile are neither loaded nor multiplied */
} T4sGemm;

int t4s_gemm(const T4sGemm* g, void* stream);

/* out[i] = (accumulate ? out[i] : 0) + sum_s ws[s*n + i]   (fp32; finishes a split-K weight-gradient GEMM) */
int t4s_reduce_splits(const float* ws, int splits, size_t n, float* out, int accumulate, void* stream);

/* 3xTF32 operand preparation for the error-compensated parity mode: gathers a (strided, optionally MN-major) fp32
 * operand into a contiguous K-major [nb2][nb1][rows][ld_dst] buffer of tf32-exact values; pattern 0 = [hi|lo|hi] (A side),
 * pattern 1 = [hi|hi|lo] (B side), so one tf32 GEMM over 3K gives hi*hi + lo*hi + hi*lo (fp32-class accuracy). */
int t4s_split_tf32(const T4sOperand* src, int K, float* dst, int64_t ld_dst /* row pitch of dst, >= 3K */, int pattern, void* stream);

/* ---- K2: LayerNorm, softmax and other row kernels ----------------------------------------------------------
 * Replace nn.LayerNorm (passt.py:360-363,410; passt_sed.py:126,203; transformerXL.py:32,34), softmax (passt.py:339;
 * transformerXL.py:549), rel_shift + softmax (transformerXL.py:254-297,514-549), exact GELU backward, bias gradients.
 * Row r of x is read at x + (r / n_inner) * x_bstride + (r % n_inner) * cols (n_inner <= 0: contiguous), which lets a
 * [B, skip:, C] token slice be normalised in place; y / mean / rstd are contiguous.  `in_scale` multiplies x first
 * (the decoder's x*sqrt(d), transformerXL.py:118). */
int t4s_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows,
                      int cols, float eps, float in_scale, int dtype, int64_t n_inner, int64_t x_bstride, void* stream);
size_t t4s_layernorm_bwd_workspace(int64_t rows, int cols);
/* dx (same addressing as x) = dLN/dx (+ dx_add, same addressing); dgamma/dbeta may be NULL (frozen). */
/* dx_colsum (optional [cols] fp32, bf16 fast path only): column sums of the OUTPUT dx (after dx_add).  dx is the gradient of the residual
 * stream, i.e. the output gradient of the linear layer that wrote it (proj / fc2), so this is that layer's bias gradient for free. */
int t4s_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dx_add,
                      void* dx, float* dgamma, float* dbeta, float* dx_colsum, float* ws, size_t ws_bytes, int64_t rows, int cols,
                      float in_scale, int dtype, int64_t n_inner, int64_t x_bstride, void* stream);
size_t t4s_colsum_workspace(int64_t rows, int cols);
/* out[c] (+)= sum_r x[r*ld + c] */
int t4s_colsum(const void* x, int dtype, int64_t rows, int cols, int64_t ld, float* ws, size_t ws_bytes, float* out, int accumulate,
               void* stream);
int t4s_gelu_bwd(const void* dy, const void* h, void* dh, size_t n, int dtype, void* stream);
int t4s_softmax_fwd(const void* s, void* p, int64_t rows, int cols, int64_t ld_s, int64_t ld_p, int dtype, void* stream);
/* dp <- p * (dp - sum(p*dp)) in place */
int t4s_softmax_bwd(const void* p, void* dp, int64_t rows, int cols, int64_t ld_p, int64_t ld_dp, int dtype, void* stream);
/* p[r, j] = softmax_j(ac[r, j] + bd[r, T-1-(r%T)+j]); rows = batch*heads*T */
int t4s_relpos_softmax_fwd(const void* ac, const void* bd, void* p, int64_t rows, int T_len, int64_t ld_ac, int64_t ld_bd,
                           int64_t ld_p, int dtype, void* stream);
/* dp <- ds in place (= d ac); dbd row <- ds scattered to the shifted columns, zero elsewhere */
int t4s_relpos_softmax_bwd(const void* p, void* dp, void* dbd, int64_t rows, int T_len, int64_t ld_p, int64_t ld_dp, int64_t ld_bd,
                           int dtype, void* stream);

/* ---- K4: fused multi-head self-attention (csrc/attn.cu) ------------------------------------------------------
 * Replaces passt.py:330-341 (Attention.forward: q k^T * hd^-0.5 -> softmax -> . v) forward and backward without ever
 * materialising the [B, H, N, N] score / probability matrices: S and P tiles live in TMEM / shared memory
 * (tcgen05.mma for QK^T, PV, dP, dV, dK, dQ; online softmax in the TMEM-load epilogue warps).  bf16 operands, fp32
 * accumulation and statistics; head_dim must be 64.  Element (b, n, h, d) of q is q[b*q_bs + n*q_ld + h*64 + d]
 * (likewise k, v, o, do, dq, dk, dv), so q/k/v may be slices of one fused [B, N, 3*H*64] buffer; all bases and pitches
 * must be multiples of 8 elements (16 bytes).
 * lse [B, H, Nl] and delta [B, H, Nl] are fp32 with Nl = t4s_attn_padded_len(N); lse is in log2 units
 * (log2 sum_j exp(scale * q.k_j)), written by the forward and consumed by the backward. */
typedef struct {
  int batch, heads, tokens, head_dim;
  float scale;
  const void* q; int64_t q_ld, q_bs;
  const void* k; int64_t k_ld, k_bs;
  const void* v; int64_t v_ld, v_bs;
  void* o;       int64_t o_ld, o_bs;
  float* lse;
  float* o32;    /* optional [B, N, heads*64] fp32 copy of o (un-rounded): when given, the backward forms
                    delta = rowsum(dO * o32) from it so that sum_j dS[i, j] cancels to fp32 rather than bf16 accuracy */
} T4sAttn;
typedef struct {
  T4sAttn fwd;          /* q, k, v, o, lse as given to / produced by t4s_attn_fwd */
  const void* d_o; int64_t do_ld, do_bs;
  float* delta;         /* workspace [B, H, Nl]: rowsum(dO * O), filled by the call */
  void* dq; int64_t dq_ld, dq_bs;
  void* dk; int64_t dk_ld, dk_bs;
  void* dv; int64_t dv_ld, dv_bs;
  float* dq32;          /* optional workspace [B, N, heads*64] fp32.  When given, ONE fused kernel computes dK / dV and adds the dQ tiles
                           into dq32 with TMA reduce-add (S, P, dS are formed once instead of twice), then dq = scale * dq32 is written;
                           the fp32 summation order over key tiles is not fixed.  NULL: the deterministic two-kernel backward. */
  float* dqkv_colsum;   /* optional [3 * heads * 64] fp32, zeroed by the caller (one-kernel backward only): column sums of dq | dk | dv as
                           written, i.e. the bias gradient of the qkv projection (passt.py:330-334), accumulated with atomics by the
                           kernels that produce dq / dk / dv instead of a separate pass over dqkv */
} T4sAttnBwd;
int64_t t4s_attn_padded_len(int tokens);
int t4s_attn_fwd(const T4sAttn* a, void* stream);
int t4s_attn_bwd(const T4sAttnBwd* a, void* stream);

/* ---- K5: fused Transformer-XL relative-position attention (csrc/attn_rel.cu) -----------------------------------
 * Replaces transformerXL.py:299-593 (RelPositionMultiheadAttention core: AC + rel_shift(BD) -> softmax -> . v) and
 * rel_shift (:254-297) forward and backward:  score[i, j] = (qu_i . k_j + qv_i . pos[T-1-i+j]) * scale with
 * qu = q + pos_bias_u, qv = q + pos_bias_v, pos = linear_pos(pos_emb) [2T-1, heads*64] (row k <-> relative position T-1-k).
 * Layout rules as for t4s_attn_*.  The backward writes dqu = scale * dS K, dk = scale * dS^T qu, dv = P^T dO and streams
 * dS back to position coordinates: dbd [B, H, T, dbd_ld] (bf16), row i receives dS[i, :] at columns T-1-i .. 2T-2-i; every
 * other column is left untouched (the caller keeps the buffer zero outside that band, e.g. by zeroing it once), so
 * d(qv) = scale * dbd . pos and d(pos) = scale * sum_b dbd^T . qv are plain GEMMs. */
typedef struct {
  int batch, heads, tokens, head_dim;
  float scale;
  const void* qu; int64_t qu_ld, qu_bs;
  const void* qv; int64_t qv_ld, qv_bs;
  const void* k;  int64_t k_ld, k_bs;
  const void* v;  int64_t v_ld, v_bs;
  const void* pos; int64_t pos_ld;
  void* o;        int64_t o_ld, o_bs;
  float* lse;
  float* o32;     /* optional, as in T4sAttn */
} T4sRelAttn;
typedef struct {
  T4sRelAttn fwd;
  const void* d_o; int64_t do_ld, do_bs;
  float* delta;
  void* dqu; int64_t dqu_ld, dqu_bs;
  void* dk;  int64_t dk_ld, dk_bs;
  void* dv;  int64_t dv_ld, dv_bs;
  void* dbd; int64_t dbd_ld;
} T4sRelAttnBwd;
int t4s_relattn_fwd(const T4sRelAttn* a, void* stream);
int t4s_relattn_bwd(const T4sRelAttnBwd* a, void* stream);

/* ---- layout / glue kernels (csrc/misc.cu) -------------------------------------------------------------------
 * passt.py:302-315 (patch conv as im2col + GEMM), :503-519,:560-569 (positional tables, cls/dist tokens),
 * passt_sed.py:199-218 (frequency mean-pool), :23-34,:258-259 (pad + linear interpolation). */
/* out [(b, f, t), patch*patch] for f < f_dim, t < t_dim (the grid may be cropped, passt.py:515) */
int t4s_patch_im2col(const void* img, int img_dtype, void* out, int out_dtype, int batch, int height, int width, int patch, int stride,
                     int f_dim, int t_dim, void* stream);
int t4s_patch_posbias(const float* time_pos, const float* freq_pos, float* out, int dim, int f_dim, int t_dim, int t_table, int t_offset,
                      void* stream);
int t4s_cls_dist_tokens(void* x, int dtype, const float* cls, const float* dist, const float* new_pos, int batch, int64_t batch_stride,
                        int dim, void* stream);
/* tmp: (2 + f_dim*t_dim) * dim floats.  Any gradient pointer may be NULL. */
int t4s_patch_small_grads(const void* dx, int dtype, float* tmp, float* d_time, float* d_freq, float* d_bias, float* d_cls, float* d_dist,
                          float* d_newpos, int batch, int64_t batch_stride, int dim, int f_dim, int t_dim, int t_table, int t_offset,
                          void* stream);
int t4s_fpool_mean_fwd(const void* y, void* out, int dtype, int batch, int f_dim, int t_dim, int dim, void* stream);
int t4s_fpool_mean_bwd(const void* dout, void* dy, int dtype, int batch, int f_dim, int t_dim, int dim, void* stream);
/* x [B, t_in, C] -> out [B, (t_in+pad)*ratio, C]; pad = 1 repeats the last frame first (99 -> 100 frames) */
int t4s_pad_interp_fwd(const void* x, void* out, int dtype, int batch, int t_in, int ratio, int dim, int pad, void* stream);
int t4s_pad_interp_bwd(const void* dout, void* dx, int dtype, int batch, int t_in, int ratio, int dim, int pad, void* stream);
/* out[r, c] = scale * x[r*ld + c] + vec[c]  (vec may be NULL) */
int t4s_add_rowvec(const void* x, int64_t ld, const float* vec, void* out, int64_t rows, int cols, float scale, int dtype, void* stream);
/* out[r*ldo + c] = alpha * x[r*ldx + c] + beta * y[r*ldy + c] */
int t4s_add2(const void* x, int64_t ldx, const void* y, int64_t ldy, void* out, int64_t ldo, int64_t rows, int cols, float alpha, float beta,
             int dtype, void* stream);
int t4s_convert(const void* in, int in_dtype, void* out, int out_dtype, size_t n, void* stream);

/* ---- sliding-window global-local fusion and MLM frame masking (csrc/window.cu) -------------------------------------------
 * src/models/encoder_slide_window.py:16-36 + src/models/passt/passt_win.py:23-41: the reference loops over 11 (train teacher) or
 * 17 (validation) time windows in Python, one backbone pass each.  Here every window is one more sequence of ONE backbone pass:
 * t4s_patch_im2col_windows gathers the patches of all windows (out [(w, b, f, t), patch*patch]; `width` is the row pitch of the
 * full image, starts[w] the first mel frame of window w; `starts` is a HOST array), and t4s_window_overlap_add_* averages the
 * per-window frame embeddings back onto the clip's time axis (out[b,t] = mean over covering windows, 0 where none covers). */
#define T4S_MAX_WINDOWS 64
int t4s_patch_im2col_windows(const void* img, int img_dtype, void* out, int out_dtype, int batch, int height, int width, const int* starts,
                             int n_windows, int patch, int stride, int f_dim, int t_dim, void* stream);
typedef struct {
  void* ptr;            /* frame 0 of clip 0 of this window: [batch][frames][dim] with clips batch_stride elements apart */
  int64_t batch_stride;
  int out_start;        /* output frame the window's frame 0 lands on (round(w_left * scale), encoder_slide_window.py:31) */
  int frames;           /* frames of this window; those past the end of the output are dropped (zero gradient) */
} T4sWindowSegment;
/* segs: HOST array of n_windows entries (<= T4S_MAX_WINDOWS). */
int t4s_window_overlap_add_fwd(const T4sWindowSegment* segs, int n_windows, void* out, int dtype, int batch, int frames, int dim, void* stream);
/* writes the gradient of every window through segs[w].ptr */
int t4s_window_overlap_add_bwd(const void* dout, const T4sWindowSegment* segs, int n_windows, int dtype, int batch, int frames, int dim,
                               void* stream);
/* src/models/transformer/mask.py:62-82 (MlmModule.setence_mask): out[r] = kind[r]==1 ? token : kind[r]==2 ? x[src[r]] : x[r]
 * (kind: 0 keep / "self", 1 mask token, 2 random other frame of the un-masked sequence).  The backward routes dout back to
 * x (including the copied-from rows, summed in ascending row order: deterministic) and to the mask token.
 * copy_rows: the n_copy_rows row indices with kind == 2, ascending (device).  dx or d_token may be NULL. */
#define T4S_MASK_GRAD_BLOCKS 64
int t4s_mask_rows_fwd(const void* x, const float* token, const unsigned char* kind, const int64_t* src, void* out, int64_t rows, int dim,
                      int dtype, void* stream);
size_t t4s_mask_rows_bwd_workspace(int64_t rows, int dim);
int t4s_mask_rows_bwd(const void* dout, const unsigned char* kind, const int64_t* src, const int64_t* copy_rows, int n_copy_rows, void* dx,
                      float* d_token, float* ws, size_t ws_bytes, int64_t rows, int dim, int dtype, void* stream);

/* ---- PMAM / DASM CNN branch and heads (csrc/cnn.cu) -----------------------------------------------------------------------
 * src/models/cnn/base.py:33-113 (CNN: Conv2d 3x3 -> BatchNorm2d(eps 1e-3, momentum .99) -> ContextGating -> Dropout -> AvgPool2d),
 * src/models/cnn_transformer/passt_cnn.py:52-62 (merge with the transformer features), recipes/desed/pmam/train.py:82-87 (prototype
 * head).  Activations are channels-last [B, H, W, C]; the convolutions and the gating Linear run as t4s_gemm on [pixels, channels]. */
/* col [(b,h,w), k_padded]: column (ky*3+kx)*C + c = in[b, h+ky-1, w+kx-1, c] (0 outside); in is addressed by element strides
 * (channel stride 1), so a [B, F, T] mel image can be read as [B, T, F, 1] in place. */
int t4s_im2col3x3(const void* in, int in_dtype, int64_t stride_b, int64_t stride_h, int64_t stride_w, void* col, int col_dtype, int batch,
                  int height, int width, int channels, int k_padded, void* stream);
int t4s_col2im3x3(const void* dcol, void* din, int dtype, int batch, int height, int width, int channels, int k_padded, void* stream);
size_t t4s_chan_stats_workspace(int64_t rows, int channels);
/* x, y [rows, channels]; training != 0: batch statistics (+ running-statistics update when the pointers are given), else running
 * statistics.  mean / rstd [channels] are outputs (kept for the backward). */
int t4s_batchnorm_fwd(const void* x, void* y, int dtype, int64_t rows, int channels, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float eps, float momentum, int training, float* mean, float* rstd, float* ws, size_t ws_bytes,
                      void* stream);
int t4s_batchnorm_bwd(const void* dy, const void* x, int dtype, int64_t rows, int channels, const float* gamma, const float* mean, const float* rstd,
                      int training, void* dx, float* dgamma, float* dbeta, float* ws, size_t ws_bytes, void* stream);
/* ContextGating product with optional inverted dropout: out = y * sigmoid(lin) * keep(seed, i) / (1 - p) */
int t4s_gate_fwd(const void* y, const void* lin, void* out, size_t n, float dropout_p, uint64_t seed, int dtype, void* stream);
int t4s_gate_bwd(const void* dout, const void* y, const void* lin, void* dy, void* dlin, size_t n, float dropout_p, uint64_t seed, int dtype,
                 void* stream);
int t4s_avgpool_fwd(const void* x, void* out, int dtype, int batch, int height, int width, int channels, int pool_h, int pool_w, void* stream);
int t4s_avgpool_bwd(const void* dout, void* dx, int dtype, int batch, int height, int width, int channels, int pool_h, int pool_w, void* stream);
/* out = a + w[0] * b (w: learnable scalar in device memory); backward: db = w dout (optional), dw = <dout, b>; ws: T4S_SCALE_ADD_BLOCKS floats */
#define T4S_SCALE_ADD_BLOCKS 256
int t4s_scale_add_fwd(const void* a, const void* b, const float* w, void* out, size_t n, int dtype, void* stream);
int t4s_scale_add_bwd(const void* dout, const void* b, const float* w, void* db, float* dw, float* ws, size_t n, int dtype, void* stream);
/* y = x / max(||x||_2, 1e-12) per row (F.normalize); inv_norm [rows] kept for the backward */
int t4s_l2norm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int cols, int dtype, void* stream);
int t4s_l2norm_bwd(const void* dy, const void* y, const float* inv_norm, void* dx, int64_t rows, int cols, int dtype, void* stream);
/* p = sigmoid((leaky_relu(s, slope) * 2 - 1) / temperature) */
int t4s_proto_act_fwd(const float* s, float* p, size_t n, float slope, float temperature, void* stream);
int t4s_proto_act_bwd(const float* s, const float* p, const float* dp, float* ds, size_t n, float slope, float temperature, void* stream);

/* ---- K6/K7: heads and losses (csrc/head.cu) ------------------------------------------------------------------
 * passt_sed.py:285-296 (sigmoid, pad mask, linear-softmax pool), pooling.py:37-51 (AttentionPooling),
 * recipes/desed/finetune/train.py:166-178 (BCE / MSE), recipes/desed/mlm/mlm_passt/train.py:36-38 (masked MSE). */
int t4s_sed_pool_fwd(const float* logits, const unsigned char* pad_mask, float temp, float* strong, float* weak, int batch, int frames,
                     int classes, void* stream);
int t4s_sed_pool_bwd(const float* strong, const float* dstrong, const float* dweak, const unsigned char* pad_mask, float temp, float* dlogits,
                     int batch, int frames, int classes, void* stream);
int t4s_sigmoid_fwd(const float* x, float* y, size_t n, void* stream);
int t4s_sigmoid_bwd(const float* y, const float* dy, float* dx, size_t n, void* stream);
/* out[0] = mean loss, out[1] = divisor; ws >= 512 floats */
int t4s_bce_fwd(const float* p, const float* y, size_t n, float* ws, float* out, void* stream);
int t4s_bce_bwd(const float* p, const float* y, const float* grad_out, size_t n, float* dp, void* stream);
int t4s_mse_fwd(const void* a, const void* b, const unsigned char* row_mask, int64_t rows, int cols, int dtype, float* ws, float* out,
                void* stream);
int t4s_mse_bwd(const void* a, const void* b, const unsigned char* row_mask, int64_t rows, int cols, int dtype, const float* grad_out,
                const float* fwd_out, void* da, void* db, void* stream);
/* kv: item i at kv + i*item_stride holds [keys, 2*dim] (k | v per key); item_stride <= 0 means keys*2*dim (a larger
 * stride skips leading tokens such as cls/dist).  q [dim] fp32 pre-scaled; ctx [items, dim]; probs [items, heads, keys] */
int t4s_attnpool_fwd(const void* kv, const float* q, void* ctx, float* probs, int items, int keys, int dim, int heads, int64_t item_stride,
                     int dtype, void* stream);
int t4s_attnpool_bwd(const void* kv, const float* q, const float* probs, const void* dctx, void* dkv, float* dq_part, int items, int keys,
                     int dim, int heads, int64_t item_stride, int dtype, void* stream);

/* ---- DASM open-vocabulary head (csrc/head.cu, csrc/cnn.cu) ----------------------------------------------------------------------
 * src/models/detect_any_sound/detect_any_sound.py:376-388: strong[b,k,t] = clamp(sigmoid(score[b,t,k]/temp) * at_out[b,k], 1e-7, 1)
 * (padded frames forced to 0 before the clamp), weak = clamp(sum_t p^2 / sum_t p, 1e-7, 1).  score is the query x frame GEMM output. */
int t4s_query_pool_fwd(const float* score, const float* at_out, const unsigned char* pad_mask, float temp, float* strong, float* weak, int batch,
                       int frames, int queries, void* stream);
int t4s_query_pool_bwd(const float* score, const float* at_out, const float* strong, const float* dstrong, const float* dweak,
                       const unsigned char* pad_mask, float temp, float* dscore, float* dat, int batch, int frames, int queries, void* stream);
/* boolean attn_mask of nn.MultiheadAttention (at_adapter.py:28-31, tgt_mask): s[(b,h,q), c] = -inf where mask[q, c] != 0 */
int t4s_mask_scores(void* s, const unsigned char* mask, int64_t rows, int cols, int64_t ld, int n_queries, int dtype, void* stream);
/* inverted dropout with a counter-based mask (seed, element index); the backward is the same call on the gradient */
int t4s_dropout(const void* x, void* out, size_t n, float dropout_p, uint64_t seed, int dtype, void* stream);

/* ---- train-loop glue next to the hot path (csrc/aug.cu), fp32 ----------------------------------------------
 * frame_shift (src/preprocess/data_aug.py:12-31): out[b, r, (t + shifts[b]) mod len] = x[b, r, t]; shifts: DEVICE int32 [batch] */
int t4s_roll_rows(const float* x, float* out, const int* shifts_dev, int batch, int rows, int len, void* stream);
/* mixup (data_aug.py:34-91): out[b] = w_self x[b] + w_other x[perm[b]] (perm: DEVICE int64 [batch]), optionally clamped to [0, 1] */
int t4s_mixup(const float* x, const int64_t* perm_dev, float* out, int batch, int64_t inner, float w_self, float w_other, int clamp01, void* stream);
/* freq_nonlinear (data_aug.py:239-254): out[b,k,t] = lerp(x[b, src_bin[k], t], x[b, src_bin[k]+1, t], weight[k]); src_bin / weight are the
 * np.interp knots of the warped frequency axis, computed once on the host (DEVICE arrays of n_freq entries) */
int t4s_freq_warp(const float* x, float* out, const int* src_bin_dev, const float* weight_dev, int batch, int n_freq, int n_frames, void* stream);
/* FilterAugment on log-mel features (data_aug.py:150-190): out[r, t] = x[r, t] + bias[r], r = (clip, mel bin); bias: DEVICE [rows] */
int t4s_add_rowbias(const float* x, const float* bias_dev, float* out, int64_t rows, int cols, void* stream);
/* class-wise median filter of the post-processing (src/postprocess/filter.py:4-36): in / out [batch, length, classes]; class c uses the
 * odd window window_sizes[c] (HOST array) with replicate padding */
#define T4S_MEDIAN_MAX_CLASSES 32
#define T4S_MEDIAN_MAX_WINDOW 255
int t4s_median_filter(const float* in, float* out, const int* window_sizes, int batch, int length, int classes, void* stream);

/* ---- callers either side of the hot path (csrc/post.cu; SURVEY §8 a1', f3, f4), fp32 ------------------------------------------------
 * TorchScaler.forward (src/preprocess/scaler.py:91-121) with statistics over every dimension but the first:
 * instance mode 0 'mean' (x - mean), 1 'standard' ((x - mean) / (std + eps), unbiased std like torch.std), 2 'minmax'
 * ((x - min) / (max - min + eps)); dataset mode with the fitted scalars `mean`, `mean_squared` (DEVICE, one float each) */
int t4s_scaler_instance(const float* x, float* out, int batch, int64_t inner, int mode, float eps, void* stream);
int t4s_scaler_dataset(const float* x, float* out, int64_t total, const float* mean_dev, const float* mean_sq_dev, int standard, float eps, void* stream);
/* scipy.ndimage.median_filter / maximum_filter per class as batched_decode_preds applies them (src/codec/decoder.py:86-92):
 * in / out [batch, length, classes]; window of class c = window_sizes[c] (HOST array, any size in [1, 255]) samples starting at
 * l - size / 2, borders mirrored with the edge sample repeated (scipy 'reflect'); op 0 = element of rank size / 2, 1 = maximum */
int t4s_rank_filter(const float* in, float* out, const int* window_sizes, int batch, int length, int classes, int op, void* stream);
/* Threshold sweep + run-length event decoding (src/codec/decoder.py:15-35 decode_pred_batch_fast, src/codec/encoder.py:51-84
 * decode_strong / find_contiguous_regions).  scores [batch, length, classes] are the filtered frame probabilities, weak [batch, classes]
 * (may be NULL) the clip probabilities, thresholds_dev DEVICE [n_thresholds].  Column q = (threshold, clip, class) is silent when
 * weak < threshold; its events are the maximal runs of score > threshold as (onset frame, offset frame = last + 1).
 * Pass 1 (events == NULL): counts[q] = number of events.  Pass 2: offsets[q] (DEVICE int64, exclusive prefix sum of counts) and
 * events [total, 5] int32 rows (threshold index, clip, class, onset, offset), ordered by q then by onset. */
int t4s_event_sweep(const float* scores, const float* weak, const float* thresholds_dev, int n_thresholds, int batch, int length, int classes, int* counts,
                    const int64_t* offsets, int* events, void* stream);
/* The six losses of the mean-teacher step and their weighted total in one pass (recipes/desed/finetune/train.py:166-188):
 *   out[0] BCE(strong[s0:s1], y[s0:s1])   out[1] BCE(weak[w0:w1], yw[w0:w1])   out[2] BCE(at[w0:w1], yw[w0:w1])
 *   out[3] MSE(strong, t_strong)          out[4] MSE(weak, t_at)               out[5] MSE(at, t_at)
 *   out[6] = out[0] + w_weak out[1] + w_at out[2] + w_cons (out[3] + w_weak_cons out[4] + w_at out[5])
 * strong / t_strong / y: [batch, strong_inner] (classes x frames in any common layout); weak / at / t_at / yw: [batch, classes].
 * Deterministic two-stage reductions (ws: 768 floats).  The backward writes d out[6] / d strong, d weak, d at times grad_total[0]. */
typedef struct {
  const float *strong, *weak, *at, *t_strong, *t_at, *y, *yw;
  int batch, classes;
  int64_t strong_inner;
  int s0, s1, w0, w1;
  float w_weak, w_at, w_cons, w_weak_cons;
} T4sSedLosses;
int t4s_sed_losses_fwd(const T4sSedLosses* p, float* ws, float* out, void* stream);
int t4s_sed_losses_bwd(const T4sSedLosses* p, const float* grad_total, float* d_strong, float* d_weak, float* d_at, void* stream);

/* ---- K9: parameter-side kernels of a training step (csrc/optim.cu) ------------------------------------------------
 * torch.optim.AdamW semantics (recipes/desed/setting.py:254-258) over a flat fp32 arena; `bf16_shadow` (optional) receives the
 * updated weights as bf16 GEMM operands in the same pass; `grad_scale` folds the 1/world_size of the gradient all-reduce. */
int t4s_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* bf16_shadow, size_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* teacher = alpha*teacher + (1-alpha)*student  (src/utils/scheduler.py:125-130) */
int t4s_ema_update(float* teacher, const float* student, void* bf16_shadow, size_t n, float alpha, void* stream);
/* Gather per-tensor gradients into the flat all-reduce buffer.  table (DEVICE memory): n_entries x {const float* src (NULL = zero fill),
 * int64 offset, int64 n}. */
int t4s_grad_pack(const void* table, int n_entries, float* flat, void* stream);
/* number of kernel launches issued by this library in this process so far */
long long t4s_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* T4S_H_ */
