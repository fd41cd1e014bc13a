"""Host-side construction of the sparse mel basis handed to the K1 kernel.

The basis is tiny (128 x 513, ~1.4 % non-zero) and, in train mode, changes every step with the
fmin/fmax jitter (reference src/models/passt/passt_feature_extraction.py:66-82), so it is built on
the host in fp32 exactly as Kaldi defines it, converted to a row-CSR of contiguous taps, cached per
(fmin, fmax) and uploaded once.
"""
import math

import numpy as np
import torch


def kaldi_mel_banks(n_mels, n_fft, sr, fmin, fmax):
    """Kaldi triangular banks [n_mels, n_fft//2 + 1] (Nyquist column zero) as fp32 numpy.
    mel(f) = 1127 ln(1 + f/700); vtln warp factor 1 (passt_feature_extraction.py:73-81).
    Evaluated with torch CPU fp32 ops so the weights are bit-identical to what the reference's
    torchaudio call produces on the host."""
    nyq = 0.5 * sr
    if fmax <= 0.0:
        fmax += nyq
    n_bins = n_fft // 2
    mel_lo = 1127.0 * math.log(1.0 + fmin / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + fmax / 700.0)
    delta = (mel_hi - mel_lo) / (n_mels + 1)
    b = torch.arange(n_mels, dtype=torch.float32).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    mel = 1127.0 * (1.0 + (sr / n_fft) * torch.arange(n_bins, dtype=torch.float32) / 700.0).log().unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    w = torch.max(torch.zeros(1), torch.min(up, down))
    return torch.nn.functional.pad(w, (0, 1)).numpy()


def htk_mel_banks(n_mels, n_fft, sr, fmin, fmax):
    """torchaudio melscale_fbanks(mel_scale='htk', norm=None) -> [n_mels, n_fft//2+1]
    (reference src/preprocess/feats_extraction.py:47-57)."""
    n_freqs = n_fft // 2 + 1
    freqs = torch.linspace(0, sr // 2, n_freqs)
    m_lo = 2595.0 * math.log10(1.0 + fmin / 700.0)
    m_hi = 2595.0 * math.log10(1.0 + fmax / 700.0)
    m_pts = torch.linspace(m_lo, m_hi, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0).t().contiguous().numpy()


def to_row_csr(basis, max_taps=128):
    """Dense [n_mels, n_bins] -> (bin_start, bin_count, w_offset int32 [n_mels], weights fp32 [nnz_span]).
    Each row is stored as its contiguous span first-nonzero..last-nonzero (interior zeros kept)."""
    n_mels = basis.shape[0]
    bs = np.zeros(n_mels, np.int32)
    bc = np.zeros(n_mels, np.int32)
    wo = np.zeros(n_mels, np.int32)
    chunks, off = [], 0
    for m in range(n_mels):
        nz = np.nonzero(basis[m])[0]
        if nz.size:
            lo, hi = int(nz[0]), int(nz[-1]) + 1
            if hi - lo > max_taps:
                raise ValueError(f"mel row {m} spans {hi - lo} bins (> {max_taps})")
            bs[m], bc[m], wo[m] = lo, hi - lo, off
            chunks.append(basis[m, lo:hi])
            off += hi - lo
        else:
            wo[m] = off
    w = np.concatenate(chunks).astype(np.float32) if chunks else np.zeros(1, np.float32)
    return bs, bc, wo, w
