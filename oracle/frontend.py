"""Oracle: STFT -> mel -> log front end (fp32 or fp64, CPU, torch/numpy).  Test infrastructure only.

Follows reference src/models/passt/passt_feature_extraction.py:46-94 (PasstFeatureExtractor) and,
for the 16 kHz DCASE-style parametrisation, src/preprocess/feats_extraction.py:41-57.
"""
import math

import torch


def kaldi_mel_banks(n_mels, n_fft, sr, fmin, fmax, dtype=torch.float32):
    """Kaldi triangular mel banks, shape [n_mels, n_fft//2] (no Nyquist column).

    Restates torchaudio.compliance.kaldi.get_mel_banks with vtln_warp_factor=1.0 as called at
    passt_feature_extraction.py:73-80: mel(f)=1127 ln(1+f/700); n_mels+2 points equally spaced in
    mel between fmin and fmax (fmax<=0 means nyquist+fmax); weight = max(0, min(up, down)).
    """
    nyquist = 0.5 * sr
    if fmax <= 0.0:
        fmax += nyquist
    n_bins = n_fft // 2
    bin_w = sr / n_fft
    mel_lo = 1127.0 * math.log(1.0 + fmin / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + fmax / 700.0)
    delta = (mel_hi - mel_lo) / (n_mels + 1)
    b = torch.arange(n_mels, dtype=dtype).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    mel = 1127.0 * (1.0 + bin_w * torch.arange(n_bins, dtype=dtype) / 700.0).log().unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1, dtype=dtype), torch.min(up, down))


def passt_mel_basis(n_mels=128, n_fft=1024, sr=32000, fmin=0.0, fmax=15000.0, dtype=torch.float32):
    """[n_mels, n_fft//2+1]: Kaldi banks zero-padded with a Nyquist column (:81)."""
    return torch.nn.functional.pad(kaldi_mel_banks(n_mels, n_fft, sr, fmin, fmax, dtype), (0, 1))


def passt_power_mel(wav, n_mels=128, sr=32000, win_length=800, hop=320, n_fft=1024,
                    fmin=0.0, fmax=15000.0, wav_norm=True, dtype=torch.float32, mel_basis=None):
    """wav [B, L] -> power mel [B, n_mels, 1 + (L-1)//hop]   (PasstFeatureExtractor.forward, eval)."""
    x = wav.to(dtype)
    if wav_norm:  # :46-51
        peak = torch.maximum(x.max(dim=1, keepdim=True)[0].abs(), x.min(dim=1, keepdim=True)[0].abs())
        x = x / (peak + 1e-10)
    x = x[:, 1:] - 0.97 * x[:, :-1]  # pre-emphasis conv1d with taps [-.97, 1]  (:44,:56)
    pad = n_fft // 2
    x = torch.nn.functional.pad(x.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)  # center=True
    frames = x.unfold(1, n_fft, hop)  # [B, T, n_fft]
    win = torch.hann_window(win_length, periodic=False, dtype=dtype, device=x.device)  # :41
    left = (n_fft - win_length) // 2  # torch.stft centres a short window inside n_fft
    win = torch.nn.functional.pad(win, (left, n_fft - win_length - left))
    spec = torch.fft.rfft(frames * win, dim=-1)
    power = spec.real ** 2 + spec.imag ** 2  # :65   [B, T, n_fft//2+1]
    if mel_basis is None:
        mel_basis = passt_mel_basis(n_mels, n_fft, sr, fmin, fmax, dtype)
    return torch.matmul(mel_basis.to(device=power.device, dtype=dtype), power.transpose(1, 2))  # :84


def passt_normalize(melspec):
    """(ln(x + 1e-5) + 4.5) / 5   (:91-94)."""
    return ((melspec + 0.00001).log() + 4.5) / 5.0


def passt_logmel(wav, **kw):
    return passt_normalize(passt_power_mel(wav, **kw))


# ----------------------------------------------------------------------------------------------
# 16 kHz DCASE-style parametrisation (reference src/preprocess/feats_extraction.py:41-57, dead code
# upstream: torchaudio MelSpectrogram(hamming non-periodic, win=n_fft, power=1, HTK mel, norm=None)
# -> AmplitudeToDB(stype='amplitude') with amin 1e-5, clamp[-50, 80]).
# ----------------------------------------------------------------------------------------------
def htk_mel_basis(n_mels, n_fft, sr, fmin, fmax, dtype=torch.float32):
    """torchaudio.functional.melscale_fbanks(mel_scale='htk', norm=None) restated: [n_mels, n_fft//2+1]."""
    n_freqs = n_fft // 2 + 1
    freqs = torch.linspace(0, sr // 2, n_freqs, dtype=dtype)
    m_lo = 2595.0 * math.log10(1.0 + fmin / 700.0)
    m_hi = 2595.0 * math.log10(1.0 + fmax / 700.0)
    m_pts = torch.linspace(m_lo, m_hi, n_mels + 2, dtype=dtype)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - freqs.unsqueeze(1)  # [n_freqs, n_mels+2]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0).t().contiguous()


def dcase_logmel(wav, sr=16000, n_fft=2048, hop=256, n_mels=128, fmin=0.0, fmax=8000.0, dtype=torch.float32):
    """wav [B, L] -> dB mel [B, n_mels, 1 + L//hop]   (setmelspectrogram + take_log)."""
    x = wav.to(dtype)
    pad = n_fft // 2
    x = torch.nn.functional.pad(x.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    frames = x.unfold(1, n_fft, hop)
    win = torch.hamming_window(n_fft, periodic=False, dtype=dtype)
    spec = torch.fft.rfft(frames * win, dim=-1)
    mag = (spec.real ** 2 + spec.imag ** 2).sqrt()  # power=1
    mel = torch.matmul(htk_mel_basis(n_mels, n_fft, sr, fmin, fmax, dtype), mag.transpose(1, 2))
    db = 20.0 * torch.log10(torch.clamp(mel, min=1e-5))  # AmplitudeToDB(amplitude): multiplier 20, amin 1e-5
    return db.clamp(min=-50.0, max=80.0)
