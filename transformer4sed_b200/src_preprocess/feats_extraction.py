"""Drop-in for the arithmetic of reference ``src/preprocess/feats_extraction.py`` (:12-13, :41-57): the DCASE-baseline style
front end (`setmelspectrogram` = torchaudio MelSpectrogram with a non-periodic hamming window of n_window samples, power 1, HTK
mel, no filter normalisation; `take_log` = AmplitudeToDB(amplitude, amin 1e-5) clamped to [-50, 80]).  Upstream it is dead code
(no recipe calls it); it is provided because the headline metric names a 16 kHz front-end case (SURVEY §8 a1').

The returned module runs the libt4s front-end kernel (csrc/mel_generic.cu for n_fft != 1024): reflect padding, windowed FFT,
|X|, sparse mel rows -- one read of the waveform, one write of the mel image; `logmel` fuses `take_log` into the same kernel.
File loading helpers (`waveform_modification`, `pad_wav`: librosa / numpy on the host) are callers' code and are not mirrored.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, mel_basis


def normalize_wav(wav):
    return wav / (torch.max(torch.max(wav), -torch.min(wav)) + 1e-10)


class MelSpectrogram(nn.Module):
    """torchaudio.transforms.MelSpectrogram(sample_rate, n_fft=win_length=n_window, hop_length, f_min, f_max, n_mels,
    window_fn=hamming_window(periodic=False), power=1): wav [B, L] -> [B, n_mels, 1 + L // hop]."""

    def __init__(self, sample_rate, n_fft, hop_length, f_min, f_max, n_mels, power=1):
        super().__init__()
        if power not in (1, 2):
            raise NotImplementedError("power must be 1 (magnitude) or 2")
        self.sample_rate, self.n_fft, self.hop_length, self.f_min, self.f_max, self.n_mels, self.power = (
            sample_rate, n_fft, hop_length, float(f_min), float(f_max if f_max is not None else sample_rate // 2), n_mels, power)
        self.register_buffer("window", torch.hamming_window(n_fft, periodic=False), persistent=False)
        self._dev = {}

    def _consts(self, dev_idx, device):
        c = self._dev.get(dev_idx)
        if c is None:
            lib = _lib.load()
            nbytes = lib.t4s_mel_tables_bytes(self.n_fft, self.n_fft)
            if nbytes == 0:
                raise _lib.T4sError(f"unsupported n_fft {self.n_fft} (1024 or a power of two in 256..4096)")
            tables = torch.empty(nbytes, dtype=torch.uint8, device=device)
            win = self.window.detach().to("cpu", torch.float32).contiguous()
            _lib.check(lib.t4s_mel_tables_init(_lib.ptr(tables), ctypes.c_void_p(win.data_ptr()), self.n_fft, self.n_fft, _lib.stream_ptr()),
                       "t4s_mel_tables_init")
            dense = mel_basis.htk_mel_banks(self.n_mels, self.n_fft, self.sample_rate, self.f_min, self.f_max)
            bs, bc, wo, w = mel_basis.to_row_csr(dense, max_taps=4096)
            c = (tables, torch.from_numpy(np.stack([bs, bc, wo])).to(device), torch.from_numpy(w).to(device))
            self._dev[dev_idx] = c
        return c

    def _run(self, x, out_mode):
        if x.dim() != 2:
            raise ValueError(f"expected wav [B, L], got {tuple(x.shape)}")
        dev_idx = _lib.ensure_device(x)
        lib = _lib.load()
        x = x.contiguous().float()
        B, L = x.shape
        n_frames = 1 + L // self.hop_length
        with torch.cuda.device(dev_idx):
            tables, idx, w = self._consts(dev_idx, x.device)
            out = torch.empty(B, self.n_mels, n_frames, dtype=torch.float32, device=x.device)
            p = _lib.MelParams(self.n_fft, self.n_fft, self.hop_length, self.n_mels, 0, 0, int(self.power == 1), out_mode, 0)
            nm, base = self.n_mels, idx.data_ptr()
            _lib.check(lib.t4s_mel_forward(_lib.ptr(x), ctypes.c_void_p(0), _lib.ptr(tables), ctypes.c_void_p(base), ctypes.c_void_p(base + 4 * nm),
                                           ctypes.c_void_p(base + 8 * nm), _lib.ptr(w), w.numel(), _lib.ptr(out), B, L, n_frames, ctypes.byref(p),
                                           _lib.stream_ptr()), "t4s_mel_forward")
        return out

    def forward(self, wav):
        return self._run(wav, 0)

    def logmel(self, wav):
        """Fused ``take_log(self(wav))``."""
        return self._run(wav, 2)


def setmelspectrogram(feature_cfg):
    return MelSpectrogram(sample_rate=feature_cfg["sample_rate"], n_fft=feature_cfg["n_window"], hop_length=feature_cfg["hop_length"],
                          f_min=feature_cfg["f_min"], f_max=feature_cfg["f_max"], n_mels=feature_cfg["n_mels"], power=1)


def take_log(feature):
    _lib.ensure_device(feature)
    f = feature.contiguous().float()
    out = torch.empty_like(f)
    with torch.cuda.device(f.device):
        _lib.check(_lib.load().t4s_amp_to_db(_lib.ptr(f), _lib.ptr(out), f.numel(), 20.0, 1e-5, -50.0, 80.0, _lib.stream_ptr()), "t4s_amp_to_db")
    return out
