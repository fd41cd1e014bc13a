"""Drop-in for reference ``src/models/passt/passt_feature_extraction.py`` (PasstFeatureExtractor).

Same constructor, same buffers, same ``forward`` / ``normalize`` contract, but the arithmetic is one
hand-written sm_100a kernel (csrc/mel.cu) behind the C ABI: peak-normalise, pre-emphasis, reflect
padding, windowed rFFT, power, sparse Kaldi mel basis and (when asked) the log normalisation fused,
so a clip costs one read of the waveform and one write of the mel image.

``forward`` returns the *power* mel like the reference (:84); ``normalize`` applies
``(ln(x+1e-5)+4.5)/5`` (:91-94).  Trainers call them back to back (reference
recipes/desed/finetune/train.py:69-73); ``logmel`` is the fused equivalent of that pair.
"""
import ctypes

import torch
import torch.nn as nn

from ... import _lib, mel_basis


class PasstFeatureExtractor(nn.Module):

    def __init__(self, n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, htk=False, fmin=0.0, fmax=None,
                 wav_norm=True, fmin_aug_range=1, fmax_aug_range=1000):
        super().__init__()
        self.win_length = win_length
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.sr = sr
        self.htk = htk
        self.fmin = fmin
        if fmax is None:
            fmax = sr // 2 - fmax_aug_range // 2
        self.fmax = fmax
        self.wav_norm = wav_norm
        self.hopsize = hopsize
        self.register_buffer("window", torch.hann_window(win_length, periodic=False), persistent=False)
        assert fmin_aug_range >= 1, f"fmin_aug_range={fmin_aug_range} should be >=1; 1 means no augmentation"
        assert fmax_aug_range >= 1, f"fmax_aug_range={fmax_aug_range} should be >=1; 1 means no augmentation"
        self.fmin_aug_range = fmin_aug_range
        self.fmax_aug_range = fmax_aug_range
        self.register_buffer("preemphasis_coefficient", torch.as_tensor([[[-.97, 1]]]), persistent=False)
        self._tables = {}   # device index -> uint8 tensor (twiddles + window)
        self._basis = {}    # (device index, fmin, fmax) -> (int32 tensor [3, n_mels], fp32 weights)

    # -- device-side constants ------------------------------------------------------------------
    def _get_tables(self, dev_idx, device):
        t = self._tables.get(dev_idx)
        if t is None:
            lib = _lib.load()
            nbytes = lib.t4s_mel_tables_bytes(self.n_fft, self.win_length)
            if nbytes == 0:
                raise _lib.T4sError(f"unsupported n_fft/win_length {self.n_fft}/{self.win_length}")
            t = torch.empty(nbytes, dtype=torch.uint8, device=device)
            win = self.window.detach().to("cpu", torch.float32).contiguous()
            _lib.check(lib.t4s_mel_tables_init(_lib.ptr(t), ctypes.c_void_p(win.data_ptr()), self.n_fft, self.win_length,
                                               _lib.stream_ptr()), "t4s_mel_tables_init")
            self._tables[dev_idx] = t
        return t

    def _get_basis(self, dev_idx, device, fmin, fmax):
        key = (dev_idx, float(fmin), float(fmax))
        b = self._basis.get(key)
        if b is None:
            dense = mel_basis.kaldi_mel_banks(self.n_mels, self.n_fft, self.sr, float(fmin), float(fmax))
            bs, bc, wo, w = mel_basis.to_row_csr(dense)
            idx = torch.from_numpy(__import__("numpy").stack([bs, bc, wo])).to(device)
            b = (idx, torch.from_numpy(w).to(device))
            if len(self._basis) > 4096:
                self._basis.clear()
            self._basis[key] = b
        return b

    def _draw_band(self):
        # same draws, same order, same global CPU RNG as the reference (:66-71)
        fmin = self.fmin + torch.randint(self.fmin_aug_range, (1,)).item()
        fmax = self.fmax + self.fmax_aug_range // 2 - torch.randint(self.fmax_aug_range, (1,)).item()
        if not self.training:
            fmin, fmax = self.fmin, self.fmax
        return fmin, fmax

    def _run(self, x, out_mode, out_dtype=torch.float32):
        if x.dim() != 2:
            raise ValueError(f"expected wav [B, L], got {tuple(x.shape)}")
        dev_idx = _lib.ensure_device(x)
        lib = _lib.load()
        x = x.contiguous().float()
        B, L = x.shape
        n_frames = 1 + (L - 1) // self.hopsize
        fmin, fmax = self._draw_band()
        with torch.cuda.device(dev_idx):
            tables = self._get_tables(dev_idx, x.device)
            idx, w = self._get_basis(dev_idx, x.device, fmin, fmax)
            peak = torch.empty(B, dtype=torch.float32, device=x.device)
            out = torch.empty(B, self.n_mels, n_frames, dtype=out_dtype, device=x.device)
            st = _lib.stream_ptr()
            if self.wav_norm:
                _lib.check(lib.t4s_wav_peak(_lib.ptr(x), _lib.ptr(peak), B, L, st), "t4s_wav_peak")
            p = _lib.MelParams(self.n_fft, self.win_length, self.hopsize, self.n_mels, 1, int(bool(self.wav_norm)), 0,
                               out_mode, 0 if out_dtype == torch.float32 else 1)
            nm = self.n_mels
            base = idx.data_ptr()
            _lib.check(lib.t4s_mel_forward(_lib.ptr(x), _lib.ptr(peak), _lib.ptr(tables),
                                           ctypes.c_void_p(base), ctypes.c_void_p(base + 4 * nm), ctypes.c_void_p(base + 8 * nm),
                                           _lib.ptr(w), w.numel(), _lib.ptr(out), B, L, n_frames, ctypes.byref(p), st),
                       "t4s_mel_forward")
        return out

    # -- reference surface ------------------------------------------------------------------------
    def forward(self, x):
        """wav [B, L] (cuda fp32) -> power mel [B, n_mels, 1 + (L-1)//hop]."""
        return self._run(x, out_mode=0)

    def normalize(self, melspec):
        """(ln(x + 1e-5) + 4.5) / 5."""
        _lib.ensure_device(melspec)
        m = melspec.contiguous().float()
        out = torch.empty_like(m)
        with torch.cuda.device(m.device):
            _lib.check(_lib.load().t4s_mel_normalize(_lib.ptr(m), _lib.ptr(out), m.numel(), _lib.stream_ptr()),
                       "t4s_mel_normalize")
        return out

    def logmel(self, x, out_dtype=torch.float32):
        """Fused ``normalize(forward(x))``: one kernel, the power mel never touches HBM."""
        return self._run(x, out_mode=1, out_dtype=out_dtype)

    def extra_repr(self):
        return "winsize={}, hopsize={}".format(self.win_length, self.hopsize)
