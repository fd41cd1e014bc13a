"""Drop-in for the label-coupled augmentations of reference ``src/preprocess/data_aug.py`` that run every training step on the
mel batch: `frame_shift` (:12-31) and `mixup` (:34-91).  The random draws are made on the host with the same generators, in the
same order, as the reference (python `random`, `torch.randperm`, `np.random.beta`), so seeded runs pick the same shifts,
permutation and mixing rate; the data movement is one libt4s kernel per tensor instead of a Python loop of `torch.roll` + stack.

`freq_nonlinear` (:239-254) and `filt_aug` (:150-195) are the label-independent feature transforms: upstream the first one moves the
batch to the host and calls `np.interp` once per (clip, frame) -- 64 000 calls per step at B = 64; here the warp is a table of 128
(source bin, weight) pairs built once on the host in float64 and one gather-lerp kernel over the batch on the GPU.
"""
import random

import numpy as np
import torch

from .. import _lib


def _roll(x, shifts):
    _lib.ensure_device(x)
    x = x.contiguous().float()
    B, R, L = x.shape
    out = torch.empty_like(x)
    sh = torch.tensor(shifts, dtype=torch.int32).to(x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().t4s_roll_rows(_lib.ptr(x), _lib.ptr(out), _lib.ptr(sh), B, R, L, _lib.stream_ptr()), "t4s_roll_rows")
    return out


def frame_shift(features, label=None, net_pooling=None, max_shift_frame=90):
    batch_size, _, _ = features.shape
    shifts = [int(random.gauss(0, max_shift_frame)) for _ in range(batch_size)]     # one draw per clip, as upstream
    shifted = _roll(features, shifts)
    if label is None:
        return shifted
    lshifts = [int(-abs(s) // net_pooling if s < 0 else s // net_pooling) for s in shifts]
    return shifted, _roll(label, lshifts)


def _mix(x, perm_dev, wa, wb, clamp):
    _lib.ensure_device(x)
    x = x.contiguous().float()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().t4s_mixup(_lib.ptr(x), _lib.ptr(perm_dev), _lib.ptr(out), x.shape[0], x[0].numel(), float(wa), float(wb), int(clamp),
                                         _lib.stream_ptr()), "t4s_mixup")
    return out


def mixup(features, label=None, permutation=None, c=None, alpha=0.2, beta=0.2, mixup_label_type="soft", power=None, repeat=True):
    with torch.no_grad():
        batch_size = features.size(0)
        if permutation is None:
            if repeat:
                permutation = torch.randperm(batch_size)
            else:
                while True:
                    permutation = torch.randperm(batch_size)
                    combine = [(min(i, int(permutation[i])), max(i, int(permutation[i]))) for i in range(batch_size)]
                    if len(set(combine)) == batch_size:
                        break
        if c is None:
            if mixup_label_type == "soft":
                c = np.random.beta(alpha, beta)
            elif mixup_label_type == "hard":
                c = np.random.beta(alpha, beta) * 0.4 + 0.3
        perm_dev = permutation.to(device=features.device, dtype=torch.int64).contiguous()
        mixed_features = _mix(features, perm_dev, c, 1 - c, False)
        if label is None:
            return mixed_features
        if mixup_label_type == "soft":
            mixed_label = _mix(label, perm_dev, c, 1 - c, True)
            if power:
                mixed_label = torch.float_power(mixed_label, power).to(mixed_label)
        elif mixup_label_type == "hard":
            mixed_label = _mix(label, perm_dev, 1.0, 1.0, True)
        else:
            raise NotImplementedError(f"mixup_label_type: {mixup_label_type} not implemented. choice in {'soft', 'hard'}")
        return mixed_features, mixed_label


def freq_warp_table(n_freq, f, bias, phase):
    """(source bin j[k], weight w[k]) such that np.interp(arange(F), ind_t, row)[k] == row[j] + w (row[j+1] - row[j]) for the warped
    knots ind_t = F * trans(arange(F) / F), trans(x) = x + bias sin(2 pi (f x + phase))  (data_aug.py:247-251).  float64, like numpy."""
    ind = np.arange(n_freq)
    xin = ind / n_freq
    ind_t = n_freq * (xin + bias * np.sin(2 * np.pi * (f * xin + phase)))      # monotone for the shipped bias <= 0.03
    j = np.clip(np.searchsorted(ind_t, ind, side="right") - 1, 0, n_freq - 2)   # ind_t[j] <= k < ind_t[j+1]
    w = (ind - ind_t[j]) / (ind_t[j + 1] - ind_t[j])
    lo, hi = ind <= ind_t[0], ind >= ind_t[-1]                                  # np.interp clamps outside the knot range
    j = np.where(lo, 0, np.where(hi, n_freq - 1, j))
    w = np.where(lo | hi, 0.0, w)
    return j.astype(np.int64), w


def freq_nonlinear(mel, f=1, bias=0.02):
    """mel [B, F, T] (cuda fp32; the reference takes / returns a numpy array) -> frequency-warped copy.  One `random.random()` draw
    (the phase), as upstream where the lambda is evaluated once on the whole index vector."""
    j, w = freq_warp_table(mel.shape[1], f, bias, random.random())
    return _launch_freq_warp(mel, j, w)


def _launch_freq_warp(mel, j, w):
    """out[b, i, t] = (1 - w[i]) * mel[b, j[i], t] + w[i] * mel[b, j[i] + 1, t] -- one gather-lerp kernel over the batch."""
    _lib.ensure_device(mel)
    x = mel.contiguous().float()
    B, Fq, T = x.shape
    j_dev = torch.from_numpy(j.astype(np.int32)).to(x.device)
    w_dev = torch.from_numpy(w.astype(np.float32)).to(x.device)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().t4s_freq_warp(_lib.ptr(x), _lib.ptr(out), _lib.ptr(j_dev), _lib.ptr(w_dev), B, Fq, T, _lib.stream_ptr()), "t4s_freq_warp")
    return out


def filt_aug(features, db_range=[-0.5, 0.5], n_band=[3, 6], min_bw=6, filter_type="step", log=False, norm_std=1):
    """FilterAugment (ICASSP 2022 variant, data_aug.py:150-195) on log-mel features [B, F, T]: the per-(clip, bin) filter is drawn and
    assembled on the host exactly as upstream (same torch CPU RNG calls), the additive log-filter is applied by one kernel."""
    batch_size, n_freq_bin, n_frames = features.shape
    n_freq_band = torch.randint(low=n_band[0], high=n_band[1], size=(1, )).item()
    if n_freq_band <= 1:
        return features
    while n_freq_bin - n_freq_band * min_bw + 1 < 0:
        min_bw -= 1
    band_bndry_freqs = torch.sort(torch.randint(0, n_freq_bin - n_freq_band * min_bw + 1, (n_freq_band - 1, )))[0] + \
        torch.arange(1, n_freq_band) * min_bw
    band_bndry_freqs = torch.cat((torch.tensor([0]), band_bndry_freqs, torch.tensor([n_freq_bin])))
    if filter_type == "step":
        band_factors = torch.rand((batch_size, n_freq_band)) * (db_range[1] - db_range[0]) + db_range[0]
        band_factors = 10 ** (band_factors / 20)
        freq_filt = torch.ones((batch_size, n_freq_bin, 1))
        for i in range(n_freq_band):
            freq_filt[:, band_bndry_freqs[i]:band_bndry_freqs[i + 1], :] = band_factors[:, i].unsqueeze(-1).unsqueeze(-1)
    elif filter_type == "linear":
        band_factors = torch.rand((batch_size, n_freq_band + 1)) * (db_range[1] - db_range[0]) + db_range[0]
        freq_filt = torch.ones((batch_size, n_freq_bin, 1))
        for i in range(n_freq_band):
            for j in range(batch_size):
                freq_filt[j, band_bndry_freqs[i]:band_bndry_freqs[i + 1], :] = torch.linspace(
                    band_factors[j, i], band_factors[j, i + 1], band_bndry_freqs[i + 1] - band_bndry_freqs[i]).unsqueeze(-1)
    else:
        raise Exception("Unkonwn filter augment type")
    if not log:
        raise NotImplementedError("[DEBUG] Don't support filter augumentation after log operation")
    return _launch_add_rowbias(features, (torch.log(freq_filt + 0.00001) / norm_std).reshape(-1))


def _launch_add_rowbias(features, bias_host):
    """out[b, f, :] = features[b, f, :] + bias[b * F + f] -- one kernel over the batch."""
    _lib.ensure_device(features)
    x = features.contiguous().float()
    batch_size, n_freq_bin, n_frames = x.shape
    bias = bias_host.to(x.device)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().t4s_add_rowbias(_lib.ptr(x), _lib.ptr(bias), _lib.ptr(out), batch_size * n_freq_bin, n_frames, _lib.stream_ptr()),
                   "t4s_add_rowbias")
    return out


def feature_transformation(features, n_transform, choice, filter_db_range, filter_bands, filter_minimum_bandwidth, filter_type,
                           freq_mask_ratio=None, noise_snrs=None, norm_std=5, log=False):
    """The label-independent transform stack the recipes call every step (data_aug.py:111-147; e.g. recipes/desed/finetune/train.py,
    mlm_passt/train.py:32): `n_transform` independently augmented copies of the mel batch (student / teacher inputs).
    ``choice`` = [FilterAugment, freq_mask, add_noise, frequency distortion]; every shipped config uses [1, 0, 0, 1].  Host RNG draws
    are made in upstream's order (per copy: `random.random()` for the warp bias, one for its phase, then FilterAugment's torch
    draws).  The features stay on the GPU -- upstream's frequency distortion round-trips the batch through host numpy."""
    if choice[1] or choice[2]:
        raise NotImplementedError("freq_mask / add_noise are not selected by any shipped config and are not on the B200 path")
    feature_list = []
    for _ in range(n_transform):
        features_temp = features
        if choice[3]:
            bias = 0.03 * random.random()
            features_temp = freq_nonlinear(features_temp, bias=bias)
        if choice[0]:
            features_temp = filt_aug(features_temp, db_range=filter_db_range, n_band=filter_bands, min_bw=filter_minimum_bandwidth,
                                     filter_type=filter_type, norm_std=norm_std, log=log)
        if features_temp is features:            # upstream works on a deep copy: every returned tensor is distinct from the input
            features_temp = features.clone()
        feature_list.append(features_temp)
    if n_transform == 1:
        return feature_list[0]
    return feature_list
