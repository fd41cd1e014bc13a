"""Drop-in for the label-coupled augmentations of reference ``src/preprocess/data_aug.py`` that run every training step on the
mel batch: `frame_shift` (:12-31) and `mixup` (:34-91).  The random draws are made on the host with the same generators, in the
same order, as the reference (python `random`, `torch.randperm`, `np.random.beta`), so seeded runs pick the same shifts,
permutation and mixing rate; the data movement is one libt4s kernel per tensor instead of a Python loop of `torch.roll` + stack.
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
