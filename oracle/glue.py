"""Oracle: train-loop glue around the model (CPU, torch / python).  Test infrastructure only (see oracle/__init__.py).

frame_shift / mixup follow reference src/preprocess/data_aug.py:12-31 / :34-91 with the random draws injected;
median_filter follows reference src/postprocess/filter.py:4-36 (including its hard-wired range(10))."""
import torch


def frame_shift(features, label, net_pooling, shifts):
    """data_aug.py:12-23: torch.roll per sample; the label shift is floor-divided by net_pooling towards -inf for negative shifts."""
    f = torch.stack([torch.roll(features[i], s, dims=-1) for i, s in enumerate(shifts)])
    lab = torch.stack([torch.roll(label[i], int(-abs(s) // net_pooling if s < 0 else s // net_pooling), dims=-1) for i, s in enumerate(shifts)])
    return f, lab


def mixup(features, label, permutation, c, mixup_label_type="soft"):
    """data_aug.py:74-89."""
    mf = c * features + (1 - c) * features[permutation, :]
    if mixup_label_type == "soft":
        ml = torch.clamp(c * label + (1 - c) * label[permutation, :], min=0, max=1)
    else:
        ml = torch.clamp(label + label[permutation, :], min=0, max=1)
    return mf, ml


def median_filter(x, filter_size, n_classes_filtered=10):
    """filter.py:24-35: per class, window made odd, replicate padding, exact median; classes >= 10 stay zero upstream."""
    B, L, C = x.shape
    out = torch.zeros_like(x)
    for c in range(n_classes_filtered):
        k = filter_size[c] + 1 if filter_size[c] % 2 == 0 else filter_size[c]
        xi = torch.nn.functional.pad(x[:, :, c].unsqueeze(1), (k // 2, k // 2), mode="replicate").squeeze(1)
        out[:, :, c] = xi.unfold(1, k, 1).median(dim=-1)[0]
    return out


def freq_nonlinear(mel, phase, f=1, bias=0.02):
    """data_aug.py:239-254 with the single `random.random()` draw injected: np.interp of every (clip, frame) column over the warped
    frequency knots (numpy float64 arithmetic, result stored as float32)."""
    import numpy as np
    m = mel.numpy().copy()
    B, F, T = m.shape
    m = np.reshape(np.transpose(m, (0, 2, 1)), (B * T, F))
    ind = np.arange(F)
    x = ind / F
    ind_t = F * (x + bias * np.sin(2 * np.pi * (f * x + phase)))
    for i in range(B * T):
        m[i, :] = np.interp(ind, ind_t, m[i, :])
    return torch.from_numpy(np.reshape(m, (B, T, F)).transpose((0, 2, 1)).copy())


def filt_aug_apply(features, freq_filt, norm_std):
    """data_aug.py:188-190 (log features): features + log(filter + 1e-5) / norm_std with filter [B, F, 1]."""
    return features + torch.log(freq_filt + 0.00001) / norm_std
