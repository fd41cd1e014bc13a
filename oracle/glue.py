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


# ---- callers either side of the hot path: scaler, decoding, mean-teacher losses (SURVEY §8 a1', f3, f4) -----------------------------
def torch_scaler(x, statistic, normtype, dims=(1, 2), eps=1e-8, mean=None, mean_squared=None):
    """src/preprocess/scaler.py:91-121."""
    if statistic is None or normtype is None:
        return x
    if statistic == "dataset":
        if normtype == "mean":
            return x - mean
        std = torch.sqrt(mean_squared - mean ** 2)
        return (x - mean) / (std + eps)
    if normtype == "mean":
        return x - torch.mean(x, dims, keepdim=True)
    if normtype == "standard":
        return (x - torch.mean(x, dims, keepdim=True)) / (torch.std(x, dims, keepdim=True) + eps)
    lo, hi = torch.amin(x, dim=dims, keepdim=True), torch.amax(x, dim=dims, keepdim=True)
    return (x - lo) / (hi - lo + eps)


def find_contiguous_regions(array):
    """src/codec/encoder.py:71-84."""
    import numpy as np
    change = np.logical_xor(array[1:], array[:-1]).nonzero()[0] + 1
    if array[0]:
        change = np.r_[0, change]
    if array[-1]:
        change = np.r_[change, array.size]
    return change.reshape((-1, 2))


def decode_pred_batch_fast(outputs, weak_preds, thresholds, filter_size):
    """src/codec/decoder.py:15-35 down to frame indices: rows (threshold index, clip, class, onset frame, offset frame) in the
    order the reference appends them (threshold, clip, class, onset)."""
    import numpy as np
    rows = []
    for ti, th in enumerate(thresholds):
        out = outputs.transpose(1, 2).clone()
        b_idx, c_idx = torch.where(weak_preds < th)
        out[b_idx, :, c_idx] = 0
        out = (median_filter(out, filter_size) > th).float().numpy()
        for b in range(out.shape[0]):
            for c, col in enumerate(out[b].T):
                for on, off in find_contiguous_regions(col):
                    rows.append((ti, b, c, int(on), int(off)))
    return np.array(rows, dtype=np.int32).reshape(-1, 5)


def rank_filter_scores(scores, sizes, filter_type="median"):
    """src/codec/decoder.py:86-92: scipy.ndimage filters per class on a [T, C] score array (first len(sizes) classes).
    For windows of more than 2 T + 1 taps scipy's 1-D fast path (taken here, as upstream) and its N-D filter disagree about the
    border; the CUDA kernel follows the N-D (periodic reflection) rule there -- tests/test_post_gpu.py pins both regimes."""
    from scipy import ndimage
    out = scores.copy()
    for idx in range(len(sizes)):
        if filter_type == "median":
            out[:, idx] = ndimage.median_filter(scores[:, idx], (sizes[idx]))
        else:
            out[:, idx] = ndimage.maximum_filter(scores[:, idx], (sizes[idx]))
    return out


def sed_losses(stu_strong, stu_weak, stu_at, tch_strong, tch_at, labels, labels_weak, mask_strong, mask_weak, w_weak, w_at, w_cons, w_weak_cons):
    """recipes/desed/finetune/train.py:166-188 (BCELoss / MSELoss, mean reduction).  Returns (total, [six parts])."""
    bce, mse = torch.nn.BCELoss(), torch.nn.MSELoss()
    l_at = bce(stu_at[mask_weak], labels_weak[mask_weak])
    c_at = mse(stu_at, tch_at.detach())
    l_strong = bce(stu_strong[mask_strong], labels[mask_strong])
    l_weak = bce(stu_weak[mask_weak], labels_weak[mask_weak])
    c_strong = mse(stu_strong, tch_strong.detach())
    c_weak = mse(stu_weak, tch_at.detach())
    self_loss = (c_strong + w_weak_cons * c_weak + w_at * c_at) * w_cons
    total = l_strong + w_weak * l_weak + self_loss + l_at * w_at
    return total, [l_strong, l_weak, l_at, c_strong, c_weak, c_at]
