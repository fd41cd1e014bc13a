"""Drop-in for the decoding arithmetic of reference ``src/codec/decoder.py`` (:15-103) and ``src/codec/encoder.py:51-84``: the
threshold sweep that turns frame probabilities into (event_label, onset, offset) rows, and the per-class score filtering that
feeds the PSDS evaluation.  Upstream this is a Python loop over thresholds x clips x classes with a host round trip per threshold;
here the class-wise filter runs once and ONE pair of kernel launches (count, fill) decodes every threshold (csrc/post.cu).

`decode_events` is the device core (frame indices, bit-exact); `decode_pred_batch_fast` / `batched_decode_preds` wrap it into the
pandas objects the recipes consume (`encoder` is the reference's ManyHotEncoder: only `labels` and `_frame_to_time` are used)."""
import ctypes
from pathlib import Path

import numpy as np
import torch

from .. import _lib
from ..src_postprocess.filter import median_filter_torch


def decode_events(scores, weak_preds, thresholds):
    """scores [B, T, C] (filtered), weak_preds [B, C] or None -> int32 array [n, 5] of (threshold index, clip, class, onset frame,
    offset frame), ordered by threshold, clip, class, onset (the order decode_pred_batch_fast appends rows in)."""
    _lib.ensure_device(scores)
    lib = _lib.load()
    x = scores.contiguous().float()
    B, T, C = x.shape
    th = torch.as_tensor(list(thresholds), dtype=torch.float32, device=x.device)
    n_th = th.numel()
    w = weak_preds.contiguous().float() if weak_preds is not None else None
    with torch.cuda.device(x.device):
        counts = torch.empty(n_th * B * C, dtype=torch.int32, device=x.device)
        _lib.check(lib.t4s_event_sweep(_lib.ptr(x), _lib.ptr(w), _lib.ptr(th), n_th, B, T, C, _lib.ptr(counts), None, None, _lib.stream_ptr()),
                   "t4s_event_sweep(count)")
        ends = torch.cumsum(counts.to(torch.int64), 0)
        total = int(ends[-1].item())
        offsets = (ends - counts).contiguous()
        events = torch.empty(max(total, 1), 5, dtype=torch.int32, device=x.device)
        if total:
            _lib.check(lib.t4s_event_sweep(_lib.ptr(x), _lib.ptr(w), _lib.ptr(th), n_th, B, T, C, None, _lib.ptr(offsets), _lib.ptr(events),
                                           _lib.stream_ptr()), "t4s_event_sweep(fill)")
    return events[:total].cpu().numpy()


def decode_pred_batch_fast(outputs, weak_preds, filenames, encoder, thresholds, median_filter, strict_upstream=True):
    """decoder.py:15-35.  outputs [B, C, T] strong probabilities, weak_preds [B, C]; returns {threshold: DataFrame(event_label, onset,
    offset, filename)}.  Zeroing the classes with weak < threshold before the median filter (upstream) equals silencing those
    columns after it, so the filter runs once for all thresholds."""
    import pandas as pd
    thresholds = list(thresholds)
    filt = median_filter_torch(outputs.transpose(1, 2).detach(), median_filter, strict_upstream=strict_upstream)
    ev = decode_events(filt, weak_preds.detach(), thresholds)
    labels = list(encoder.labels)
    stems = [Path(f).stem + ".wav" for f in filenames]
    onset = np.clip(encoder._frame_to_time(ev[:, 3]), a_min=0, a_max=encoder.audio_len) if len(ev) else np.zeros(0)
    offset = np.clip(encoder._frame_to_time(ev[:, 4]), a_min=0, a_max=encoder.audio_len) if len(ev) else np.zeros(0)
    pred_dfs = {}
    for ti, th in enumerate(thresholds):
        sel = np.nonzero(ev[:, 0] == ti)[0] if len(ev) else np.zeros(0, dtype=np.int64)
        pred_dfs[th] = pd.DataFrame({"event_label": [labels[c] for c in ev[sel, 2]], "onset": onset[sel], "offset": offset[sel],
                                     "filename": [stems[b] for b in ev[sel, 1]]})
    return pred_dfs


def _score_dataframe(scores, timestamps, event_classes):
    """sed_scores_eval.base_modules.scores.create_score_dataframe: columns onset, offset, then one column per class."""
    try:
        from sed_scores_eval.base_modules.scores import create_score_dataframe
        return create_score_dataframe(scores=scores, timestamps=timestamps, event_classes=event_classes)
    except ImportError:
        import pandas as pd
        return pd.DataFrame(np.concatenate((timestamps[:-1, None], timestamps[1:, None], scores), axis=1),
                            columns=["onset", "offset", *event_classes])


def filter_scores(strong_preds, filter=7, filter_type="median", weak_preds=None, need_weak_mask=None):
    """The device part of batched_decode_preds (decoder.py:61-95): optional soft weak mask, then scipy.ndimage median / maximum
    filter per class.  strong_preds [B, C, T] -> (raw, post-processed) float32 [B, T, C]."""
    _lib.ensure_device(strong_preds)
    lib = _lib.load()
    x = strong_preds.detach().transpose(1, 2).contiguous().float()
    if need_weak_mask and weak_preds is not None:
        x = x * weak_preds.detach().float().unsqueeze(1)
    if not filter:
        return x, x
    B, T, C = x.shape
    sizes = list(filter)
    if len(sizes) > C:
        raise IndexError("more filter sizes than classes")
    full = sizes + [1] * (C - len(sizes))       # upstream filters the first len(filter) classes only
    out = torch.empty_like(x)
    arr = (ctypes.c_int * C)(*[int(k) for k in full])
    with torch.cuda.device(x.device):
        _lib.check(lib.t4s_rank_filter(_lib.ptr(x), _lib.ptr(out), arr, B, T, C, {"median": 0, "max": 1}[filter_type], _lib.stream_ptr()),
                   "t4s_rank_filter")
    return x, out


def batched_decode_preds(strong_preds, filenames, encoder, filter=7, filter_type="median", pad_indx=None, weak_preds=None, need_weak_mask=None):
    """decoder.py:38-103: {audio_id: score DataFrame} before and after the class-wise filter.  (Upstream's `pad_indx` slice acts on
    the class axis, `c_scores[:true_len]` of a [n_class, frame] array; every shipped recipe passes None, and so must callers here.)"""
    if pad_indx is not None:
        raise NotImplementedError("pad_indx is not used by any shipped recipe (upstream slices the class axis with it)")
    raw, post = filter_scores(strong_preds, filter, filter_type, weak_preds, need_weak_mask)
    raw, post = raw.cpu().numpy(), post.cpu().numpy()
    scores_raw, scores_post = {}, {}
    for j in range(raw.shape[0]):
        audio_id = Path(filenames[j]).stem
        ts = encoder._frame_to_time(np.arange(raw.shape[1] + 1))
        scores_raw[audio_id] = _score_dataframe(raw[j], ts, encoder.labels)
        scores_post[audio_id] = _score_dataframe(post[j], ts, encoder.labels) if filter else scores_raw[audio_id]
    return scores_raw, scores_post
