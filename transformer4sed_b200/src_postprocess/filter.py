"""Drop-in for reference ``src/postprocess/filter.py`` (`median_filter_torch`): class-wise median filtering of the frame
probabilities before thresholding.  One kernel for all classes (the reference loops over classes with pad + unfold + median).

Upstream quirk kept by default: the reference loop is hard-wired to ``range(10)``, so with more than 10 classes the
remaining ones come back as zeros and with fewer it raises; ``strict_upstream=False`` filters every class.
"""
import ctypes

import torch

from .. import _lib


def median_filter_torch(input_tensor, filter_size: list, strict_upstream=True):
    if len(input_tensor.shape) != 3:
        raise ValueError("input_tensor must have shape (Batch, Length, Classes)")
    batch, length, num_classes = input_tensor.shape
    if len(filter_size) != num_classes:
        raise ValueError("Length of median_filter_sizes must match the number of classes")
    n_filtered = num_classes
    if strict_upstream:
        if num_classes < 10:
            raise IndexError("index 10 classes hard-coded upstream (src/postprocess/filter.py:25)")
        n_filtered = 10
    _lib.ensure_device(input_tensor)
    x = input_tensor.contiguous().float()
    sizes = [int(k) + 1 if int(k) % 2 == 0 else int(k) for k in filter_size]
    if n_filtered == num_classes:
        src, out = x, torch.empty_like(x)
    else:                                   # classes >= 10 stay zero upstream
        src = x[:, :, :n_filtered].contiguous()
        out = torch.empty_like(src)
    arr = (ctypes.c_int * n_filtered)(*sizes[:n_filtered])
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().t4s_median_filter(_lib.ptr(src), _lib.ptr(out), arr, batch, length, n_filtered, _lib.stream_ptr()),
                   "t4s_median_filter")
    if n_filtered != num_classes:
        full = torch.zeros_like(x)
        full[:, :, :n_filtered] = out
        out = full
    return out.to(input_tensor.dtype)
