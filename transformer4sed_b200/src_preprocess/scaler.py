"""Drop-in for reference ``src/preprocess/scaler.py`` (`TorchScaler`, :5-121): instance- or dataset-level normalisation of the
feature tensor.  Same constructor, `fit`, `forward`, state-dict behaviour; `forward` is one libt4s kernel (csrc/post.cu: one CTA
per instance, two reads + one write) instead of torch's mean / std / amin / amax passes.

Supported layout: statistics over every dimension but the first (the reference default ``dims=(1, 2)`` on [B, F, T] features, or
``dims=(1,)`` on [B, L] waveforms); any other `dims` raises NotImplementedError (no shipped config uses one)."""
import torch
import torch.nn as nn

from .. import _lib

_MODES = {"mean": 0, "standard": 1, "minmax": 2}


class TorchScaler(nn.Module):
    def __init__(self, statistic="dataset", normtype="standard", dims=(1, 2), eps=1e-8):
        super().__init__()
        assert statistic in ["dataset", "instance", None]
        assert normtype in ["standard", "mean", "minmax", None]
        if statistic == "dataset" and normtype == "minmax":
            raise NotImplementedError("statistic==dataset and normtype==minmax is not currently implemented.")
        self.statistic, self.normtype, self.dims, self.eps = statistic, normtype, tuple(dims), eps

    # the reference only restores buffers for dataset statistics (scaler.py:36-57)
    def load_state_dict(self, state_dict, strict=True):
        if self.statistic == "dataset":
            for k in ("mean", "mean_squared"):
                if k in state_dict and not hasattr(self, k):
                    self.register_buffer(k, torch.zeros_like(state_dict[k]))
            return super().load_state_dict(state_dict, strict)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if self.statistic == "dataset":
            super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def fit(self, dataloader, transform_func=lambda x: x[0]):
        """scaler.py:59-89: running mean of the per-batch mean and mean of squares (a one-off pass over the data set: torch ops)."""
        indx = 0
        mean = mean_squared = None
        for batch in dataloader:
            feats = transform_func(batch)
            m = torch.mean(feats, self.dims, keepdim=True).mean(0).unsqueeze(0)
            q = torch.mean(feats ** 2, self.dims, keepdim=True).mean(0).unsqueeze(0)
            mean = m if indx == 0 else mean + m
            mean_squared = q if indx == 0 else mean_squared + q
            indx += 1
        mean /= indx
        mean_squared /= indx
        self.register_buffer("mean", mean)
        self.register_buffer("mean_squared", mean_squared)

    def _check_dims(self, tensor):
        want = tuple(range(1, tensor.ndim))
        dims = tuple(d % tensor.ndim for d in self.dims)
        if tuple(sorted(dims)) != want:
            raise NotImplementedError(f"TorchScaler: statistics over dims {self.dims} of a {tensor.ndim}-d tensor are not implemented "
                                      f"(only every dimension but the first, e.g. dims=(1, 2) on [B, F, T])")

    def forward(self, tensor):
        if self.statistic is None or self.normtype is None:
            return tensor
        _lib.ensure_device(tensor)
        lib = _lib.load()
        x = tensor.contiguous().float()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            if self.statistic == "dataset":
                assert hasattr(self, "mean") and hasattr(self, "mean_squared"), "TorchScaler should be fit before used if statistics=dataset"
                assert tensor.ndim == self.mean.ndim, "Pre-computed statistics "
                if self.mean.numel() != 1:
                    raise NotImplementedError("TorchScaler: only scalar dataset statistics (dims = every dimension but the first) are implemented")
                mean = self.mean.to(x.device, torch.float32).reshape(1).contiguous()
                msq = self.mean_squared.to(x.device, torch.float32).reshape(1).contiguous()
                _lib.check(lib.t4s_scaler_dataset(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.ptr(mean), _lib.ptr(msq),
                                                  int(self.normtype == "standard"), float(self.eps), _lib.stream_ptr()), "t4s_scaler_dataset")
            else:
                self._check_dims(tensor)
                _lib.check(lib.t4s_scaler_instance(_lib.ptr(x), _lib.ptr(out), x.shape[0], x[0].numel(), _MODES[self.normtype], float(self.eps),
                                                   _lib.stream_ptr()), "t4s_scaler_instance")
        return out.to(tensor.dtype)
