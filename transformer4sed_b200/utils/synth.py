"""Synthetic, seed-reproducible checkpoints and inputs.

There is no network for the pretrained PaSST checkpoint the reference loads
(reference src/models/passt/passt_sed.py:114), so parity tests, the golden-vector generator and
bench.py all build weights from a seed.  The generator depends only on (seed, key name, shape),
never on module construction order, so the unmodified reference, the CPU oracle and the CUDA
modules can be handed bit-identical tensors through ``load_state_dict``.

Scales are chosen so activations are "realistic" (softmax far from uniform, LayerNorm gains
away from 1) rather than the std=.02 default init, which would hide most numerical bugs.
"""
import zlib

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    return g


def synth_tensor(seed: int, name: str, shape, kind: str = None) -> torch.Tensor:
    """One fp32 CPU tensor for state-dict key ``name``."""
    shape = tuple(shape)
    g = _gen(seed, name)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = name.rsplit(".", 1)[-1]
    if kind is None:
        if "norm" in name and leaf == "weight" and len(shape) == 1:
            kind = "ln_w"
        elif leaf in ("bias", "in_proj_bias") and len(shape) == 1:
            kind = "bias"
        elif len(shape) >= 2 and leaf in ("weight", "in_proj_weight"):
            kind = "matrix"
        else:
            kind = "embed"
    if kind == "ln_w":
        return 1.0 + 0.1 * r
    if kind == "bias":
        return 0.05 * r
    if kind == "matrix":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r * (1.0 / fan_in) ** 0.5
    return 0.3 * r  # tokens, positional embeddings, pos_bias_u/v, mask_token, query tokens


def synth_state_dict(shapes: dict, seed: int) -> dict:
    """``shapes``: {key: shape}.  Returns {key: fp32 tensor}."""
    out = {}
    for k, s in sorted(shapes.items()):
        if k.endswith("running_var"):        # same rules as synth_state_dict_like
            out[k] = 0.5 + torch.rand(tuple(s), generator=_gen(seed, k))
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(tuple(s), generator=_gen(seed, k))
        else:
            out[k] = synth_tensor(seed, k, s)
    return out


def synth_state_dict_like(module: torch.nn.Module, seed: int) -> dict:
    """Seeded replacement for every floating-point entry of ``module.state_dict()``
    (BatchNorm running_var stays positive, integer buffers are left alone)."""
    out = {}
    for k, v in module.state_dict().items():
        if not torch.is_floating_point(v):
            out[k] = v.clone()
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(v.shape, generator=_gen(seed, k))
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(v.shape, generator=_gen(seed, k))
        elif "bn" in k.lower() and k.endswith("weight") and v.ndim == 1:
            out[k] = synth_tensor(seed, k, v.shape, "ln_w")
        else:
            out[k] = synth_tensor(seed, k, v.shape)
    return out


def synth_wav(batch: int, n_samples: int, seed: int, device="cpu") -> torch.Tensor:
    """``0.1*randn`` plus a few chirps/tones per clip so the mel image has
    structure across >6 decades (pure noise would keep log-mel in a 1-decade band)."""
    g = _gen(seed, "wav")
    x = 0.1 * torch.randn(batch, n_samples, generator=g)
    t = torch.arange(n_samples, dtype=torch.float32) / 32000.0
    for b in range(batch):
        for j in range(3):
            f0 = 100.0 + 3000.0 * torch.rand((), generator=g).item() * (j + 1)
            sweep = 2000.0 * (torch.rand((), generator=g).item() - 0.5)
            amp = 10 ** (-2.0 * torch.rand((), generator=g).item())
            on = int(torch.rand((), generator=g).item() * n_samples * 0.7)
            off = min(n_samples, on + int((0.05 + 0.3 * torch.rand((), generator=g).item()) * n_samples))
            ph = 2 * torch.pi * (f0 * t[on:off] + 0.5 * sweep * t[on:off] ** 2 / max(t[-1].item(), 1e-3))
            x[b, on:off] += amp * torch.sin(ph)
    return x.to(device)


def synth_strong_labels(batch: int, n_class: int, n_frames: int, seed: int) -> torch.Tensor:
    """per (clip, class) with p=0.2 one interval of 0.5-5 s (50-500 frames at 100 fps)."""
    g = _gen(seed, "labels")
    y = torch.zeros(batch, n_class, n_frames)
    for b in range(batch):
        for c in range(n_class):
            if torch.rand((), generator=g).item() < 0.2:
                ln = int(n_frames * (0.05 + 0.45 * torch.rand((), generator=g).item()))
                st = int(torch.rand((), generator=g).item() * max(1, n_frames - ln))
                y[b, c, st:st + ln] = 1.0
    return y
