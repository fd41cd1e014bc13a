"""Drop-in for reference ``src/models/encoder_slide_window.py`` (`EncoderSlideWindow`): global-local sliding-window fusion.

The reference walks the windows in a Python loop -- one full backbone pass per window (11 with ``win_param=[512, 49]``,
17 with ``[512, 31]``), accumulating into two fp32 buffers and dividing.  Here windows that yield the same patch grid are
folded into the batch dimension (`encode_windows`), so the whole thing is one or two backbone passes over 602-token
sequences, and a single overlap-add kernel (csrc/window.cu) produces the averaged embedding.  Window placement, the
`round(w_left * scale)` output offset, the shorter last window and the zero fill where no window lands follow
encoder_slide_window.py:16-36 exactly.
"""
from abc import ABC, abstractmethod

import torch

from .. import functional as F


class EncoderSlideWindow(ABC):
    #: upper bound on sequences per backbone pass (window-major chunks); None = all windows of a group at once
    max_sequences = None

    def __init__(self, net, win_param=[512, 31], out_dim=768):
        super().__init__()
        self.net = net
        self.out_dim = out_dim
        self.win = win_param

    def window_starts(self, input_len):
        """[(w_left, width)] exactly as the reference loop enumerates them (encoder_slide_window.py:29-30)."""
        win_width, step = self.win
        return [(w_left, min(w_left + win_width, input_len) - w_left) for w_left in range(0, input_len + step - win_width, step)]

    def __call__(self, input: torch.Tensor, emb_len):
        batch_size, _, input_len = input.shape
        scale = emb_len / input_len
        wins = self.window_starts(input_len)
        if not wins:
            return torch.zeros(batch_size, emb_len, self.out_dim, dtype=F.act_dtype(), device=input.device)
        # group consecutive windows that produce the same patch grid (all but possibly the last, which is cut at the clip end)
        groups = []
        for w_left, width in wins:
            key = self.frames_per_window(width)
            if groups and groups[-1][0] == key:
                groups[-1][1].append(w_left)
            else:
                groups.append((key, [w_left], width))
        outs = []
        for key, starts, width in groups:
            chunk = len(starts)
            if self.max_sequences:
                chunk = max(1, min(chunk, self.max_sequences // max(1, batch_size)))
            for i in range(0, len(starts), chunk):
                part = starts[i:i + chunk]
                local = self.encode_windows(input, part, width)          # [(w, b), frames, C]
                outs.append((local, [round(s * scale) for s in part]))
        return F.window_overlap_add(outs, batch_size, emb_len)

    def frames_per_window(self, width):
        """Grouping key: windows with equal keys must produce equally shaped `encode_windows` outputs."""
        return width

    def encode_windows(self, input, starts, width):
        """Default: one `encode` per window (subclasses fold the windows into the batch instead)."""
        return torch.cat([self.encode(input[:, :, s:s + width]) for s in starts], dim=0)

    @abstractmethod
    def encode(self, input: torch.Tensor) -> torch.Tensor:
        pass
