"""Drop-in for reference ``src/models/transformer/mask.py``: `MlmModule` (block / random frame masking, :49-107).

Mask *sampling* consumes the same torch RNG draws in the same order as the reference so seeded runs select the same
frames.  Upstream quirk kept (refined by oracle/make_golden.py): in `PaSST_SED` the token replacement
writes through ``clone().reshape(-1, C)`` of a non-contiguous tensor, which is a copy whenever B > 1, so the decoder
input is left unchanged (only the loss uses the mask).  `apply=False` reproduces that; `apply=True` performs the
replacement (what `PaSST_CNN`, whose input is contiguous, and B == 1 do upstream).
"""
import torch


class MlmModule:
    def __init__(self, mask_rate=0.15, mask_style=(0.8, 0.1, 0.1), strategy="random", block_width=10, device=None, *arg, **kwarg) -> None:
        self.mask_rate = mask_rate
        self.mask_style = {"mask_token": mask_style[0], "random": mask_style[1], "self": mask_style[2]}
        self.strategy = strategy
        self.device = device
        self.block_width = block_width

    def setence_mask(self, token_seq, mask_token, apply=True):
        B, T, C = token_seq.shape
        dev = token_seq.device
        rng_dev = self.device if self.device is not None else dev   # the reference draws on `self.device` (mask.py:71,79,88,96)
        mask_id_seq = self.get_mask_id_seq(B, T, rng_dev)
        mask_id_flat = mask_id_seq.view(-1)
        probs = torch.rand(B * T, device=rng_dev)
        mask_mask = mask_id_flat & (probs < self.mask_style["mask_token"])
        random_mask = mask_id_flat & (probs >= self.mask_style["mask_token"]) & (
            probs < self.mask_style["mask_token"] + self.mask_style["random"])
        random_indices = torch.randint(0, B * T, (int(random_mask.sum().item()),), device=rng_dev)
        mask_id_seq = mask_id_seq.to(dev)
        if not apply:
            return token_seq, mask_id_seq
        from ... import functional as F
        return F.mask_rows(token_seq, mask_token, mask_mask.to(dev), random_mask.to(dev), random_indices.to(dev)), mask_id_seq

    def get_mask_id_seq(self, batch_len, seq_len, device=None):
        if self.strategy == "random":
            return self.random_mask(batch_len, seq_len, device)
        if self.strategy == "block":
            return self.block_mask(batch_len, seq_len, self.block_width, device)
        raise ValueError("Unknown mask strategy")

    def random_mask(self, batch_len, seq_len, device=None):
        noise = torch.rand(batch_len, seq_len, device=device or self.device)
        return noise <= self.mask_rate

    def block_mask(self, batch_len, seq_len, block_width=10, device=None):
        device = device or self.device
        num_seg = seq_len // block_width
        noise = torch.rand(batch_len, num_seg, device=device)
        noise_sort, _ = noise.sort()
        threshold = noise_sort[:, min(int(num_seg * self.mask_rate), num_seg - 1)]
        id_seq = torch.zeros(batch_len, seq_len, dtype=bool, device=device)
        id_seq[:, :num_seg * block_width] = (noise <= torch.unsqueeze(threshold, dim=-1)).repeat_interleave(block_width, dim=1)
        return id_seq
