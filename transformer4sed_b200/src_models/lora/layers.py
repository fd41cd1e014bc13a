"""Drop-in for the parts of reference ``src/models/lora/layers.py`` the PMAM / DASM backbones use: `LoRALayer`, `Linear`
(:13-153), plus the two helpers of ``src/models/lora/utils.py`` recipes call.

Same parameters (`weight`, `bias`, `lora_A [r, in]`, `lora_B [out, r]`), same init, same train()/eval() weight merging
(`merge_weights`: eval folds ``B A * alpha/r`` into `weight`, train takes it out again).  The arithmetic is
`functional.lora_linear`: the rank-r update rides inside the ordinary tcgen05 GEMM as an effective weight, and its gradient
goes through the two [tokens, r] factors only.
"""
import math

import torch.nn as nn

from ... import functional as F
from ... import ops


class LoRALayer:
    def __init__(self, r: int, lora_alpha: int, lora_dropout: float, merge_weights: bool):
        self.r = r
        self.lora_alpha = lora_alpha
        if lora_dropout > 0.:
            raise NotImplementedError("lora_dropout is 0 in every shipped config and is not implemented")
        self.lora_dropout = lambda x: x
        self.merged = False
        self.merge_weights = merge_weights


class Linear(nn.Linear, LoRALayer):
    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1, lora_dropout: float = 0.,
                 fan_in_fan_out: bool = False, merge_weights: bool = True, requires_grad_pretrain: bool = False, **kwargs):
        nn.Linear.__init__(self, in_features, out_features, **kwargs)
        LoRALayer.__init__(self, r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout, merge_weights=merge_weights)
        if fan_in_fan_out:
            raise NotImplementedError("fan_in_fan_out layers are not used by the SED models")
        self.fan_in_fan_out = fan_in_fan_out
        if r > 0:
            self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = requires_grad_pretrain
        self.reset_parameters()

    def reset_parameters(self):
        nn.Linear.reset_parameters(self)
        if hasattr(self, 'lora_A'):
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def train(self, mode: bool = True):
        nn.Linear.train(self, mode)
        if self.r > 0 and self.merge_weights and (self.merged == mode):
            # train: un-merge; eval: merge  (layers.py:124-141).  A [out, in] x rank-r outer product once per mode switch.
            delta = (self.lora_B.data @ self.lora_A.data) * self.scaling
            self.weight.data += delta if not mode else -delta
            F.invalidate_weight_cache(self.weight)
            self.merged = not mode
        return self

    def forward(self, x, residual=None, act=ops.ACT_NONE):
        if self.r > 0 and not self.merged:
            return F.lora_linear(x, self.weight, self.bias, self.lora_A, self.lora_B, self.scaling, residual=residual, act=act)
        return F.linear(x, self.weight, self.bias, residual=residual, act=act)


def mark_only_lora_as_trainable(model: nn.Module, bias: str = 'none') -> None:
    for n, p in model.named_parameters():
        if 'lora_' not in n:
            p.requires_grad = False
    if bias == 'all':
        for n, p in model.named_parameters():
            if 'bias' in n:
                p.requires_grad = True
    elif bias == 'lora_only':
        for m in model.modules():
            if isinstance(m, LoRALayer) and hasattr(m, 'bias') and m.bias is not None:
                m.bias.requires_grad = True


def lora_state_dict(model: nn.Module, bias: str = 'none'):
    sd = model.state_dict()
    if bias == 'none':
        return {k: sd[k] for k in sd if 'lora_' in k}
    if bias == 'all':
        return {k: sd[k] for k in sd if 'lora_' in k or 'bias' in k}
    raise NotImplementedError
