"""Drop-in for reference ``src/models/pooling.py`` (`AttentionPooling`, the only pooler the shipped configs select)."""
import torch
import torch.nn as nn

from .. import functional as F


class AttentionPooling(nn.Module):
    """One learned query token attending over the keys via nn.MultiheadAttention semantics (pooling.py:37-51).
    `frequency_att` is kept as an `nn.MultiheadAttention` parameter holder so state-dict keys match
    (``f_att_token``, ``frequency_att.in_proj_weight`` ...); the arithmetic is GEMM + the attnpool kernel."""

    def __init__(self, embed_dim, num_head=4):
        super().__init__()
        self.f_att_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        torch.nn.init.normal_(self.f_att_token, std=.02)
        self.frequency_att = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_head, batch_first=True)
        self.num_head = num_head

    def forward(self, x, skip=0):
        """x [items, keys, C] -> [items, C]; the first `skip` keys of every item are not attended to."""
        att = self.frequency_att
        return F.mha_pool(F.to_act(x), self.f_att_token, att.in_proj_weight, att.in_proj_bias, att.out_proj.weight,
                          att.out_proj.bias, self.num_head, skip=skip)
