"""Drop-in for reference ``src/models/transformer/transformerXL.py``: `TransformerXL` block,
`RelPositionalEncoding`, `RelPositionMultiheadAttention` (reference :23-593).

Upstream semantics kept bit-for-bit in structure: the block's residual is taken from the *normalised* input
(``x = norm1(x); x = x + attn(x)``, :31-35), scores are ``((q+u)k^T + shift((q+v)p^T)) / sqrt(hd)`` with
``p = linear_pos(pos_emb)`` (no bias), and row k of the position table holds relative position T-1-k.
Activations are stored batch-major [B, T, C]; a block takes and returns the reference's (T, B, C) layout as a strided VIEW of
that storage (no copy), so forward hooks on ``decoder.encoder_blocks[i]`` see what they see upstream
(recipes/desed/pmam/extractor_feature.py:83-89 does ``fea_out.transpose(0, 1).reshape(-1, C)``).
"""
import math

import torch
import torch.nn as nn

from ... import functional as F
from ... import ops
from ..passt.passt import Mlp


class RelPositionalEncoding(nn.Module):
    def __init__(self, d_model: int, dropout_rate: float, max_len: int = 5000) -> None:
        super().__init__()
        if dropout_rate:
            raise NotImplementedError("dropout is 0 in every shipped recipe and is not implemented")
        self.d_model = d_model
        self.xscale = math.sqrt(d_model)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.pe = None

    def forward(self, x, left_context: int = 0):
        """Returns (x, pos_emb [2T-1, C]).  The x*sqrt(d) scale (:118) is applied inside the first LayerNorm kernel of each
        consumer (`in_scale`), so x is returned unscaled together with `self.xscale`."""
        if left_context:
            raise NotImplementedError("streaming left_context is unused by the SED recipes")
        T = x.shape[1]
        self.pe = F.rel_pos_table(T, self.d_model, x.device, x.dtype)
        return x, self.pe


class RelPositionMultiheadAttention(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0) -> None:
        super().__init__()
        if dropout:
            raise NotImplementedError("attention dropout is 0 in every shipped recipe and is not implemented")
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.in_proj = nn.Linear(embed_dim, 3 * embed_dim, bias=True)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.linear_pos = nn.Linear(embed_dim, embed_dim, bias=False)
        self.pos_bias_u = nn.Parameter(torch.Tensor(num_heads, self.head_dim))
        self.pos_bias_v = nn.Parameter(torch.Tensor(num_heads, self.head_dim))
        self._reset_parameters()

    def _reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.in_proj.weight)
        nn.init.constant_(self.in_proj.bias, 0.0)
        nn.init.constant_(self.out_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)

    def forward(self, x, pos_emb, residual=None):
        """x [B, T, C] (query = key = value), pos_emb [2T-1, C] -> [B, T, C]"""
        qkv = F.linear(x, self.in_proj.weight, self.in_proj.bias)
        p = F.linear(pos_emb, self.linear_pos.weight)  # batch independent: once per layer, not per clip
        o = F.relpos_attention(qkv, p, self.pos_bias_u, self.pos_bias_v, self.num_heads)
        return F.linear(o, self.out_proj.weight, self.out_proj.bias, residual=residual)


class TransformerXL(nn.Module):
    """timm-0.4.5 `Block` members (norm1, attn, drop_path, norm2, mlp) with the relative-position attention (:23-35)."""

    def __init__(self, dim, num_heads, mlp_ratio=4, qkv_bias=False, qk_scale=None, drop=0, attn_drop=0, drop_path=0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if drop or drop_path:
            raise NotImplementedError("dropout / drop_path are 0 in every shipped recipe and are not implemented")
        self.norm1 = norm_layer(dim)
        self.attn = RelPositionMultiheadAttention(embed_dim=dim, num_heads=num_heads, dropout=attn_drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, pos_emb, att_mask=None, in_scale=1.0):
        """x (T, B, C) -> (T, B, C), both views of batch-major storage (a genuinely time-major tensor is copied once)."""
        if att_mask is not None:
            raise NotImplementedError("decoder_win_len attention masks are unused by the shipped configs")
        x = x.transpose(0, 1)
        x = F.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, in_scale=in_scale)
        x = self.attn(x, pos_emb, residual=x)           # residual from the NORMALISED input (upstream semantics)
        y, res = F.layer_norm_res(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return self.mlp(y, residual=res).transpose(0, 1)
