"""Drop-in for reference ``src/models/passt/passt_lora.py``: the PaSST backbone with every block Linear (qkv, proj, fc1, fc2) and
the classification heads replaced by `lora.Linear` (:46-48, :122-124, :288-299).  Parameter names match the reference
(``blocks.N.attn.qkv.lora_A`` ...), so PMAM / DASM checkpoints load with strict=True."""
import torch.nn as nn

from ... import functional as F
from ... import ops
from .. import lora
from . import passt as base


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0., lora_config=dict()):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU or drop:
            raise NotImplementedError("libt4s fuses exact-erf GELU only; dropout is 0 in every shipped recipe")
        self.fc1 = lora.Linear(in_features, hidden_features, **lora_config)
        self.act = act_layer()
        self.fc2 = lora.Linear(hidden_features, out_features, **lora_config)
        self.drop = nn.Dropout(drop)

    def forward(self, x, residual=None):
        return self.fc2(self.fc1(x, act=ops.ACT_GELU), residual=residual)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0., lora_config=dict()):
        super().__init__()
        if attn_drop or proj_drop:
            raise NotImplementedError("dropout is 0 in every shipped recipe and is not implemented")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = lora.Linear(dim, dim * 3, bias=qkv_bias, **lora_config)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = lora.Linear(dim, dim, **lora_config)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, att_mask=None, residual=None):
        if att_mask is not None:
            raise NotImplementedError("att_mask is unused by the shipped recipes")
        return self.proj(F.attention(self.qkv(x), self.num_heads), residual=residual)


class Block(base.Block):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, lora_config=dict()):
        nn.Module.__init__(self)
        if drop_path:
            raise NotImplementedError("drop_path is 0 in every shipped recipe and is not implemented")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop, lora_config=lora_config)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop, lora_config=lora_config)


class PaSST(base.PaSST):
    def __init__(self, *args, lora_config=dict(), **kwargs):
        self._lora_config = dict(lora_config)
        super().__init__(*args, **kwargs)

    def make_block(self, **kw):
        return Block(lora_config=self._lora_config, **kw)

    def make_head_linear(self, in_features, out_features):
        return lora.Linear(in_features, out_features, **self._lora_config)
