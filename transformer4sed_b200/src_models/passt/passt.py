"""Drop-in for the parts of reference ``src/models/passt/passt.py`` the SED recipes use: `PaSST`, `Block`, `Attention`,
`Mlp`, `PatchEmbed` (reference :257-596).

Parameter names, shapes and submodule names match the reference so checkpoints and the recipes' parameter-group regexes
(``blocks.(\\d+)``, ``norm.``; reference recipes/desed/finetune/passt/setting.py:43-83) keep working.  The modules are
parameter holders; the arithmetic is libt4s kernels through `transformer4sed_b200.functional`:
  patch conv            -> im2col + tcgen05 GEMM with bias / positional tables fused in the epilogue
  LN -> qkv             -> LayerNorm kernel, GEMM(+bias)
  softmax(q k^T/8) v    -> batched GEMMs reading q/k/v in place from the fused qkv buffer + row softmax
  proj (+residual)      -> GEMM epilogue adds bias and the residual stream
  fc1 + GELU, fc2 (+residual) -> GEMM epilogues (exact-erf GELU; pre-activation kept for backward)
"""
import math
import warnings
from functools import partial

import torch
import torch.nn as nn

from ... import functional as F
from ... import ops


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("libt4s fuses exact-erf GELU only")
        if drop:
            raise NotImplementedError("dropout is 0 in every shipped recipe and is not implemented")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x, residual=None):
        return F.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual=residual)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, stride=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        img_size, patch_size, stride = to_2tuple(img_size), to_2tuple(patch_size), to_2tuple(stride)
        if in_chans != 1 or patch_size[0] != patch_size[1] or stride[0] != stride[1]:
            raise NotImplementedError("PatchEmbed: single-channel square patches only (the PaSST configuration)")
        self.img_size, self.patch_size, self.stride = img_size, patch_size, stride
        self.grid_size = (img_size[0] // stride[0], img_size[1] // stride[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        if attn_drop or proj_drop:
            raise NotImplementedError("dropout is 0 in every shipped recipe and is not implemented")
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, att_mask=None, residual=None):
        if att_mask is not None:
            raise NotImplementedError("att_mask is unused by the shipped recipes")
        qkv = F.linear(x, self.qkv.weight, self.qkv.bias)
        o = F.attention(qkv, self.num_heads)
        return F.linear(o, self.proj.weight, self.proj.bias, residual=residual)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm):
        super().__init__()
        if drop_path:
            raise NotImplementedError("drop_path is 0 in every shipped recipe and is not implemented")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x, att_mask=None):
        # x = x + attn(norm1(x)); x = x + mlp(norm2(x))   (passt.py:360-363); residual adds ride in the GEMM epilogues
        y, res = F.layer_norm_res(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        x = self.attn(y, att_mask, residual=res)
        y, res = F.layer_norm_res(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return self.mlp(y, residual=res)


class PaSST(nn.Module):
    def __init__(self, u_patchout=0, s_patchout_t=0, s_patchout_f=0, img_size=(128, 998), patch_size=16, stride=16, in_chans=1,
                 num_classes=527, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4., qkv_bias=True, representation_size=None,
                 distilled=False, drop_rate=0., attn_drop_rate=0., drop_path_rate=0., embed_layer=PatchEmbed, norm_layer=None,
                 act_layer=None, weight_init=''):
        super().__init__()
        if not distilled:
            raise NotImplementedError("PaSST: only the distilled (cls + dist token) variant the SED models build")
        if u_patchout or s_patchout_t or s_patchout_f:
            raise NotImplementedError("patchout is 0 in every SED recipe (passt_sed.py:77) and is not implemented")
        self.num_classes = num_classes
        self.u_patchout, self.s_patchout_t, self.s_patchout_f = u_patchout, s_patchout_t, s_patchout_f
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 2
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        act_layer = act_layer or nn.GELU
        self.patch_embed = embed_layer(img_size=img_size, patch_size=patch_size, stride=stride, in_chans=in_chans,
                                       embed_dim=embed_dim, flatten=False)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.new_pos_embed = nn.Parameter(torch.zeros(1, self.num_tokens, embed_dim))
        self.freq_new_pos_embed = nn.Parameter(torch.zeros(1, embed_dim, self.patch_embed.grid_size[0], 1))
        self.time_new_pos_embed = nn.Parameter(torch.zeros(1, embed_dim, 1, self.patch_embed.grid_size[1]))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.blocks = nn.Sequential(*[
            self.make_block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                            attn_drop=attn_drop_rate, drop_path=0., norm_layer=norm_layer, act_layer=act_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.pre_logits = nn.Identity()
        # classification heads are unused by the SED path but kept so that PaSST checkpoints load with strict=True
        self.head = nn.Sequential(nn.LayerNorm(self.num_features),
                                  self.make_head_linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity())
        self.head_dist = self.make_head_linear(self.embed_dim, self.num_classes) if num_classes > 0 else nn.Identity()
        self.init_weights(weight_init)

    # construction hooks (passt_lora.PaSST swaps in LoRA layers)
    def make_block(self, **kw):
        return Block(**kw)

    def make_head_linear(self, in_features, out_features):
        return nn.Linear(in_features, out_features)

    def init_weights(self, mode=''):
        for p in (self.new_pos_embed, self.freq_new_pos_embed, self.time_new_pos_embed, self.dist_token, self.cls_token):
            nn.init.trunc_normal_(p, std=.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.zeros_(m.bias)
                nn.init.ones_(m.weight)

    def no_weight_decay(self):
        return {'new_pos_embed', 'freq_new_pos_embed', 'time_new_pos_embed', 'cls_token', 'dist_token'}

    # ---- token-major fast path used by PaSST_SED ---------------------------------------------------------------
    def forward_tokens(self, mel, feature_layers=(), windows=None):
        """mel [B, n_mels, T] -> (dict {layer k: tokens [B, N, D] after block k}, final-norm tokens [B, N, D], f_dim, t_dim).
        Token-major tensors, no transposes, only the requested layers are kept (the reference materialises 12 fp32
        transposed copies, passt.py:574-576).

        `windows` = (starts, width): instead of the whole image, every crop mel[:, :, s:s+width] (s in starts) is embedded as its
        own sequence and all of them ride through the blocks as ONE batch, window-major [(w, b), N_w, D] -- the batched form
        of the reference's per-window backbone calls (encoder_slide_window.py:29-33)."""
        pe = self.patch_embed
        H, W = mel.shape[-2], mel.shape[-1]
        if windows is not None:
            W = int(windows[1])
        if not (H == pe.img_size[0] and W == pe.img_size[1]) and W not in getattr(self, "_warned", ()):
            warnings.warn(f"Input image size ({H}*{W}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]}).")
            self._warned = getattr(self, "_warned", ()) + (W,)
        P, S = pe.patch_size[0], pe.stride[0]
        f_dim = (H - P) // S + 1
        t_full = (W - P) // S + 1
        t_table = self.time_new_pos_embed.shape[-1]
        t_dim = min(t_full, t_table)

        def draw():
            if t_full < t_table and self.training:
                return torch.randint(1 + t_table - t_full, (1,)).item()  # same draw as the reference (passt.py:508)
            return 0

        spec, toffset = None, 0
        if windows is None:
            toffset = draw()
        else:
            starts = [int(v) for v in windows[0]]
            spec = (starts, t_dim, [draw() for _ in starts])    # one draw per window, in window order, like the reference loop
        x = F.patch_embed(mel, pe.proj.weight, pe.proj.bias, self.time_new_pos_embed.reshape(self.embed_dim, t_table),
                          self.freq_new_pos_embed.reshape(self.embed_dim, f_dim), self.cls_token.reshape(-1),
                          self.dist_token.reshape(-1), self.new_pos_embed.reshape(2, -1), stride=S, t_offset=toffset, windows=spec)
        feats = {}
        for k, block in enumerate(self.blocks):
            x = block(x)
            if (k + 1) in feature_layers:
                feats[k + 1] = x
        frame = F.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return feats, frame, f_dim, t_dim

    # ---- reference-compatible dict output -----------------------------------------------------------------------
    def forward(self, x):
        """x [B, 1, n_mels, T] -> dict with 'layer{k}_out' [B, C, P] fp32, 'frame', 'f_dim', 't_dim' (passt.py:492-591)."""
        feats, frame, f_dim, t_dim = self.forward_tokens(x[:, 0], feature_layers=range(1, len(self.blocks) + 1))
        out = {'origin_f_dim': f_dim, 'origin_t_dim': t_dim, 'f_dim': f_dim, 't_dim': t_dim}
        for k, v in feats.items():
            out['layer{}_out'.format(k)] = F.cast(v, torch.float32).transpose(1, 2)
        out['frame'] = F.cast(frame, torch.float32).transpose(1, 2)
        return out
