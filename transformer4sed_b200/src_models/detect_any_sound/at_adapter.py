"""Drop-in for reference ``src/models/detect_any_sound/at_adapter.py``: the query-based audio-tagging decoder of DASM.

`CrossAttentionFirstDecoderLayer` stays an `nn.TransformerDecoderLayer` subclass, so parameter names / shapes
(``self_attn.in_proj_weight``, ``multihead_attn.*``, ``linear1``, ``linear2``, ``norm1..3``) are torch's own and reference
checkpoints load unchanged; only `forward` is replaced: cross-attention over the encoder memory FIRST, then (masked)
self-attention among the queries, then the GELU feed-forward, post-norm (reference :16-32) -- on libt4s kernels.
"""
from typing import Optional

import torch.nn as nn
from torch import Tensor

from ... import functional as F
from ... import ops


class CrossAttentionFirstDecoderLayer(nn.TransformerDecoderLayer):

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.norm_first:
            raise NotImplementedError("norm_first decoder layers are not used by DASM")

    def _attend(self, x, memory, att, mask, p):
        """norm(x + dropout(MHA(x, memory)));  without dropout the residual rides in the out_proj GEMM epilogue."""
        fused = not (self.training and p > 0)
        y = F.multi_head_attention(x, memory, att.in_proj_weight, att.in_proj_bias, att.out_proj.weight, att.out_proj.bias, att.num_heads,
                                   attn_mask=mask, dropout_p=att.dropout, training=self.training, residual=x if fused else None)
        return y if fused else F.add(x, F.dropout(y, p, True))

    def forward(self, tgt: Tensor, memory: Tensor, tgt_mask: Optional[Tensor] = None, memory_mask: Optional[Tensor] = None,
                tgt_key_padding_mask: Optional[Tensor] = None, memory_key_padding_mask: Optional[Tensor] = None, tgt_is_causal: bool = False,
                memory_is_causal: bool = False) -> Tensor:
        if memory_mask is not None or tgt_key_padding_mask is not None or memory_key_padding_mask is not None:
            raise NotImplementedError("only the boolean tgt_mask of DASM.forward is supported")
        x = tgt
        x = self._attend(x, memory, self.multihead_attn, None, self.dropout2.p)
        x = F.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        x = self._attend(x, x, self.self_attn, tgt_mask, self.dropout1.p)
        x = F.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        p_in, p_out = self.dropout.p, self.dropout3.p
        h = F.linear(x, self.linear1.weight, self.linear1.bias, act=ops.ACT_GELU)
        if self.training and (p_in > 0 or p_out > 0):
            y = F.linear(F.dropout(h, p_in, True), self.linear2.weight, self.linear2.bias)
            x = F.add(x, F.dropout(y, p_out, True))
        else:
            x = F.linear(h, self.linear2.weight, self.linear2.bias, residual=x)
        return F.layer_norm(x, self.norm3.weight, self.norm3.bias, self.norm3.eps)


class QueryBasedAudioTaggingDecoder(nn.Module):

    def __init__(self, n_layers, d_model, nhead, dim_ffn, activation="gelu"):
        super().__init__()
        if activation != "gelu":
            raise NotImplementedError("libt4s fuses exact-erf GELU only")
        decoder_layer = CrossAttentionFirstDecoderLayer(d_model=d_model, nhead=nhead, dim_feedforward=dim_ffn, activation=activation,
                                                        batch_first=True)
        self.decoder = nn.TransformerDecoder(decoder_layer, num_layers=n_layers)

    def forward(self, feat_encoder: Tensor, queries: Tensor, tgt_mask=None):
        out = F.to_act(queries)
        memory = F.to_act(feat_encoder)
        for layer in self.decoder.layers:
            out = layer(out, memory, tgt_mask=tgt_mask)
        return out
