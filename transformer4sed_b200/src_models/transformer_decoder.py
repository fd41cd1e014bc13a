"""Drop-in for reference ``src/models/transformer_decoder.py``: `TransformerXLDecoder` (reference :74-122).
(`TransformerDecoder` is broken upstream and `ConformerDecoder` is not selected by any shipped config: SURVEY §2 row 6.)"""
import torch.nn as nn

from .transformer.transformerXL import RelPositionalEncoding, TransformerXL


class TransformerXLDecoder(nn.Module):

    def __init__(self, input_dim, seq_len=1000, window_len=None, decoder_layer_num=2, attn_drop=0, num_heads=12, mlp_ratio=1) -> None:
        super().__init__()
        if window_len is not None:
            raise NotImplementedError("decoder_win_len band masks are unused by the shipped configs")
        self.pos_embedding = RelPositionalEncoding(d_model=input_dim, dropout_rate=0, max_len=seq_len)
        self.encoder_blocks = nn.ModuleList([
            TransformerXL(input_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, norm_layer=nn.LayerNorm, attn_drop=attn_drop)
            for _ in range(decoder_layer_num)
        ])
        self.register_buffer("att_mask", None)

    def forward(self, x):
        """x [B, T, C] -> [B, T, C]"""
        x, pos_emb = self.pos_embedding(x)
        scale = self.pos_embedding.xscale
        x = x.transpose(0, 1)          # (T, B, C) view, as upstream permutes (reference :112); the blocks undo it without a copy
        for i, block in enumerate(self.encoder_blocks):
            x = block(x, pos_emb, in_scale=scale if i == 0 else 1.0)  # x*sqrt(d) folded into the first LayerNorm
        return x.transpose(0, 1)
