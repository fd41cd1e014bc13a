"""Drop-in for reference ``src/models/cnn_transformer/passt_cnn.py``: `PaSST_CNN` (the PMAM model: PaSST + LoRA backbone, CNN
branch, learnable merge weight, TransformerXL context net, MLM head).  Same constructor (`passt_sed_param`, `cnn_param`), same
forward signature and return tuples, same parameter names."""
import torch
import torch.nn as nn

from ... import functional as F
from ..cnn import CNN
from ..passt.passt_sed import PaSST_SED


class PaSST_CNN(PaSST_SED):

    def __init__(self, passt_sed_param, cnn_param) -> None:
        super().__init__(**passt_sed_param)
        if cnn_param is not None:
            cnn_param = dict(cnn_param)
            self.init_cnn(cnn_param)
            self.cnn_feat_dim = cnn_param["nb_filters"][-1] if "cnn_1d_dict" not in cnn_param else cnn_param["cnn_1d_dict"]["filters"][-1]
            self.cnn_projector = torch.nn.Linear(self.cnn_feat_dim, self.decoder_dim)
            self.merge_weight = torch.nn.Parameter(torch.Tensor([0.5]), requires_grad=self.mlm)
        self.transformer_projector = torch.nn.Linear(self.embed_dim, self.decoder_dim)

    def init_cnn(self, cnn_param: dict):
        self.cnn_name = cnn_param.pop('cnn_name', "base")
        if self.cnn_name != "base":
            raise NotImplementedError(f"cnn encoder '{self.cnn_name}' is not on the B200 path (the shipped PMAM / DASM configs use 'base')")
        self.cnn = CNN(**cnn_param)

    def merge_cnn(self, x, input):
        """x [B, T, embed_dim] transformer frames, input log-mel [B, F, T] -> decoder input [B, T, decoder_dim] (passt_cnn.py:52-62)."""
        if hasattr(self, "cnn"):
            cnn_feat = self.cnn.forward_cl(input, mel_layout=True)                # [B, T', 1, C]
            Bc, cnn_t, cnn_f, cnn_channel = cnn_feat.shape
            assert cnn_channel == self.cnn_feat_dim
            assert cnn_f == 1
            cnn_feat = cnn_feat.reshape(Bc, cnn_t, cnn_channel)
            if x.shape[1] % cnn_t:
                raise NotImplementedError("CNN time resolution must divide the frame count")
            cnn_feat = F.pad_interpolate(cnn_feat, x.shape[1] // cnn_t, pad=False)  # F.interpolate(size=T, mode='linear')
            a = F.linear(x, self.transformer_projector.weight, self.transformer_projector.bias)
            b = F.linear(cnn_feat, self.cnn_projector.weight, self.cnn_projector.bias)
            return F.scale_add(a, b, self.merge_weight)
        return F.linear(x, self.transformer_projector.weight, self.transformer_projector.bias)

    def forward(self, input, encoder_win=False, mix_rate=0.5, win_param=[512, 49], temp_w=1, pad_mask=None):
        other_dict = {}
        feats, frame, f_dim, t_dim = self.backbone.forward_tokens(input, feature_layers=(self.passt_feature_layer,))
        x = self.f_pool(feats[self.passt_feature_layer], f_dim, t_dim)
        x = self.interpolate_module(x, self.decode_ratio, pad_last=True)
        if encoder_win:
            from ..passt.passt_win import PasstWithSlide
            slide_window_model = PasstWithSlide(net=self, win_param=win_param)
            x_local = self.slide_window_layer(slide_window_model(input, emb_len=x.shape[1]))
            x = F.lerp(x, x_local, mix_rate)
        x = self.merge_cnn(x, input)
        x = self.decoder_step(x, other_dict)
        if self.at_adpater:
            other_dict = self.at_forward(frame, other_dict, skip=2)
        if self.mlm:
            h = F.linear(x, self.mlm_mlp[0].weight, self.mlm_mlp[0].bias, act=F.ops.ACT_GELU)
            return F.linear(h, self.mlm_mlp[2].weight, self.mlm_mlp[2].bias), other_dict
        logits = F.linear(x, self.classifier.weight, self.classifier.bias, out_dtype=torch.float32)
        sed_out, at_out = F.sed_pool(logits, temp_w, pad_mask)
        return sed_out, at_out, other_dict

    def get_model_name(self):
        return "PaSST_CNN"
