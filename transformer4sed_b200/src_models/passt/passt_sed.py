"""Drop-in for reference ``src/models/passt/passt_sed.py``: `PaSST_SED` (MAT-SED) and `InterpolateModule`.

Same constructor kwargs (straight from the recipes' YAML), same ``forward`` signature and return tuples, same parameter
names, same hookable submodules (`interpolate_module`, `decoder.encoder_blocks[i]`, `backbone.blocks`), but every
arithmetic step runs in libt4s kernels.  Only the decoder the MAT-SED / PMAM / DASM configs select is provided:
``decoder='transformerXL'`` (or ``'no'``); GRU / vanilla / conformer decoders are not on the hot path.
"""
import torch
import torch.nn as nn

from ... import functional as F
from ..pooling import AttentionPooling
from ..sed_model import SEDModel
from ..transformer.mask import MlmModule
from ..transformer_decoder import TransformerXLDecoder
from .passt_feature_extraction import PasstFeatureExtractor


class InterpolateModule(nn.Module):
    """Kept as a module so forward hooks can capture the frame sequence (reference :13-34)."""

    def __init__(self, mode="linear") -> None:
        super().__init__()
        if mode != "linear":
            raise NotImplementedError("only linear interpolation is used by the shipped configs")
        self.mode = mode

    def forward(self, seq, ratio, pad_last=False):
        """seq [B, T, C] -> [B, T*ratio, C]; `pad_last` repeats the final frame first (fused 99 -> 100 -> 1000)."""
        if ratio == 1 and not pad_last:
            return seq
        return F.pad_interpolate(seq, ratio, pad=pad_last)


class PaSST_SED(SEDModel):

    def __init__(self, decode_ratio=10, interpolate_mode='linear', passt_feature_layer=10, embed_dim=768, decoder_dim=768,
                 f_pool='mean_pool', s_patchout_f=0, s_patchout_t=0, decoder='gru', decoder_layer_num=2, decoder_pos_emd_len=1000,
                 load_pretrained_model=True, class_num=10, at_adapter=False, decoder_win_len=None, mlm=False, mlm_dict=dict(),
                 lora_config=None, strict_upstream=True):
        super().__init__()
        self.mel_trans = PasstFeatureExtractor(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, htk=False, fmin=0.0,
                                               fmax=None, wav_norm=True, fmin_aug_range=10, fmax_aug_range=2000)
        assert s_patchout_t == 0, "SED task do not support temporal patchout"
        passt_params_dict = dict(u_patchout=0, s_patchout_t=s_patchout_t, s_patchout_f=s_patchout_f, img_size=(128, 998), patch_size=16,
                                 stride=10, in_chans=1, num_classes=527, embed_dim=embed_dim, depth=12, num_heads=12, mlp_ratio=4,
                                 qkv_bias=True, representation_size=None, distilled=True, drop_rate=0, attn_drop_rate=0,
                                 drop_path_rate=0., norm_layer=None, act_layer=None, weight_init='')
        if lora_config is not None:
            from .passt_lora import PaSST
            passt_params_dict["lora_config"] = lora_config
        else:
            from .passt import PaSST
        self.backbone = PaSST(**passt_params_dict)
        if load_pretrained_model:
            sd = torch.load('./pretrained_model/passt-s-f128-p16-s10-ap.476-swa.pt', map_location="cpu")
            self.backbone.load_state_dict(sd, strict=False)

        self.f_pool_name = f_pool
        self.passt_feature_layer = passt_feature_layer
        self.decoder_name = decoder
        self.decode_ratio = decode_ratio
        self.class_num = class_num
        self.embed_dim = embed_dim
        self.decoder_dim = decoder_dim
        self.strict_upstream = strict_upstream

        self.out_norm = nn.LayerNorm(embed_dim)
        self.init_f_pool(pool_name=f_pool, embed_dim=embed_dim)
        self.interpolate_module = InterpolateModule(mode=interpolate_mode)
        self.slide_window_layer = nn.Identity()
        self.mlm = mlm
        if mlm:
            self.init_mlm(device=None, mlm_dict=mlm_dict)
        self.init_decoder(decoder_win_len, decoder_layer_num, decoder_pos_emd_len)
        self.at_adpater = at_adapter  # (upstream spelling; state-dict keys depend on it)
        if self.at_adpater:
            self.at_adpater = nn.Sequential(AttentionPooling(embed_dim=embed_dim, num_head=12), nn.Linear(embed_dim, class_num))

    def init_f_pool(self, pool_name, embed_dim):
        if pool_name == 'mean_pool':
            pass
        elif pool_name == "attention":
            self.f_pool_module = AttentionPooling(embed_dim=embed_dim, num_head=6)
        else:
            raise NotImplementedError("pool method {0} hasn't been implemneted yet".format(pool_name))

    def init_decoder(self, win_len, decoder_layer_num, decoder_pos_emd_len):
        self.decoder_layer_num = decoder_layer_num
        if self.decoder_name == "transformerXL":
            self.decoder = TransformerXLDecoder(input_dim=self.decoder_dim, seq_len=decoder_pos_emd_len, window_len=win_len,
                                                decoder_layer_num=decoder_layer_num)
        elif self.decoder_name == 'no':
            self.decoder = torch.nn.Identity()
        else:
            raise NotImplementedError(f"decoder '{self.decoder_name}' is not on the B200 hot path; use 'transformerXL' (all shipped "
                                      "MAT-SED / PMAM / DASM configs) or 'no'")
        self.classifier = nn.Linear(self.decoder_dim, self.class_num)

    def init_mlm(self, device, mlm_dict=dict()):
        out_dim = mlm_dict["out_dim"]
        self.mlm_tool = MlmModule(device=device, **mlm_dict)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.decoder_dim))
        torch.nn.init.normal_(self.mask_token, std=.02)
        self.mlm_mlp = nn.Sequential(torch.nn.Linear(self.decoder_dim, self.decoder_dim), torch.nn.GELU(),
                                     torch.nn.Linear(self.decoder_dim, out_dim))

    # ---- pieces of forward ------------------------------------------------------------------------------------------
    def f_pool(self, feat_tokens, f_dim, t_dim):
        """layer-k tokens [B, 2 + F*T', C] -> frame sequence [B, T', C]: drop cls/dist, out_norm, pool over frequency."""
        y = F.layer_norm(feat_tokens, self.out_norm.weight, self.out_norm.bias, self.out_norm.eps, skip=2)
        if self.f_pool_name == 'mean_pool':
            return F.fpool_mean(y, f_dim, t_dim)
        B, _, C = y.shape
        y = y.reshape(B, f_dim, t_dim, C).transpose(1, 2).reshape(B * t_dim, f_dim, C)
        return self.f_pool_module(y).reshape(B, t_dim, C)

    def decoder_step(self, x, other_dict, encoder_win=False):
        other_dict["frame_before_mask"] = x
        if self.mlm:
            # Upstream writes the mask through `token_seq.clone().reshape(-1, C)` (mask.py:65-79): a view only when the frame sequence is
            # contiguous.  From the transposed interpolation output with B > 1 it is not (the write is lost: SURVEY §9.1); with the
            # sliding-window fusion `mix_rate * x_local + ...` takes x_local's contiguous layout and the mask IS applied.
            noop = self.strict_upstream and x.shape[0] > 1 and type(self) is PaSST_SED and not encoder_win
            x, mask_id_seq = self.mlm_tool.setence_mask(x, self.mask_token, apply=not noop)
            other_dict["mask_id_seq"] = mask_id_seq
        return self.decoder(x)

    def at_forward(self, at_embedding, other_dict, skip=0):
        pooled = self.at_adpater[0](at_embedding, skip=skip)
        logit = F.linear(pooled, self.at_adpater[1].weight, self.at_adpater[1].bias, out_dtype=torch.float32)
        other_dict['at_out'] = F.sigmoid(logit)
        return other_dict

    def forward(self, input: torch.Tensor, encoder_win=False, mix_rate=0.5, win_param=[512, 49], temp_w=1, pad_mask=None):
        """input: log-mel [B, 128, T].  Returns (strong [B, C, T'], weak [B, C], other_dict), or (pred, other_dict) when mlm."""
        other_dict = {}
        feats, frame, f_dim, t_dim = self.backbone.forward_tokens(input, feature_layers=(self.passt_feature_layer,))
        x = self.f_pool(feats[self.passt_feature_layer], f_dim, t_dim)
        x = self.interpolate_module(x, self.decode_ratio, pad_last=True)  # 99 -> (pad) 100 -> x10 = 1000 frames
        assert x.shape[1] == 1000
        if encoder_win:
            from .passt_win import PasstWithSlide
            slide_window_model = PasstWithSlide(net=self, win_param=win_param)
            x_local = self.slide_window_layer(slide_window_model(input, emb_len=x.shape[1]))
            x = F.lerp(x, x_local, mix_rate)
        x = self.decoder_step(x, other_dict, encoder_win=bool(encoder_win))
        if self.at_adpater:
            other_dict = self.at_forward(frame, other_dict, skip=2)  # patch tokens only; no [B,1188,C] slice copy
        if self.mlm:
            h = F.linear(x, self.mlm_mlp[0].weight, self.mlm_mlp[0].bias, act=F.ops.ACT_GELU)
            return F.linear(h, self.mlm_mlp[2].weight, self.mlm_mlp[2].bias), other_dict
        logits = F.linear(x, self.classifier.weight, self.classifier.bias, out_dtype=torch.float32)
        sed_out, at_out = F.sed_pool(logits, temp_w, pad_mask)
        return sed_out, at_out, other_dict

    def get_feature_extractor(self):
        return self.mel_trans

    def get_model_name(self):
        return "PaSST_SED"

    def get_backbone_upsample_ratio(self):
        return self.decode_ratio

    def get_backbone(self):
        return self.backbone
