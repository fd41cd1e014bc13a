"""Drop-in for reference ``src/models/detect_any_sound/detect_any_sound.py``: `DASM` (open-vocabulary, query-based sound event
detection) and its `MLP`.  Same constructor dictionaries, forward signature and return tuples, same parameter names.

Supported (everything the released inference / survey configurations use): PaSST backbone with or without LoRA, the `base` CNN
branch or none, attention frequency pooling, `decoder` in {'transformerXL', 'no'}, learnable or externally supplied queries
(single `query_projector`), `out_type` in {'sigmoid', 'logit', None}, optional MLM.  GRU / vanilla / conformer decoders and the
multi-modal query list are not on the B200 path and raise `NotImplementedError`.
"""
from typing import List, Literal, Optional, Union

import torch
import torch.nn as nn

from ... import functional as F
from ... import ops
from ..cnn import CNN
from ..passt.passt_feature_extraction import PasstFeatureExtractor
from ..passt.passt_sed import InterpolateModule
from ..pooling import AttentionPooling
from ..sed_model import SEDModel
from ..transformer.mask import MlmModule
from ..transformer_decoder import TransformerXLDecoder
from .at_adapter import QueryBasedAudioTaggingDecoder


class MLP(nn.Module):
    """Linear (+ GELU) x num_layers (reference :404-416)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x, out_dtype=None):
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            x = F.linear(x, layer.weight, layer.bias, act=ops.ACT_NONE if last else ops.ACT_GELU, out_dtype=out_dtype if last else None)
        return x


class DASM(SEDModel):

    def __init__(self, cnn_param, backbone_param={"embed_dim": 768, "passt_feature_layer": 10, "pretrain_model_path": None, "lora_config": None},
                 at_param={"at_decoder_layer": 0, "query_projector": False, "query_dim": 768, "out_type": "logit", "query": None},
                 mlm_dict=None, backbone_upsample_ratio=10, decoder_dim=768, num_heads=12, decoder='gru', decoder_layer_num=2,
                 decoder_pos_emd_len=1000, decoder_expand_rate=1, class_num=10) -> None:
        super().__init__()
        if cnn_param is not None:
            self._init_cnn(dict(cnn_param))
        self.backbone_param = backbone_param
        self._init_transformer_backbone(backbone_param)
        self.decoder_name = decoder
        self.backbone_upsample_ratio = backbone_upsample_ratio
        self.class_num = class_num
        self.decoder_dim = decoder_dim
        self.num_heads = num_heads
        self.decoder_expand_rate = decoder_expand_rate
        self.mlm_dict = mlm_dict
        if mlm_dict is not None:
            self._init_mlm(device=None, mlm_dict=mlm_dict)
        self._init_SED_decoder(decoder_layer_num, decoder_pos_emd_len, at_param)
        self._init_AT_head(at_param)
        self._init_joint_layers(transformer_embed_dim=backbone_param["embed_dim"])

    def _init_joint_layers(self, transformer_embed_dim):
        self.interpolate_module = InterpolateModule(mode='linear')
        self.f_pool_module = AttentionPooling(embed_dim=transformer_embed_dim, num_head=6)
        if hasattr(self, "cnn"):
            self.cnn_projector = torch.nn.Linear(self.cnn_feat_dim, self.decoder_dim)
            self.merge_weight = torch.nn.Parameter(torch.Tensor([0.5]), requires_grad=self.mlm_dict is not None)
        self.transformer_projector = torch.nn.Linear(transformer_embed_dim, self.decoder_dim)
        self.at_projector = torch.nn.Linear(transformer_embed_dim, self.decoder_dim)
        self.norm_before_pool = nn.LayerNorm(transformer_embed_dim)
        self.norm_after_merge = nn.LayerNorm(self.decoder_dim)

    def _init_SED_decoder(self, decoder_layer_num, decoder_pos_emd_len, at_param):
        self.decoder_layer_num = decoder_layer_num
        if self.decoder_name == "transformerXL":
            self.sed_decoder = TransformerXLDecoder(input_dim=self.decoder_dim, seq_len=decoder_pos_emd_len, decoder_layer_num=decoder_layer_num,
                                                    mlp_ratio=self.decoder_expand_rate, num_heads=self.num_heads)
        elif self.decoder_name == 'no':
            self.sed_decoder = torch.nn.Identity()
        else:
            raise NotImplementedError(f"decoder '{self.decoder_name}' is not on the B200 hot path; use 'transformerXL' or 'no'")
        self.mask_embedding_layer = MLP(self.decoder_dim, self.decoder_dim, self.decoder_dim, 3) if at_param["out_type"] else nn.Identity()
        self.sed_head = nn.Linear(self.decoder_dim, self.decoder_dim)

    def _init_query(self, query_projector: bool = False, query: Union[str, List[str], torch.Tensor] = None, query_dim: Union[int, List[int]] = None):
        if not query_projector:
            self.at_query = nn.Parameter(torch.zeros(self.class_num, self.decoder_dim))
            torch.nn.init.normal_(self.at_query, std=.02)
            return
        if not isinstance(query_dim, int):
            raise NotImplementedError("multi-modal query lists are not on the B200 path (single query_projector only)")
        self.query_projector = nn.Sequential(nn.Linear(query_dim, self.decoder_dim), nn.GELU())
        if query is not None:
            if isinstance(query, str):
                query = torch.load(query, map_location="cpu")
            assert isinstance(query, torch.Tensor), "query are expected to be torch.Tensor, but not {0}".format(type(query))
            assert query.shape[0] == self.class_num
            self.at_query = nn.Parameter(query)

    def _init_AT_head(self, at_param: dict):
        self._init_query(at_param["query_projector"], at_param["query"], at_param["query_dim"])
        self.at_decoder = QueryBasedAudioTaggingDecoder(n_layers=at_param["at_decoder_layer"], nhead=self.num_heads, d_model=self.decoder_dim,
                                                        dim_ffn=self.decoder_dim * self.decoder_expand_rate)
        if at_param["out_type"] == "logit":
            self.at_head = MLP(self.decoder_dim, self.decoder_dim, self.class_num + 1, 2)
        elif at_param["out_type"] == "sigmoid":
            self.at_head = MLP(self.decoder_dim, self.decoder_dim, 1, 2)
        elif at_param["out_type"] is None:
            self.at_head = None
        else:
            raise RuntimeError("Unknown output type for classification branch")

    def _init_mlm(self, device, mlm_dict=dict()):
        out_dim = mlm_dict["out_dim"]
        self.mlm_tool = MlmModule(device=device, **mlm_dict)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.decoder_dim))
        torch.nn.init.normal_(self.mask_token, std=.02)
        self.mlm_mlp = nn.Sequential(torch.nn.Linear(self.decoder_dim, self.decoder_dim), torch.nn.GELU(), torch.nn.Linear(self.decoder_dim, out_dim))

    def _init_transformer_backbone(self, backbone_param):
        self.mel_trans = PasstFeatureExtractor(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, htk=False, fmin=0.0, fmax=None,
                                               wav_norm=True, fmin_aug_range=10, fmax_aug_range=2000)
        kw = dict(img_size=(128, 998), patch_size=16, stride=10, in_chans=1, embed_dim=backbone_param["embed_dim"], depth=12, num_heads=12,
                  mlp_ratio=4, qkv_bias=True, distilled=True)
        if backbone_param["lora_config"] is not None:
            kw["lora_config"] = backbone_param["lora_config"]
            from ..passt.passt_lora import PaSST
        else:
            from ..passt.passt import PaSST
        self.backbone = PaSST(**kw)
        if backbone_param["pretrain_model_path"] is not None:
            sd = torch.load(backbone_param['pretrain_model_path'], map_location="cpu")
            self.backbone.load_state_dict(sd, strict=False)

    def _init_cnn(self, cnn_param):
        self.cnn_feat_dim = cnn_param["nb_filters"][-1] if "cnn_1d_dict" not in cnn_param else cnn_param["cnn_1d_dict"]["filters"][-1]
        self.cnn = CNN(**cnn_param)

    # ---- pieces of forward ------------------------------------------------------------------------------------------
    def f_pool(self, feat_tokens, f_dim, t_dim):
        """layer-k tokens [B, 2 + F*T', C] -> [B, T', C]: norm_before_pool on the patch tokens, attention pooling over frequency (:242-254)."""
        y = F.layer_norm(feat_tokens, self.norm_before_pool.weight, self.norm_before_pool.bias, self.norm_before_pool.eps, skip=2)
        B, _, C = y.shape
        y = y.reshape(B, f_dim, t_dim, C).transpose(1, 2).reshape(B * t_dim, f_dim, C)
        return self.f_pool_module(y).reshape(B, t_dim, C)

    def sed_branch(self, x):
        return self.sed_decoder(x)

    def at_branch(self, mask_feat, query=None, query_type=None, tgt_mask=None):
        query = self.at_query if query is None else query
        if isinstance(query, (list, nn.ParameterList)):
            raise NotImplementedError("multi-modal query lists are not on the B200 path")
        if hasattr(self, "query_projector"):
            lin = self.query_projector[0]
            query = F.linear(F.to_act(query.to(mask_feat.device)), lin.weight, lin.bias, act=ops.ACT_GELU)
        else:
            query = F.to_act(query)
        mask_feat = self.at_decoder(feat_encoder=mask_feat, queries=query.expand(mask_feat.shape[0], -1, -1), tgt_mask=tgt_mask)
        at_out = None
        if getattr(self, 'at_head', None) is not None:
            at_out = self.at_head(mask_feat, out_dtype=torch.float32)
            if at_out.shape[-1] == 1:
                at_out = F.sigmoid(at_out.squeeze(-1))
        return at_out, mask_feat

    def forward(self, input, encoder_win=False, mix_rate=0.5, win_param=[512, 49], temp_w=0.1, pad_mask=None,
                query: Union[torch.Tensor, list] = None, query_type: Optional[Literal['text', 'audio']] = None, tgt_mask=None):
        other_dict = {}
        layer = self.backbone_param["passt_feature_layer"]
        feats, frame, f_dim, t_dim = self.backbone.forward_tokens(input, feature_layers=(layer,))
        x = self.f_pool(feats[layer], f_dim, t_dim)
        x = self.interpolate_module(x, self.backbone_upsample_ratio, pad_last=True)
        if encoder_win:
            raise NotImplementedError("DASM.forward(encoder_win=True) imports a non-existent class upstream (detect_any_sound.py:328)")
        x = F.linear(x, self.transformer_projector.weight, self.transformer_projector.bias)
        if hasattr(self, "cnn"):
            cnn_feat = self.cnn.forward_cl(input, mel_layout=True)
            Bc, cnn_t, cnn_f, cnn_channel = cnn_feat.shape
            assert cnn_channel == self.cnn_feat_dim
            assert cnn_f == 1
            cnn_feat = F.pad_interpolate(cnn_feat.reshape(Bc, cnn_t, cnn_channel), x.shape[1] // cnn_t, pad=False)
            x = F.scale_add(x, F.linear(cnn_feat, self.cnn_projector.weight, self.cnn_projector.bias), self.merge_weight)
        x = F.layer_norm(x, self.norm_after_merge.weight, self.norm_after_merge.bias, self.norm_after_merge.eps)

        # AT decoder over the final-norm patch tokens (cls / dist tokens dropped)
        at_feat = F.linear(frame[:, 2:, :], self.at_projector.weight, self.at_projector.bias)
        if isinstance(query, torch.Tensor) and query.ndim == 3:
            query = query[0, :, :]
        if tgt_mask is not None and tgt_mask.ndim == 3:
            tgt_mask = tgt_mask[0, :, :]
        other_dict["at_out"], mask_feat = self.at_branch(at_feat, query, query_type, tgt_mask)
        if self.mlm_dict is not None:
            other_dict["frame_before_mask"] = x
            x, mask_id_seq = self.mlm_tool.setence_mask(x, self.mask_token)
            other_dict["mask_id_seq"] = mask_id_seq
        x = self.sed_branch(x)
        if self.mlm_dict is not None:
            h = F.linear(x, self.mlm_mlp[0].weight, self.mlm_mlp[0].bias, act=ops.ACT_GELU)
            return F.linear(h, self.mlm_mlp[2].weight, self.mlm_mlp[2].bias), other_dict

        x = F.linear(x, self.sed_head.weight, self.sed_head.bias)
        mask_embedding = self.mask_embedding_layer(mask_feat)
        score = F.query_frame_score(x, mask_embedding)                    # [B, T, K] = einsum('bqc,bct->bqt') transposed
        sed_out, weak_out = F.query_pool(score, other_dict["at_out"], temp_w, pad_mask)
        return sed_out, weak_out, other_dict

    def get_model_name(self):
        return "DASM"

    def get_feature_extractor(self):
        return self.mel_trans

    def get_backbone(self):
        return self.backbone

    def get_backbone_upsample_ratio(self):
        return self.backbone_upsample_ratio
