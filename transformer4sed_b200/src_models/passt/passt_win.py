"""Drop-in for reference ``src/models/passt/passt_win.py`` (`PasstWithSlide`): PaSST backbone under the sliding window.

`encode` keeps the reference's per-crop contract (mel crop [B, 128, w] -> frames [B, t*ratio, C]); `encode_windows` is what
`EncoderSlideWindow.__call__` actually uses: all crops of one group go through patch-embed (windowed im2col), the 12 blocks,
out_norm + frequency pooling and the x`ratio` interpolation as one batch.
"""
import torch

from ... import functional as F
from ..encoder_slide_window import EncoderSlideWindow


class PasstWithSlide(EncoderSlideWindow):

    def __init__(self, net, win_param=[512, 29]):
        super().__init__(net, win_param, out_dim=net.embed_dim)

    def frames_per_window(self, width):
        pe = self.net.get_backbone().patch_embed
        return min((width - pe.patch_size[0]) // pe.stride[0] + 1, self.net.get_backbone().time_new_pos_embed.shape[-1])

    def _frames(self, feats, f_dim, t_dim):
        x = self.net.f_pool(feats[self.net.passt_feature_layer], f_dim, t_dim)
        ratio = self.net.get_backbone_upsample_ratio()
        if ratio != 1:
            # F.interpolate directly, not net.interpolate_module: hooks on that module must not fire per window (passt_win.py:34-40)
            x = F.pad_interpolate(x, ratio, pad=False)
        return x

    def encode_windows(self, input, starts, width):
        feats, _, f_dim, t_dim = self.net.get_backbone().forward_tokens(input, feature_layers=(self.net.passt_feature_layer,),
                                                                        windows=(starts, width))
        return self._frames(feats, f_dim, t_dim)

    def encode(self, input: torch.Tensor) -> torch.Tensor:
        feats, _, f_dim, t_dim = self.net.get_backbone().forward_tokens(input.contiguous(), feature_layers=(self.net.passt_feature_layer,))
        return self._frames(feats, f_dim, t_dim)
