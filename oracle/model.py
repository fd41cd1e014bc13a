"""Oracle: MAT-SED (`PaSST_SED`) forward as plain functions over a state dict (CPU, torch fp32/fp64).

Test infrastructure only (see oracle/__init__.py).  Each function cites the reference lines it follows.
State-dict keys are the reference's.  Autograd works through everything here, so the
oracle also provides reference gradients.
"""
import math

import torch
import torch.nn.functional as F


def layer_norm(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


LORA_SCALING = 1.0 / 8.0   # lora_alpha / r of the shipped PMAM / DASM configs (config/pmam/post_pretrain.yaml:61-63)


def linear(x, sd, prefix, bias=True):
    """nn.Linear, or lora.Linear when the state dict carries lora_A / lora_B for this layer (lora/layers.py:143-153, un-merged:
    W x + (x A^T B^T) * alpha / r)."""
    y = F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"] if bias else None)
    if prefix + ".lora_A" in sd:
        y = y + (x @ sd[prefix + ".lora_A"].transpose(0, 1) @ sd[prefix + ".lora_B"].transpose(0, 1)) * LORA_SCALING
    return y


# ------------------------------------------------------------------------------------------------
# PaSST backbone  (reference src/models/passt/passt.py)
# ------------------------------------------------------------------------------------------------
def patch_embed_tokens(mel, sd, p="backbone.", t_offset=0):
    """mel [B,128,T] -> tokens [B, 2+F*Tp, D]; patch conv (:302-315) + time/freq pos (:503-519) +
    cls/dist tokens (:560-569).  Eval path: time pos-embed cropped from offset 0 (:511), or the
    patch grid cropped to the table's 99 columns (:515).  `t_offset` injects the train-mode random
    crop offset of the time table (:506-509)."""
    x = F.conv2d(mel.unsqueeze(1), sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=10)
    tpos = sd[p + "time_new_pos_embed"]
    if x.shape[-1] < tpos.shape[-1]:
        tpos = tpos[:, :, :, t_offset:t_offset + x.shape[-1]]
    else:
        x = x[:, :, :, :tpos.shape[-1]]
    x = x + tpos + sd[p + "freq_new_pos_embed"]
    B, D, Fd, Td = x.shape
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "cls_token"].expand(B, -1, -1) + sd[p + "new_pos_embed"][:, :1]
    dist = sd[p + "dist_token"].expand(B, -1, -1) + sd[p + "new_pos_embed"][:, 1:]
    return torch.cat((cls, dist, x), dim=1), Fd, Td


def vit_attention(x, sd, p, num_heads):
    """passt.py:330-344: fused qkv, scale hd^-1/2, softmax, proj."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = linear(x, sd, p + "qkv").reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
    return linear((attn @ v).transpose(1, 2).reshape(B, N, C), sd, p + "proj")


def vit_mlp(x, sd, p):
    """passt.py:270-276 (exact-erf GELU)."""
    return linear(F.gelu(linear(x, sd, p + "fc1")), sd, p + "fc2")


def vit_block(x, sd, p, num_heads, eps=1e-6):
    """passt.py:360-363: pre-norm residual block, LayerNorm eps 1e-6 (:410)."""
    x = x + vit_attention(layer_norm(x, sd, p + "norm1", eps), sd, p + "attn.", num_heads)
    x = x + vit_mlp(layer_norm(x, sd, p + "norm2", eps), sd, p + "mlp.")
    return x


def passt_backbone(mel, sd, depth=12, num_heads=12, feature_layer=10, p="backbone.", t_offset=0):
    """passt.py:492-583.  Returns (layer{feature_layer}_out [B,N,D], final-norm tokens [B,N,D], F, T')."""
    x, Fd, Td = patch_embed_tokens(mel, sd, p, t_offset)
    feat = None
    for k in range(depth):
        x = vit_block(x, sd, f"{p}blocks.{k}.", num_heads)
        if k + 1 == feature_layer:
            feat = x
    return feat, layer_norm(x, sd, p + "norm", 1e-6), Fd, Td


# ------------------------------------------------------------------------------------------------
# Frame sequence  (reference src/models/passt/passt_sed.py:199-218, 258-259, 23-34)
# ------------------------------------------------------------------------------------------------
def mha_pool(x, sd, p, num_heads):
    """pooling.py:37-51 AttentionPooling: one learned query over keys x [B',K,C] via nn.MultiheadAttention."""
    Bp, K, C = x.shape
    hd = C // num_heads
    w, b = sd[p + "frequency_att.in_proj_weight"], sd[p + "frequency_att.in_proj_bias"]
    q = F.linear(sd[p + "f_att_token"].reshape(1, C), w[:C], b[:C])  # batch independent
    k = F.linear(x, w[C:2 * C], b[C:2 * C]).reshape(Bp, K, num_heads, hd).transpose(1, 2)
    v = F.linear(x, w[2 * C:], b[2 * C:]).reshape(Bp, K, num_heads, hd).transpose(1, 2)
    q = q.reshape(1, num_heads, 1, hd)
    attn = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)  # [B',H,1,K]
    o = (attn @ v).transpose(1, 2).reshape(Bp, C)
    return F.linear(o, sd[p + "frequency_att.out_proj.weight"], sd[p + "frequency_att.out_proj.bias"])


def f_pool(feat, sd, Fd, Td, mode="mean_pool"):
    """passt_sed.py:199-218: drop cls/dist, out_norm (eps 1e-5), pool over the F patch rows."""
    x = layer_norm(feat[:, 2:], sd, "out_norm", 1e-5)
    B, _, C = x.shape
    x = x.reshape(B, Fd, Td, C)
    if mode == "mean_pool":
        return x.mean(dim=1)
    if mode == "attention":
        return mha_pool(x.transpose(1, 2).reshape(B * Td, Fd, C), sd, "f_pool_module.", 6).reshape(B, Td, C)
    raise NotImplementedError(mode)


def pad_interpolate(x, ratio=10):
    """passt_sed.py:258-259 + InterpolateModule (:23-34): repeat last frame, F.interpolate(linear,
    align_corners=False) x ratio."""
    x = torch.cat((x, x[:, -1:, :]), dim=1)
    if ratio == 1:
        return x
    return F.interpolate(x.transpose(1, 2), scale_factor=ratio, mode="linear").transpose(1, 2)


# ------------------------------------------------------------------------------------------------
# TransformerXL context network  (reference src/models/transformer/transformerXL.py,
# src/models/transformer_decoder.py:74-122)
# ------------------------------------------------------------------------------------------------
def rel_pos_table(T, d_model, dtype=torch.float32):
    """transformerXL.py:68-101,120-126: rows k=0..2T-2 hold relative position T-1-k
    (sin on even, cos on odd channels)."""
    rel = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(2 * T - 1, d_model)
    pe[:, 0::2] = torch.sin(rel * div)
    pe[:, 1::2] = torch.cos(rel * div)
    return pe.to(dtype)


def rel_shift(bd):
    """transformerXL.py:254-297: out[..., i, j] = bd[..., i, T-1-i+j]."""
    T = bd.shape[-2]
    idx = (T - 1 - torch.arange(T).unsqueeze(1)) + torch.arange(T).unsqueeze(0)
    return bd.gather(-1, idx.to(bd.device).expand(bd.shape[:-2] + (T, T)))


def relpos_attention(x, pos, sd, p, num_heads):
    """transformerXL.py:299-593 with query=key=value=x [B,T,C] (we keep batch-major; the reference's
    (T,B,C) permutes are layout only).  score = ((q+u)k^T + shift((q+v)p^T)) * hd^-1/2."""
    B, T, C = x.shape
    hd = C // num_heads
    q, k, v = linear(x, sd, p + "in_proj").chunk(3, dim=-1)
    q = q.reshape(B, T, num_heads, hd)
    k = k.reshape(B, T, num_heads, hd).permute(0, 2, 3, 1)
    v = v.reshape(B, T, num_heads, hd).transpose(1, 2)
    pp = F.linear(pos, sd[p + "linear_pos.weight"]).reshape(-1, num_heads, hd).permute(1, 2, 0)  # [H,hd,2T-1]
    qu = (q + sd[p + "pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[p + "pos_bias_v"]).transpose(1, 2)
    ac = qu @ k
    bd = rel_shift(qv @ pp)
    attn = ((ac + bd) * hd ** -0.5).softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, T, C)
    return linear(o, sd, p + "out_proj")


def txl_block(x, pos, sd, p, num_heads):
    """transformerXL.py:31-35: residual taken from the NORMALISED input; timm Mlp ratio 1."""
    x = layer_norm(x, sd, p + "norm1", 1e-5)
    x = x + relpos_attention(x, pos, sd, p + "attn.", num_heads)
    x = x + vit_mlp(layer_norm(x, sd, p + "norm2", 1e-5), sd, p + "mlp.")
    return x


def txl_decoder(x, sd, n_layers, num_heads=12, p="decoder."):
    """transformer_decoder.py:110-122 + RelPositionalEncoding.forward (transformerXL.py:104-127)."""
    B, T, C = x.shape
    pos = rel_pos_table(T, C, x.dtype).to(x.device)
    x = x * math.sqrt(C)
    for i in range(n_layers):
        x = txl_block(x, pos, sd, f"{p}encoder_blocks.{i}.", num_heads)
    return x


# ------------------------------------------------------------------------------------------------
# Heads and losses
# ------------------------------------------------------------------------------------------------
def sed_head(x, sd, temp_w=1.0, pad_mask=None):
    """passt_sed.py:285-296: classifier, sigmoid(x/temp), zero padded frames, linear-softmax pool."""
    p = torch.sigmoid(linear(x, sd, "classifier") / temp_w)
    if pad_mask is not None:
        p = p.masked_fill(pad_mask.unsqueeze(-1) if pad_mask.dim() == 2 else pad_mask, 0.0)
    weak = torch.clamp((p * p).sum(dim=1) / p.sum(dim=1), 1e-7, 1.0)
    return p.transpose(1, 2), weak


def at_branch(frame, sd):
    """passt_sed.py:236-240,276-278: AttentionPooling(12 heads) over final-norm patch tokens -> Linear -> sigmoid."""
    emb = mha_pool(frame[:, 2:], sd, "at_adpater.0.", 12)
    return torch.sigmoid(F.linear(emb, sd["at_adpater.1.weight"], sd["at_adpater.1.bias"]))


def mlm_head(x, sd):
    """passt_sed.py:194-196: Linear-GELU-Linear."""
    return linear(F.gelu(linear(x, sd, "mlm_mlp.0")), sd, "mlm_mlp.2")


def block_mask_from_noise(noise, mask_rate, block_width, seq_len):
    """mask.py:93-100 with the `torch.rand` draw injected: threshold at sorted-noise index int(n*rate)."""
    n_seg = noise.shape[1]
    thr = noise.sort()[0][:, min(int(n_seg * mask_rate), n_seg - 1)]
    m = torch.zeros(noise.shape[0], seq_len, dtype=torch.bool, device=noise.device)
    m[:, :n_seg * block_width] = (noise <= thr.unsqueeze(-1)).repeat_interleave(block_width, dim=1)
    return m


def apply_mask(x, mask_id, probs, rand_idx, mask_token, style=(0.8, 0.1, 0.1), strict_upstream=False):
    """mask.py:62-83 with RNG draws injected.  strict_upstream=True reproduces the upstream no-op for
    non-contiguous inputs: the masked sequence equals the input."""
    if strict_upstream:
        return x
    B, T, C = x.shape
    flat = x.reshape(-1, C)
    m = mask_id.reshape(-1)
    out = flat.clone()
    mm = m & (probs < style[0])
    out[mm] = mask_token.reshape(1, C).to(out.dtype)
    rm = m & (probs >= style[0]) & (probs < style[0] + style[1])
    out[rm] = flat[rand_idx[: int(rm.sum())]]
    return out.reshape(B, T, C)


def window_starts(input_len, win_param):
    """encoder_slide_window.py:29-30: [(w_left, w_right)] of the Python loop."""
    win, step = win_param
    return [(w, min(w + win, input_len)) for w in range(0, input_len + step - win, step)]


def slide_window_embed(mel, sd, win_param, emb_len, feature_layer=10, f_pool_mode="mean_pool", ratio=10, t_offsets=None):
    """EncoderSlideWindow.__call__ (encoder_slide_window.py:16-36) with PasstWithSlide.encode (passt_win.py:23-41): one
    backbone pass per window, x`ratio` linear interpolation WITHOUT the last-frame pad, overlap-add mean, nan -> 0.
    `t_offsets[i]` injects window i's train-mode time-table offset (passt.py:506-509); None = eval (0)."""
    B, _, L = mel.shape
    scale = emb_len / L
    emb = acc = None
    for i, (w_left, w_right) in enumerate(window_starts(L, win_param)):
        feat, _, Fd, Td = passt_backbone(mel[:, :, w_left:w_right], sd, feature_layer=feature_layer,
                                         t_offset=0 if t_offsets is None else t_offsets[i])
        out = f_pool(feat, sd, Fd, Td, f_pool_mode)
        if ratio != 1:
            out = F.interpolate(out.transpose(1, 2), scale_factor=ratio, mode="linear").transpose(1, 2)
        if emb is None:
            emb = torch.zeros(B, emb_len, out.shape[-1], dtype=out.dtype, device=out.device)
            acc = torch.zeros_like(emb)
        out_left = round(w_left * scale)
        out_right = int(min(emb_len, out_left + out.shape[1]))
        emb[:, out_left:out_right] = emb[:, out_left:out_right] + out
        acc[:, out_left:out_right] += 1
    emb = emb / acc
    return torch.where(torch.isnan(emb), torch.zeros_like(emb), emb)


def mat_sed_forward(mel, sd, decoder_layers=3, feature_layer=10, f_pool_mode="mean_pool", decode_ratio=10,
                    temp_w=1.0, pad_mask=None, mlm=False, decoder_input_override=None, stages=None,
                    encoder_win=False, mix_rate=0.5, win_param=(512, 49), win_t_offsets=None):
    """PaSST_SED.forward (passt_sed.py:242-296).  Returns (strong, weak, other) or (pred, other) in MLM mode.
    `stages` (dict) receives intermediate tensors when given.  encoder_win: sliding-window fusion (:266-271)."""
    feat, frame, Fd, Td = passt_backbone(mel, sd, feature_layer=feature_layer)
    x = pad_interpolate(f_pool(feat, sd, Fd, Td, f_pool_mode), decode_ratio)
    if encoder_win:
        x_local = slide_window_embed(mel, sd, win_param, x.shape[1], feature_layer, f_pool_mode, decode_ratio, win_t_offsets)
        if stages is not None:
            stages.update(x_global=x, x_local=x_local)
        x = mix_rate * x_local + (1 - mix_rate) * x
    other = {"frame_before_mask": x}
    dec_in = x if decoder_input_override is None else decoder_input_override(x, other)
    y = txl_decoder(dec_in, sd, decoder_layers)
    if "at_adpater.1.weight" in sd:
        other["at_out"] = at_branch(frame, sd)
    if stages is not None:
        stages.update(layer_feat=feat, frame=frame, frame_before_mask=x, decoder_out=y)
    if mlm:
        return mlm_head(y, sd), other
    strong, weak = sed_head(y, sd, temp_w, pad_mask)
    return strong, weak, other


# ------------------------------------------------------------------------------------------------
# PMAM: CNN branch, PaSST_CNN forward, prototype head
# ------------------------------------------------------------------------------------------------
def cnn_forward(x, sd, nb_filters, pooling, training=False, p="cnn.cnn.", new_stats=None):
    """cnn/base.py:33-113 with activation 'cg' (ContextGating :19-30), BatchNorm2d(eps 1e-3, momentum .99), AvgPool2d; dropout
    is not applied (parity runs use eval mode or conv_dropout=0).  x [B, 1, T, F].  `new_stats` (dict) receives the running
    statistics a training-mode pass leaves behind."""
    for i in range(len(nb_filters)):
        x = F.conv2d(x, sd[f"{p}conv{i}.weight"], sd[f"{p}conv{i}.bias"], padding=1)
        rm, rv = sd[f"{p}batchnorm{i}.running_mean"].clone(), sd[f"{p}batchnorm{i}.running_var"].clone()
        x = F.batch_norm(x, rm, rv, sd[f"{p}batchnorm{i}.weight"], sd[f"{p}batchnorm{i}.bias"], training, 0.99, 0.001)
        if new_stats is not None:
            new_stats[f"{p}batchnorm{i}.running_mean"], new_stats[f"{p}batchnorm{i}.running_var"] = rm, rv
        lin = F.linear(x.permute(0, 2, 3, 1), sd[f"{p}cg{i}.linear.weight"], sd[f"{p}cg{i}.linear.bias"]).permute(0, 3, 1, 2)
        x = x * torch.sigmoid(lin)
        x = F.avg_pool2d(x, tuple(pooling[i]))
    return x


def passt_cnn_forward(mel, sd, nb_filters, pooling, decoder_layers=3, feature_layer=10, f_pool_mode="attention", decode_ratio=10,
                      training=False, decoder_input_override=None, stages=None, new_stats=None, mlm=True, temp_w=1.0, pad_mask=None):
    """PaSST_CNN.forward (cnn_transformer/passt_cnn.py:32-88), encoder_win=False -> (pred [B,T,out_dim], other) with mlm, else
    (strong [B,C,T], weak [B,C], other) through the classifier + linear-softmax pooling (:73-86)."""
    feat, frame, Fd, Td = passt_backbone(mel, sd, feature_layer=feature_layer)
    x = pad_interpolate(f_pool(feat, sd, Fd, Td, f_pool_mode), decode_ratio)
    cnn_feat = cnn_forward(mel.transpose(1, 2).unsqueeze(1), sd, nb_filters, pooling, training, new_stats=new_stats)
    cnn_feat = F.interpolate(cnn_feat.squeeze(-1), size=x.shape[1], mode="linear").transpose(1, 2)
    x = linear(x, sd, "transformer_projector") + sd["merge_weight"] * linear(cnn_feat, sd, "cnn_projector")
    other = {"frame_before_mask": x}
    dec_in = x if decoder_input_override is None else decoder_input_override(x, other)
    y = txl_decoder(dec_in, sd, decoder_layers)
    other["at_out"] = at_branch(frame, sd)
    if stages is not None:
        stages.update(cnn_feat=cnn_feat, frame_before_mask=x, decoder_in=dec_in, decoder_out=y)
    if mlm:
        return mlm_head(y, sd), other
    strong, weak = sed_head(y, sd, temp_w, pad_mask)
    return strong, weak, other


def prototype_predict(logit, prototypes, temperature=0.1):
    """recipes/desed/pmam/train.py:82-87."""
    s = F.normalize(logit, dim=-1) @ prototypes.T
    return torch.sigmoid((F.leaky_relu(s, negative_slope=0.2) * 2 - 1) / temperature)


# ------------------------------------------------------------------------------------------------
# DASM: query-based open-vocabulary detection  (reference src/models/detect_any_sound/)
# ------------------------------------------------------------------------------------------------
def mha(q_in, kv_in, sd, p, num_heads, attn_mask=None):
    """nn.MultiheadAttention(batch_first=True) forward, eval mode (no dropout): packed in_proj, scaled dot-product with an optional
    boolean attn_mask [Nq, Nk] (True = masked out), out_proj."""
    B, Nq, D = q_in.shape
    Nk, hd = kv_in.shape[1], D // num_heads
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(q_in, w[:D], b[:D]).reshape(B, Nq, num_heads, hd).transpose(1, 2)
    k = F.linear(kv_in, w[D:2 * D], b[D:2 * D]).reshape(B, Nk, num_heads, hd).transpose(1, 2)
    v = F.linear(kv_in, w[2 * D:], b[2 * D:]).reshape(B, Nk, num_heads, hd).transpose(1, 2)
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask, float("-inf"))
    o = (s.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, Nq, D)
    return F.linear(o, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def at_decoder(queries, memory, sd, n_layers, num_heads, tgt_mask=None, p="at_decoder.decoder.layers."):
    """at_adapter.py:7-50: nn.TransformerDecoder of post-norm layers that cross-attend FIRST (:28-31), GELU FFN, eps 1e-5."""
    x = queries
    for i in range(n_layers):
        q = f"{p}{i}."
        x = layer_norm(x + mha(x, memory, sd, q + "multihead_attn.", num_heads), sd, q + "norm1", 1e-5)
        x = layer_norm(x + mha(x, x, sd, q + "self_attn.", num_heads, tgt_mask), sd, q + "norm2", 1e-5)
        x = layer_norm(x + linear(F.gelu(linear(x, sd, q + "linear1")), sd, q + "linear2"), sd, q + "norm3", 1e-5)
    return x


def mlp_head(x, sd, p, n_layers):
    """detect_any_sound.py:404-416 (`MLP`): Linear (+GELU) x n."""
    for i in range(n_layers):
        x = linear(x, sd, f"{p}layers.{i}")
        if i < n_layers - 1:
            x = F.gelu(x)
    return x


def dasm_forward(mel, query, sd, nb_filters, pooling, at_layers=2, decoder_layers=3, num_heads=12, feature_layer=10, ratio=10, temp_w=0.1,
                 pad_mask=None, tgt_mask=None, training=False, stages=None):
    """DASM.forward (detect_any_sound.py:304-389): query_projector given, out_type 'sigmoid', no MLM.  query [K, query_dim]."""
    feat, frame, Fd, Td = passt_backbone(mel, sd, feature_layer=feature_layer)
    # f_pool (:242-254): norm_before_pool on the patch tokens, attention pooling (6 heads) over frequency
    y = layer_norm(feat[:, 2:], sd, "norm_before_pool", 1e-5)
    B, _, C = y.shape
    y = y.reshape(B, Fd, Td, C).transpose(1, 2).reshape(B * Td, Fd, C)
    x = pad_interpolate(mha_pool(y, sd, "f_pool_module.", 6).reshape(B, Td, C), ratio)
    cnn_feat = cnn_forward(mel.transpose(1, 2).unsqueeze(1), sd, nb_filters, pooling, training)
    cnn_feat = F.interpolate(cnn_feat.squeeze(-1), size=x.shape[1], mode="linear").transpose(1, 2)
    x = linear(x, sd, "transformer_projector") + sd["merge_weight"] * linear(cnn_feat, sd, "cnn_projector")
    x = layer_norm(x, sd, "norm_after_merge", 1e-5)
    at_feat = linear(frame[:, 2:], sd, "at_projector")
    q = F.gelu(linear(query, sd, "query_projector.0"))
    mask_feat = at_decoder(q.expand(B, -1, -1), at_feat, sd, at_layers, num_heads, tgt_mask)
    at_out = torch.sigmoid(mlp_head(mask_feat, sd, "at_head.", 2).squeeze(-1))
    xd = linear(txl_decoder(x, sd, decoder_layers, num_heads, p="sed_decoder."), sd, "sed_head")
    emb = mlp_head(mask_feat, sd, "mask_embedding_layer.", 3)
    score = torch.einsum("bqc,bct->bqt", emb, xd.transpose(1, 2)).transpose(1, 2)          # [B, T, K]
    sed = torch.sigmoid(score / temp_w) * at_out.unsqueeze(1)
    if pad_mask is not None:
        sed = sed.masked_fill(pad_mask.unsqueeze(-1), 0.0)
    sed = torch.clamp(sed, 1e-7, 1.0)
    weak = torch.clamp((sed * sed).sum(dim=1) / sed.sum(dim=1), 1e-7, 1.0)
    if stages is not None:
        stages.update(x_merged=x, mask_feat=mask_feat, score=score)
    return sed.transpose(1, 2), weak, {"at_out": at_out}


def bce(p, y):
    """torch.nn.BCELoss semantics (log clamped at -100), mean reduction (finetune/train.py:166-173)."""
    return -(y * torch.clamp(p.log(), min=-100.0) + (1 - y) * torch.clamp((1 - p).log(), min=-100.0)).mean()
