"""State-dict schema of the reference models as {key: shape}.

The drop-in promise is that checkpoints move between the reference and this package unchanged, so
the key names (including the upstream spelling ``at_adpater``) and shapes are part of the boundary.
"""


def _ln(d, p, dim):
    d[p + ".weight"] = (dim,)
    d[p + ".bias"] = (dim,)


def _lin(d, p, n_out, n_in, bias=True):
    d[p + ".weight"] = (n_out, n_in)
    if bias:
        d[p + ".bias"] = (n_out,)


def passt_shapes(embed_dim=768, depth=12, mlp_ratio=4, f_dim=12, t_dim=99, patch=16, num_classes=527, prefix="backbone."):
    """reference src/models/passt/passt.py:366-455 (distilled=True)."""
    d, D, p = {}, embed_dim, prefix
    d[p + "cls_token"] = (1, 1, D)
    d[p + "dist_token"] = (1, 1, D)
    d[p + "new_pos_embed"] = (1, 2, D)
    d[p + "freq_new_pos_embed"] = (1, D, f_dim, 1)
    d[p + "time_new_pos_embed"] = (1, D, 1, t_dim)
    d[p + "patch_embed.proj.weight"] = (D, 1, patch, patch)
    d[p + "patch_embed.proj.bias"] = (D,)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        _ln(d, b + "norm1", D)
        _lin(d, b + "attn.qkv", 3 * D, D)
        _lin(d, b + "attn.proj", D, D)
        _ln(d, b + "norm2", D)
        _lin(d, b + "mlp.fc1", int(D * mlp_ratio), D)
        _lin(d, b + "mlp.fc2", D, int(D * mlp_ratio))
    _ln(d, p + "norm", D)
    _ln(d, p + "head.0", D)
    _lin(d, p + "head.1", num_classes, D)
    _lin(d, p + "head_dist", num_classes, D)
    return d


def attention_pooling_shapes(prefix, dim):
    """reference src/models/pooling.py:37-44."""
    d = {prefix + "f_att_token": (1, 1, dim),
         prefix + "frequency_att.in_proj_weight": (3 * dim, dim),
         prefix + "frequency_att.in_proj_bias": (3 * dim,)}
    _lin(d, prefix + "frequency_att.out_proj", dim, dim)
    return d


def txl_decoder_shapes(dim, n_layers, num_heads=12, mlp_ratio=1, prefix="decoder."):
    """reference src/models/transformer_decoder.py:74-94, src/models/transformer/transformerXL.py:23-28,148-178."""
    d = {}
    for i in range(n_layers):
        b = f"{prefix}encoder_blocks.{i}."
        _ln(d, b + "norm1", dim)
        _ln(d, b + "norm2", dim)
        d[b + "attn.pos_bias_u"] = (num_heads, dim // num_heads)
        d[b + "attn.pos_bias_v"] = (num_heads, dim // num_heads)
        _lin(d, b + "attn.in_proj", 3 * dim, dim)
        _lin(d, b + "attn.out_proj", dim, dim)
        _lin(d, b + "attn.linear_pos", dim, dim, bias=False)
        _lin(d, b + "mlp.fc1", int(dim * mlp_ratio), dim)
        _lin(d, b + "mlp.fc2", dim, int(dim * mlp_ratio))
    return d


def mat_sed_shapes(embed_dim=768, decoder_dim=768, decoder_layer_num=3, class_num=10, at_adapter=True,
                   f_pool="mean_pool", mlm=False, mlm_out_dim=768):
    """reference src/models/passt/passt_sed.py:39-148 (decoder='transformerXL')."""
    d = passt_shapes(embed_dim)
    _ln(d, "out_norm", embed_dim)
    if f_pool == "attention":
        d.update(attention_pooling_shapes("f_pool_module.", embed_dim))
    if mlm:
        d["mask_token"] = (1, 1, decoder_dim)
        _lin(d, "mlm_mlp.0", decoder_dim, decoder_dim)
        _lin(d, "mlm_mlp.2", mlm_out_dim, decoder_dim)
    d.update(txl_decoder_shapes(decoder_dim, decoder_layer_num))
    _lin(d, "classifier", class_num, decoder_dim)
    if at_adapter:
        d.update(attention_pooling_shapes("at_adpater.0.", embed_dim))
        _lin(d, "at_adpater.1", class_num, embed_dim)
    return d


def cnn_shapes(nb_filters, n_in_channel=1, prefix="cnn.cnn."):
    """reference src/models/cnn/base.py:62-103 (activation 'cg', BatchNorm): conv{i}, batchnorm{i} (+ running stats), cg{i}.linear."""
    d = {}
    for i, n_out in enumerate(nb_filters):
        n_in = n_in_channel if i == 0 else nb_filters[i - 1]
        d[f"{prefix}conv{i}.weight"] = (n_out, n_in, 3, 3)
        d[f"{prefix}conv{i}.bias"] = (n_out,)
        for k in ("weight", "bias", "running_mean", "running_var"):
            d[f"{prefix}batchnorm{i}.{k}"] = (n_out,)
        _lin(d, f"{prefix}cg{i}.linear", n_out, n_out)
    return d


def add_lora(d, r=8, prefix="backbone."):
    """reference src/models/passt/passt_lora.py: every block Linear and both heads carry lora_A [r, in] / lora_B [out, r]."""
    for k in [k for k in d if k.startswith(prefix) and k.endswith(".weight") and len(d[k]) == 2 and
              any(t in k for t in (".attn.qkv.", ".attn.proj.", ".mlp.fc1.", ".mlp.fc2.", "head.1.", "head_dist."))]:
        n_out, n_in = d[k]
        d[k[:-len("weight")] + "lora_A"] = (r, n_in)
        d[k[:-len("weight")] + "lora_B"] = (n_out, r)
    return d


PMAM_FILTERS = (16, 16, 32, 32, 64, 64, 128, 128, 256, 384)
PMAM_POOLING = ((2, 2), (1, 1), (2, 2), (1, 1), (1, 2), (1, 2), (1, 2), (1, 2), (1, 2), (1, 1))


def passt_cnn_shapes(embed_dim=768, decoder_dim=384, decoder_layer_num=3, class_num=30, f_pool="attention", mlm=True, lora_r=8,
                     nb_filters=PMAM_FILTERS):
    # config/pmam/finetune{1,2}.yaml: class_num=10, mlm=False, lora_r=0
    """reference src/models/cnn_transformer/passt_cnn.py:9-19 on top of PaSST_SED (config/pmam/post_pretrain.yaml:48-79)."""
    d = mat_sed_shapes(embed_dim, decoder_dim, decoder_layer_num, class_num, True, f_pool, mlm, 768)
    if lora_r:
        add_lora(d, lora_r)
    d.update(cnn_shapes(nb_filters))
    _lin(d, "cnn_projector", decoder_dim, nb_filters[-1])
    d["merge_weight"] = (1,)
    _lin(d, "transformer_projector", decoder_dim, embed_dim)
    return d


def mha_shapes(prefix, dim):
    d = {prefix + "in_proj_weight": (3 * dim, dim), prefix + "in_proj_bias": (3 * dim,)}
    _lin(d, prefix + "out_proj", dim, dim)
    return d


def dasm_shapes(embed_dim=768, decoder_dim=384, decoder_layer_num=3, at_layers=2, query_dim=768, num_heads=12, expand=1, lora_r=8,
                nb_filters=PMAM_FILTERS):
    """reference src/models/detect_any_sound/detect_any_sound.py:20-232 (query_projector=True, out_type='sigmoid', no MLM, no at_query)."""
    d = passt_shapes(embed_dim)
    if lora_r:
        add_lora(d, lora_r)
    d.update(cnn_shapes(nb_filters))
    D = decoder_dim
    d.update(txl_decoder_shapes(D, decoder_layer_num, num_heads=num_heads, mlp_ratio=expand, prefix="sed_decoder."))
    for i in range(3):
        _lin(d, f"mask_embedding_layer.layers.{i}", D, D)
    _lin(d, "sed_head", D, D)
    _lin(d, "query_projector.0", D, query_dim)
    for i in range(at_layers):
        p = f"at_decoder.decoder.layers.{i}."
        d.update(mha_shapes(p + "self_attn.", D))
        d.update(mha_shapes(p + "multihead_attn.", D))
        _lin(d, p + "linear1", D * expand, D)
        _lin(d, p + "linear2", D, D * expand)
        for n in ("norm1", "norm2", "norm3"):
            _ln(d, p + n, D)
    _lin(d, "at_head.layers.0", D, D)
    _lin(d, "at_head.layers.1", 1, D)
    d.update(attention_pooling_shapes("f_pool_module.", embed_dim))
    _lin(d, "cnn_projector", D, nb_filters[-1])
    d["merge_weight"] = (1,)
    _lin(d, "transformer_projector", D, embed_dim)
    _lin(d, "at_projector", D, embed_dim)
    _ln(d, "norm_before_pool", embed_dim)
    _ln(d, "norm_after_merge", D)
    return d
