"""CPU: oracle MAT-SED model vs golden vectors from the unmodified reference (small + base + MLM + ops)."""
import numpy as np
import pytest
import torch

from conftest import checksum
from oracle import frontend as F
from oracle import model as M
from transformer4sed_b200 import schema
from transformer4sed_b200.utils import synth

torch.set_num_threads(max(1, __import__("os").cpu_count()))


def _sd(shapes, seed, grad=False):
    sd = synth.synth_state_dict(shapes, seed)
    if grad:
        for v in sd.values():
            v.requires_grad_(True)
    return sd


def _run(tag, shapes, seed, batch, decoder_layers, golden):
    g = golden(f"matsed_{tag}.npz")
    sd = _sd(shapes, seed, grad=True)
    np.testing.assert_allclose(checksum(torch.cat([v.flatten() for _, v in sorted(sd.items())])), g["sd_ck"], rtol=1e-12)
    wav = synth.synth_wav(batch, 320000, seed=seed + 1)
    np.testing.assert_allclose(checksum(wav), g["wav_ck"], rtol=1e-12)
    mel = F.passt_logmel(wav)
    np.testing.assert_allclose(checksum(mel), g["mel_ck"], rtol=1e-9)
    labels = synth.synth_strong_labels(batch, 10, 1000, seed + 2)
    weak_labels = (labels.sum(-1) > 0).float()
    st = {}
    strong, weak, other = M.mat_sed_forward(mel, sd, decoder_layers=decoder_layers, stages=st)
    tol = dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(st["layer_feat"][:, ::17, ::4].detach().numpy(), g["layer_feat"], **tol)
    np.testing.assert_allclose(st["frame"][:, ::17, ::4].detach().numpy(), g["frame"], **tol)
    np.testing.assert_allclose(other["frame_before_mask"][:, ::8, ::4].detach().numpy(), g["frame_before_mask"], **tol)
    np.testing.assert_allclose(st["decoder_out"][:, ::8, ::4].detach().numpy(), g["decoder_out"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(strong.detach().numpy(), g["strong"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(weak.detach().numpy(), g["weak"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(other["at_out"].detach().numpy(), g["at_out"], rtol=1e-4, atol=1e-6)
    assert (strong.argmax(dim=1).numpy() == g["argmax"]).all()
    loss = M.bce(strong, labels) + 0.5 * M.bce(weak, weak_labels) + 2.0 * M.bce(other["at_out"], weak_labels)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)
    loss.backward()
    for name, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
        gr = sd[str(name)].grad
        assert gr is not None, name
        np.testing.assert_allclose(gr.double().norm().item(), norm, rtol=2e-3, atol=1e-7, err_msg=str(name))
        n = min(8, gr.numel())
        np.testing.assert_allclose(gr.flatten()[:n].double().numpy(), head[:n], rtol=5e-3, atol=1e-5 * max(norm, 1e-3), err_msg=str(name))
    with torch.no_grad():
        pad = torch.zeros(batch, 1000, dtype=torch.bool)
        pad[-1, 900:] = True
        sp, wp, _ = M.mat_sed_forward(mel, sd, decoder_layers=decoder_layers, pad_mask=pad)
        np.testing.assert_allclose(sp.numpy(), g["strong_pad"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(wp.numpy(), g["weak_pad"], rtol=1e-4, atol=1e-6)
        s5, w5, _ = M.mat_sed_forward(mel, sd, decoder_layers=decoder_layers, temp_w=0.5)
        np.testing.assert_allclose(s5.numpy(), g["strong_t05"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(w5.numpy(), g["weak_t05"], rtol=1e-4, atol=1e-6)


def test_small_config(golden):
    _run("small", schema.mat_sed_shapes(embed_dim=192, decoder_dim=192, decoder_layer_num=1), 3, 2, 1, golden)


def test_base_config(golden):
    _run("base", schema.mat_sed_shapes(), 4, 1, 3, golden)


@pytest.mark.parametrize("batch", [1, 2])
def test_mlm_pretrain(golden, batch):
    """B>1: upstream masking is a silent no-op; B=1: masking applies (refined: the
    `reshape(-1,C)` of the cloned non-contiguous tensor is a view only when B==1)."""
    g = golden(f"matsed_mlm_base_b{batch}.npz")
    seed = 6
    sd = _sd(schema.mat_sed_shapes(mlm=True), seed)
    wav = synth.synth_wav(batch, 320000, seed=seed + 1)
    np.testing.assert_allclose(checksum(wav), g["wav_ck"], rtol=1e-12)
    mel = F.passt_logmel(wav)
    noise = torch.from_numpy(g["noise"])
    mask = M.block_mask_from_noise(noise, 0.75, 10, 1000)
    assert (np.packbits(mask.numpy()) == g["mask"]).all()
    noop = bool(g["decoder_in_equals_input"])
    assert noop == (batch > 1)

    def dec_in(x, other):
        return M.apply_mask(x, mask, torch.from_numpy(g["probs"]), torch.from_numpy(g["rand_idx"]).long(),
                            sd["mask_token"], style=(0.8, 0.1, 0.1), strict_upstream=noop)

    with torch.no_grad():
        pred, other = M.mat_sed_forward(mel, sd, mlm=True, decoder_input_override=dec_in)
        np.testing.assert_allclose(dec_in(other["frame_before_mask"], None)[:, ::8, ::4].numpy(), g["decoder_in"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(pred[:, ::8, ::4].numpy(), g["pred"], rtol=1e-3, atol=1e-3)
        fbm = other["frame_before_mask"]
        loss = torch.nn.functional.mse_loss(fbm[mask], pred[mask])
        np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-4)


def test_ops(golden):
    g = golden("ops.npz")
    np.testing.assert_array_equal(M.rel_shift(torch.from_numpy(g["rel_shift_in"])).numpy(), g["rel_shift_out"])
    m = M.block_mask_from_noise(torch.from_numpy(g["block_noise"]), 0.75, 10, 1000)
    assert (np.packbits(m.numpy()) == g["block_mask"]).all()
    sd = synth.synth_state_dict(schema.txl_decoder_shapes(48, 2, num_heads=4, prefix=""), 9)
    xin = synth.synth_tensor(9, "txl_in", (3, 50, 48))
    np.testing.assert_allclose(checksum(xin), g["txl_in_ck"], rtol=1e-12)
    out = M.txl_decoder(xin, sd, 2, num_heads=4, p="")
    np.testing.assert_allclose(out.numpy(), g["txl_out"], rtol=1e-4, atol=1e-5)


def test_sliding_window(golden):
    """Oracle sliding-window fusion (encoder_slide_window.py + passt_win.py) vs the unmodified reference: validation kwargs
    (eval, [512, 31], temp 0.5) and the teacher's training kwargs (train mode, [512, 49], recorded RNG offsets)."""
    g = golden("matsed_window_base.npz")
    seed = 8
    sd = _sd(schema.mat_sed_shapes(), seed)
    wav = synth.synth_wav(1, 320000, seed=seed + 1)
    np.testing.assert_allclose(checksum(wav), g["wav_ck"], rtol=1e-12)
    mel = F.passt_logmel(wav)
    assert [b - a for a, b in M.window_starts(1000, (512, 31))][-2:] == [512, 504] and len(M.window_starts(1000, (512, 31))) == 17
    assert len(M.window_starts(1000, (512, 49))) == 11
    with torch.no_grad():
        st = {}
        s, w, _ = M.mat_sed_forward(mel, sd, encoder_win=True, win_param=(512, 31), temp_w=0.5, stages=st)
        np.testing.assert_allclose(st["x_local"][:, ::4, ::4].numpy(), g["local_val"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(s.numpy(), g["strong_val"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(w.numpy(), g["weak_val"], rtol=1e-4, atol=1e-6)
        st = {}
        s, w, _ = M.mat_sed_forward(mel, sd, encoder_win=True, win_param=(512, 49), win_t_offsets=[int(v) for v in g["train_offsets"]], stages=st)
        np.testing.assert_allclose(st["x_local"][:, ::4, ::4].numpy(), g["local_train"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(s.numpy(), g["strong_train"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(w.numpy(), g["weak_train"], rtol=1e-4, atol=1e-6)


def _pmam_mask(seed, batch):
    """Replay MlmModule's CPU draws (mask.py:62-100) for config/pmam/post_pretrain.yaml: block masking, rate .8, style (.9, .05, .05)."""
    torch.manual_seed(seed)
    noise = torch.rand(batch, 100)
    mask = M.block_mask_from_noise(noise, 0.8, 10, 1000)
    probs = torch.rand(batch * 1000)
    n_rand = int((mask.view(-1) & (probs >= 0.9) & (probs < 0.95)).sum())
    rand_idx = torch.randint(0, batch * 1000, (n_rand,))
    return mask, probs, rand_idx


def test_pmam_passt_cnn(golden):
    """Oracle PaSST_CNN (PaSST + LoRA, CNN branch, attention f_pool, TXL d=384, MLM + prototype head) vs the unmodified reference:
    eval-mode forward, and train-mode forward + loss + gradients of every trainable tensor (conv_dropout = 0)."""
    g = golden("pmam_base.npz")
    seed, batch = 10, 2
    shapes = schema.passt_cnn_shapes()
    assert sorted(shapes) == [k for k in g["sd_keys"] if not str(k).endswith("num_batches_tracked")]
    sd = _sd(shapes, seed)
    np.testing.assert_allclose(checksum(torch.cat([v.flatten() for _, v in sorted(sd.items())])), g["sd_ck"], rtol=1e-12)
    trainable = set(str(k) for k in g["trainable"])
    for k, v in sd.items():
        v.requires_grad_(k in trainable)
    wav = synth.synth_wav(batch, 320000, seed=seed + 1)
    mel = F.passt_logmel(wav)
    np.testing.assert_allclose(checksum(mel), g["mel_ck"], rtol=1e-9)
    protos = torch.nn.functional.normalize(synth.synth_tensor(seed, "prototypes", (30, 768)), dim=-1)
    labels = synth.synth_strong_labels(batch, 30, 1000, seed + 2)
    weak_labels = (labels.sum(-1) >= 1).float()

    def run(mask_seed, training, tag):
        mask, probs, ridx = _pmam_mask(mask_seed, batch)
        assert (np.packbits(mask.numpy()) == g[f"{tag}_mask"]).all()
        st, stats = {}, {}
        pred, other = M.passt_cnn_forward(mel, sd, schema.PMAM_FILTERS, schema.PMAM_POOLING, training=training, stages=st, new_stats=stats,
                                          decoder_input_override=lambda x, o: M.apply_mask(x, mask, probs, ridx, sd["mask_token"], style=(0.9, 0.05, 0.05)))
        np.testing.assert_allclose(st["decoder_in"][:, ::8, ::4].detach().numpy(), g[f"{tag}_dec_in"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(pred[:, ::8, ::4].detach().numpy(), g[f"{tag}_pred"], rtol=1e-3, atol=1e-3)
        np.testing.assert_allclose(other["at_out"].detach().numpy(), g[f"{tag}_at"], rtol=1e-4, atol=1e-6)
        return pred, other, mask, stats

    with torch.no_grad():
        pred, other, _, _ = run(seed + 3, False, "eval")
        np.testing.assert_allclose(M.prototype_predict(pred, protos)[:, ::8].numpy(), g["eval_strong"], rtol=1e-3, atol=2e-4)
    pred, other, mask, stats = run(seed + 4, True, "train")
    strong = M.prototype_predict(pred, protos)
    loss = M.bce(strong[mask], labels.transpose(1, 2)[mask]) + 0.5 * M.bce(other["at_out"], weak_labels)
    np.testing.assert_allclose(loss.item(), g["train_loss"], rtol=1e-4)
    for k in ("cnn.cnn.batchnorm0.running_mean", "cnn.cnn.batchnorm9.running_var"):
        np.testing.assert_allclose(stats[k].numpy(), g["bn_" + k], rtol=1e-4, atol=1e-6)
    loss.backward()
    assert sorted(str(n) for n in g["grad_names"]) == sorted(k for k in trainable if sd[k].grad is not None)
    for name, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
        gr = sd[str(name)].grad
        np.testing.assert_allclose(gr.double().norm().item(), norm, rtol=5e-3, atol=1e-7, err_msg=str(name))


def test_dasm(golden):
    """Oracle DASM (query projector, cross-attention-first decoder with boolean tgt_mask, query x frame scores, sigmoid * at_out) vs
    the unmodified reference: eval forward (plain / masked), train-mode forward + loss + gradients (dropouts 0)."""
    g = golden("dasm_base.npz")
    seed, batch, K = 12, 2, 407
    shapes = schema.dasm_shapes()
    assert sorted(shapes) == [k for k in g["sd_keys"] if not str(k).endswith("num_batches_tracked")]
    sd = _sd(shapes, seed)
    np.testing.assert_allclose(checksum(torch.cat([v.flatten() for _, v in sorted(sd.items())])), g["sd_ck"], rtol=1e-12)
    trainable = set(str(k) for k in g["trainable"])
    for k, v in sd.items():
        v.requires_grad_(k in trainable)
    wav = synth.synth_wav(batch, 320000, seed=seed + 1)
    mel = F.passt_logmel(wav)
    query = torch.nn.functional.normalize(synth.synth_tensor(seed, "queries", (K, 768)), dim=-1) * 3.0
    np.testing.assert_allclose(checksum(query), g["query_ck"], rtol=1e-12)
    tgt_mask = torch.from_numpy(np.unpackbits(g["tgt_mask"])[:K * K].astype(bool)).view(K, K)
    pad = torch.zeros(batch, 1000, dtype=torch.bool)
    pad[-1, 900:] = True
    kw = dict(nb_filters=schema.PMAM_FILTERS, pooling=schema.PMAM_POOLING, temp_w=4.0)
    with torch.no_grad():
        s, w, o = M.dasm_forward(mel, query, sd, **kw)
        np.testing.assert_allclose(s[:, ::3, ::4].numpy(), g["eval_strong"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(w.numpy(), g["eval_weak"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(o["at_out"].numpy(), g["eval_at"], rtol=1e-4, atol=1e-6)
        assert (s.argmax(dim=1).numpy() == g["eval_argmax"]).mean() > 0.999
        s, w, o = M.dasm_forward(mel, query, sd, pad_mask=pad, tgt_mask=tgt_mask, **kw)
        np.testing.assert_allclose(s[:, ::3, ::4].numpy(), g["evalm_strong"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(w.numpy(), g["evalm_weak"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(o["at_out"].numpy(), g["evalm_at"], rtol=1e-4, atol=1e-6)
    labels = synth.synth_strong_labels(batch, K, 1000, seed + 2)
    weak_labels = (labels.sum(-1) >= 1).float()
    s, w, o = M.dasm_forward(mel, query, sd, training=True, **kw)
    np.testing.assert_allclose(s[:, ::3, ::4].detach().numpy(), g["train_strong"], rtol=1e-3, atol=1e-6)
    loss = M.bce(s, labels) + 0.5 * M.bce(w, weak_labels) + 0.5 * M.bce(o["at_out"], weak_labels)
    np.testing.assert_allclose(loss.item(), g["train_loss"], rtol=1e-4)
    loss.backward()
    for name, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
        gr = sd[str(name)].grad
        assert gr is not None, name
        if str(name).startswith("cnn.cnn.conv") and str(name).endswith(".bias"):
            continue    # zero gradient in front of a training-mode BatchNorm (rounding noise only)
        np.testing.assert_allclose(gr.double().norm().item(), norm, rtol=5e-3, atol=1e-9, err_msg=str(name))


def test_pmam_finetune_passt_cnn(golden):
    """Oracle PaSST_CNN in its fine-tuning form (config/pmam/finetune1.yaml: no LoRA, no MLM, 10 classes): strong / weak / AT outputs,
    frame-label argmax and the pad-mask + temperature path vs the unmodified reference."""
    g = golden("pmam_finetune.npz")
    seed, batch = 14, 2
    shapes = schema.passt_cnn_shapes(class_num=10, mlm=False, lora_r=0)
    assert sorted(shapes) == [k for k in g["sd_keys"] if not str(k).endswith("num_batches_tracked")]
    sd = _sd(shapes, seed)
    np.testing.assert_allclose(checksum(torch.cat([v.flatten() for _, v in sorted(sd.items())])), g["sd_ck"], rtol=1e-12)
    mel = F.passt_logmel(synth.synth_wav(batch, 320000, seed=seed + 1))
    np.testing.assert_allclose(checksum(mel), g["mel_ck"], rtol=1e-9)
    pad = torch.zeros(batch, 1000, dtype=torch.bool)
    pad[-1, 850:] = True
    with torch.no_grad():
        s, w, o = M.passt_cnn_forward(mel, sd, schema.PMAM_FILTERS, schema.PMAM_POOLING, mlm=False)
        np.testing.assert_allclose(s.numpy(), g["strong"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(w.numpy(), g["weak"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(o["at_out"].numpy(), g["at_out"], rtol=1e-4, atol=1e-6)
        assert (s.argmax(dim=1).numpy() == g["argmax"]).all()
        s, w, _ = M.passt_cnn_forward(mel, sd, schema.PMAM_FILTERS, schema.PMAM_POOLING, mlm=False, temp_w=0.5, pad_mask=pad)
        np.testing.assert_allclose(s.numpy(), g["strong_pad"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(w.numpy(), g["weak_pad"], rtol=1e-4, atol=1e-6)
