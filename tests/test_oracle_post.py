"""CPU: oracle restatements of TorchScaler, the event decoding and the mean-teacher loss block vs golden vectors recorded from the
unmodified reference (oracle/make_golden.py post), plus the host logic of the checkpoint hand-off helpers."""
import numpy as np
import torch

from conftest import checksum
from oracle import glue as G
from transformer4sed_b200.utils import synth


def _feat(g):
    x = synth.synth_tensor(41, "post_feat", (5, 128, 250)) * 3.0 + 1.5
    np.testing.assert_allclose(checksum(x), g["feat_ck"], rtol=1e-12)
    return x


def decode_inputs(g, tag, T):
    B = 4
    smooth = torch.nn.functional.avg_pool1d(synth.synth_tensor(43, f"post_strong{tag}", (B, 10, T + 8)), 9, 1)
    strong = torch.sigmoid(6.0 * smooth)
    weak = torch.sigmoid(2.0 * synth.synth_tensor(43, f"post_weak{tag}", (B, 10)))
    np.testing.assert_allclose(checksum(strong), g[f"strong{tag}_ck"], rtol=1e-12)
    np.testing.assert_allclose(checksum(weak), g[f"weak{tag}_ck"], rtol=1e-12)
    return strong, weak


def loss_inputs():
    B, C, T = 12, 10, 1000
    mk = lambda name, shape: torch.sigmoid(synth.synth_tensor(47, name, shape))  # noqa: E731
    stu = [mk("l_ss", (B, C, T)), mk("l_sw", (B, C)), mk("l_sa", (B, C))]
    tch = [mk("l_ts", (B, C, T)), mk("l_ta", (B, C))]
    y = (synth.synth_tensor(47, "l_y", (B, C, T)) > 0.5).float()
    yw = (synth.synth_tensor(47, "l_yw", (B, C)) > 0.3).float()
    return stu, tch, y, yw


def frame_to_time(frames, net_pooling, hop, sr, audio_len=10):
    return np.clip(frames * net_pooling * hop / sr, a_min=0, a_max=audio_len)     # src/codec/encoder.py:26-28


def test_scaler_oracle_matches_reference(golden):
    g = golden("post.npz")
    x = _feat(g)
    for nt in ("mean", "standard", "minmax"):
        r = G.torch_scaler(x, "instance", nt)
        np.testing.assert_allclose(r[:, ::4, ::5].numpy(), g[f"scaler_instance_{nt}"], rtol=1e-6, atol=1e-7)
    mean, msq = torch.from_numpy(g["scaler_mean"]), torch.from_numpy(g["scaler_mean_squared"])
    for nt in ("standard", "mean"):
        r = G.torch_scaler(x, "dataset", nt, mean=mean, mean_squared=msq)
        np.testing.assert_allclose(r[:, ::4, ::5].numpy(), g[f"scaler_dataset_{nt}"], rtol=1e-6, atol=1e-7)


def test_decoding_oracle_matches_reference(golden):
    g = golden("post.npz")
    for tag, T, grid in (("156", 156, (4, 256, 16000)), ("1000", 1000, (1, 320, 32000))):
        strong, weak = decode_inputs(g, tag, T)
        sizes = [int(k) for k in g[f"sizes{tag}"]]
        ev = G.decode_pred_batch_fast(strong, weak, [0.25, 0.5, 0.75], sizes)
        np.testing.assert_array_equal(ev[:, :3], g[f"events{tag}_idx"])
        np.testing.assert_array_equal(frame_to_time(ev[:, 3:].astype(np.float64), *grid), g[f"events{tag}_time"])
        for ft in ("median", "max"):
            for j in range(2):
                scores = (strong[j].transpose(0, 1) * weak[j, :]).numpy()
                np.testing.assert_array_equal(G.rank_filter_scores(scores, sizes, ft), g[f"scores{tag}_{ft}"][j])


def test_losses_oracle_matches_reference(golden):
    g = golden("post.npz")
    stu, tch, y, yw = loss_inputs()
    stu = [t.requires_grad_() for t in stu]
    ms, mw = torch.zeros(12).bool(), torch.zeros(12).bool()
    ms[:4] = 1
    mw[4:8] = 1
    total, parts = G.sed_losses(*stu, *tch, y, yw, ms, mw, *[float(v) for v in g["loss_weights"]])
    total.backward()
    np.testing.assert_allclose(total.item(), g["loss_total"], rtol=1e-6)
    np.testing.assert_allclose([p.item() for p in parts], g["loss_parts"], rtol=1e-6)
    np.testing.assert_allclose(stu[0].grad[:, ::3, ::25].numpy(), g["d_strong"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(stu[1].grad.numpy(), g["d_weak"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(stu[2].grad.numpy(), g["d_at"], rtol=1e-5, atol=1e-10)


def test_checkpoint_stage_filters():
    """recipes/desed/finetune/passt/main.py:60-64 and recipes/desed/pmam/main.py:188-191, plus the DataParallel prefix."""
    from transformer4sed_b200.utils import checkpoint as ck
    sd = {f"module.{k}": torch.zeros(1) for k in ("backbone.blocks.0.attn.qkv.weight", "classifier.weight", "at_adpater.1.weight",
                                                   "at_adpater.0.weight", "mlm_mlp.0.weight", "decoder.x")}
    plain = ck.strip_data_parallel(sd)
    assert set(plain) == {k[len("module."):] for k in sd}
    assert ck.strip_data_parallel(plain) == plain                  # no prefix: untouched
    mixed = dict(plain, **{"module.extra": torch.zeros(1)})
    assert set(ck.strip_data_parallel(mixed)) == set(mixed)        # torch prefixes every key or none
    ft = ck.filter_stage_keys(plain, "finetune_from_mlm")
    assert set(ft) == {"backbone.blocks.0.attn.qkv.weight", "at_adpater.0.weight", "mlm_mlp.0.weight", "decoder.x"}
    pm = ck.filter_stage_keys(plain, "pmam_from_existing")
    assert "mlm_mlp.0.weight" not in pm and "classifier.weight" in pm
    net = torch.nn.Sequential(torch.nn.Linear(3, 2))
    src = {"module.0.weight": torch.ones(2, 3), "module.0.bias": torch.ones(2), "module.classifier.weight": torch.ones(1)}
    res = ck.load_reference_checkpoint(net, src, stage="finetune_from_mlm")
    assert not res.missing_keys and not res.unexpected_keys and torch.equal(net[0].weight, torch.ones(2, 3))
    import pytest
    with pytest.raises(ValueError):
        ck.filter_stage_keys(plain, "nope")
