"""GPU: TorchScaler, the GPU threshold sweep / event decoding, the scipy-style class-wise filters and the fused six-loss kernel
(csrc/post.cu) against golden vectors of the unmodified reference and the CPU oracle on the same seeded inputs.
Event indices and filtered scores are bit-exact; scaler / loss values within 1e-5 (fp32 reductions in a different order)."""
import numpy as np
import pytest
import torch

from oracle import glue as G
from test_oracle_post import _feat, decode_inputs, frame_to_time, loss_inputs

pytestmark = pytest.mark.gpu


class _Enc:
    """The two members of the reference's ManyHotEncoder the decoder touches (src/codec/encoder.py:9-28)."""

    def __init__(self, net_pooling, hop, sr):
        self.labels = [f"class{i}" for i in range(10)]
        self.audio_len, self.net_pooling, self.frame_hop, self.sr = 10, net_pooling, hop, sr

    def _frame_to_time(self, frame):
        return np.clip(frame * self.net_pooling * self.frame_hop / self.sr, a_min=0, a_max=self.audio_len)


def test_torch_scaler_matches_reference(golden):
    from transformer4sed_b200.src_preprocess.scaler import TorchScaler
    g = golden("post.npz")
    x = _feat(g).cuda()
    for nt in ("mean", "standard", "minmax"):
        r = TorchScaler("instance", nt, dims=(1, 2))(x)
        np.testing.assert_allclose(r[:, ::4, ::5].cpu().numpy(), g[f"scaler_instance_{nt}"], rtol=2e-5, atol=2e-6)
    for nt in ("standard", "mean"):
        sc = TorchScaler("dataset", nt, dims=(1, 2))
        sc.fit([(x[:2],), (x[2:4],), (x[4:],)])
        np.testing.assert_allclose(sc.mean.cpu().numpy(), g["scaler_mean"], rtol=1e-5)
        r = sc(x)
        np.testing.assert_allclose(r[:, ::4, ::5].cpu().numpy(), g[f"scaler_dataset_{nt}"], rtol=2e-5, atol=2e-6)
        sd = sc.state_dict()
        sc2 = TorchScaler("dataset", nt, dims=(1, 2))
        sc2.load_state_dict(sd)
        assert torch.equal(sc2(x), r)
    assert TorchScaler(None, None)(x) is x
    with pytest.raises(NotImplementedError):
        TorchScaler("instance", "mean", dims=(2,))(x)
    # waveforms: dims=(1,) on [B, L]
    w = x[:, 0, :].contiguous()
    np.testing.assert_allclose(TorchScaler("instance", "minmax", dims=(1,))(w).cpu().numpy(), G.torch_scaler(w.cpu(), "instance", "minmax", dims=(1,)).numpy(),
                               rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("tag,T,grid", [("156", 156, (4, 256, 16000)), ("1000", 1000, (1, 320, 32000))])
def test_threshold_sweep_and_event_decoding(golden, tag, T, grid):
    from transformer4sed_b200.src_codec import decoder as D
    g = golden("post.npz")
    strong, weak = decode_inputs(g, tag, T)
    sizes = [int(k) for k in g[f"sizes{tag}"]]
    ths = [0.25, 0.5, 0.75]
    filt = D.median_filter_torch(strong.cuda().transpose(1, 2), sizes)
    ev = D.decode_events(filt, weak.cuda(), ths)
    np.testing.assert_array_equal(ev[:, :3], g[f"events{tag}_idx"])                          # reference order and content
    np.testing.assert_array_equal(frame_to_time(ev[:, 3:].astype(np.float64), *grid), g[f"events{tag}_time"])
    np.testing.assert_array_equal(ev, G.decode_pred_batch_fast(strong, weak, ths, sizes))    # oracle, frame indices
    dfs = D.decode_pred_batch_fast(strong.cuda(), weak.cuda(), [f"clip{i}.flac" for i in range(4)], _Enc(*grid), ths, sizes)
    assert list(dfs) == ths
    n = 0
    for ti, th in enumerate(ths):
        df = dfs[th]
        assert list(df.columns) == ["event_label", "onset", "offset", "filename"]
        sel = g[f"events{tag}_idx"][:, 0] == ti
        assert len(df) == int(sel.sum())
        np.testing.assert_array_equal(df[["onset", "offset"]].to_numpy(), g[f"events{tag}_time"][sel])
        assert list(df["event_label"]) == [f"class{c}" for c in g[f"events{tag}_idx"][sel, 2]]
        assert list(df["filename"]) == [f"clip{b}.wav" for b in g[f"events{tag}_idx"][sel, 1]]
        n += len(df)
    assert n == len(ev)
    # edge cases: nothing above threshold, everything above threshold, a single frame
    zeros = torch.zeros(2, 7, 3, device="cuda")
    assert D.decode_events(zeros, None, [0.5]).shape == (0, 5)
    ones = torch.ones(2, 7, 3, device="cuda")
    full = D.decode_events(ones, None, [0.5, 2.0])
    np.testing.assert_array_equal(full, [[0, b, c, 0, 7] for b in range(2) for c in range(3)])
    one = D.decode_events(torch.tensor([[[0.9], [0.1], [0.9]]], device="cuda"), torch.tensor([[0.6]], device="cuda"), [0.5, 0.7])
    np.testing.assert_array_equal(one, [[0, 0, 0, 0, 1], [0, 0, 0, 2, 3]])                   # weak 0.6 < 0.7 silences threshold 1


@pytest.mark.parametrize("tag,T", [("156", 156), ("1000", 1000)])
def test_scipy_style_score_filters(golden, tag, T):
    from transformer4sed_b200.src_codec import decoder as D
    g = golden("post.npz")
    strong, weak = decode_inputs(g, tag, T)
    sizes = [int(k) for k in g[f"sizes{tag}"]]
    for ft in ("median", "max"):
        raw, post = D.filter_scores(strong.cuda(), sizes, ft, weak.cuda(), need_weak_mask=True)
        np.testing.assert_array_equal(post[:2].cpu().numpy(), g[f"scores{tag}_{ft}"])
        np.testing.assert_array_equal(raw[:2].cpu().numpy(), (strong[:2].transpose(1, 2) * weak[:2].unsqueeze(1)).numpy())
    sr, sp = D.batched_decode_preds(strong.cuda(), ["a/x0.wav", "x1.wav", "x2.wav", "x3.wav"], _Enc(1, 320, 32000), filter=sizes, weak_preds=weak.cuda(),
                                    need_weak_mask=True)
    assert list(sr) == ["x0", "x1", "x2", "x3"] and list(sp["x0"].columns)[:2] == ["onset", "offset"]
    np.testing.assert_array_equal(sp["x1"].to_numpy()[:, 2:].astype(np.float32), g[f"scores{tag}_median"][1])
    # windows longer than the signal and even windows, against scipy itself (the 1-D call the reference makes)
    x = torch.rand(1, 9, 2, generator=torch.Generator().manual_seed(7))
    _, post = D.filter_scores(x.cuda().transpose(1, 2), [17, 4], "median")
    np.testing.assert_array_equal(post[0].cpu().numpy(), G.rank_filter_scores(x[0].numpy(), [17, 4], "median"))
    # beyond 2 L + 1 taps the border needs more than one reflection: scipy's N-D filter reflects periodically (so do we), its 1-D fast path
    # (scipy >= 1.11, the one a 1-D input takes) does not -- no recipe filters a clip shorter than half its window, so the N-D rule is the oracle
    from scipy import ndimage
    _, post = D.filter_scores(x.cuda().transpose(1, 2), [20, 26], "median")
    want = np.stack([ndimage.median_filter(x[0].numpy()[:, c:c + 1], size=(k, 1))[:, 0] for c, k in enumerate([20, 26])], 1)
    np.testing.assert_array_equal(post[0].cpu().numpy(), want)


def test_fused_sed_losses(golden):
    from transformer4sed_b200 import training as TR
    g = golden("post.npz")
    stu, tch, y, yw = loss_inputs()
    stu = [t.cuda().requires_grad_() for t in stu]
    w_weak, w_at, w_cons, w_wc = [float(v) for v in g["loss_weights"]]
    total, parts = TR.sed_losses(*stu, tch[0].cuda(), tch[1].cuda(), y.cuda(), yw.cuda(), (0, 4), (4, 8), w_weak=w_weak, w_at=w_at, w_cons=w_cons,
                                 w_weak_cons=w_wc)
    total.backward()
    np.testing.assert_allclose(total.item(), g["loss_total"], rtol=1e-5)
    np.testing.assert_allclose(parts.cpu().numpy(), g["loss_parts"], rtol=1e-5)
    np.testing.assert_allclose(stu[0].grad[:, ::3, ::25].cpu().numpy(), g["d_strong"], rtol=1e-4, atol=1e-10)
    np.testing.assert_allclose(stu[1].grad.cpu().numpy(), g["d_weak"], rtol=1e-4, atol=1e-10)
    np.testing.assert_allclose(stu[2].grad.cpu().numpy(), g["d_at"], rtol=1e-4, atol=1e-10)
    # deterministic: two runs agree bit for bit
    total2, _ = TR.sed_losses(*[t.detach() for t in stu], tch[0].cuda(), tch[1].cuda(), y.cuda(), yw.cuda(), (0, 4), (4, 8), w_weak=w_weak, w_at=w_at,
                              w_cons=w_cons, w_weak_cons=w_wc)
    assert total2.item() == total.item()
    with pytest.raises(Exception):
        TR.sed_losses(*stu, tch[0].cuda(), tch[1].cuda(), y.cuda(), yw.cuda(), (0, 0), (4, 8))
