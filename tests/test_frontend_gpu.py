"""GPU parity: K1 fused front end (through the C ABI) vs the golden vectors of the unmodified reference and vs the
CPU oracle.  Tolerances: the log-mel is compared in the normalised-log domain the model consumes; fp32 evaluation
noise of that quantity is ~5e-5 (tests/test_oracle_frontend.py::test_fp64_oracle_brackets_fp32), contract is <=1e-3."""
import numpy as np
import pytest
import torch

from conftest import checksum
from oracle import frontend as OF
from transformer4sed_b200.utils import synth

pytestmark = pytest.mark.gpu
ATOL_LOG = 1e-3


def _ext(**kw):
    from transformer4sed_b200.src_models.passt.passt_feature_extraction import PasstFeatureExtractor
    args = dict(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, htk=False, fmin=0.0, fmax=None,
                wav_norm=True, fmin_aug_range=10, fmax_aug_range=2000)
    args.update(kw)
    return PasstFeatureExtractor(**args).cuda().eval()


def test_golden_2s_and_10s(golden):
    g = golden("frontend_passt.npz")
    ext = _ext()
    wav_a = synth.synth_wav(2, 64000, seed=11).cuda()
    power = ext(wav_a)
    assert power.shape == (2, 128, 200)
    ref = torch.from_numpy(g["a_power"]).cuda()
    rel = ((power - ref).abs().max() / ref.abs().max()).item()
    assert rel < 1e-5, rel
    lm = ext.normalize(power)
    assert (lm.cpu() - torch.from_numpy(g["a_logmel"])).abs().max().item() < ATOL_LOG
    fused = ext.logmel(wav_a)
    assert (fused - lm).abs().max().item() < 1e-6
    wav_b = synth.synth_wav(1, 320000, seed=12).cuda()
    out = ext.logmel(wav_b)
    assert out.shape == (1, 128, 1000)
    err = (out.cpu() - torch.from_numpy(g["b_logmel"])).abs()
    assert err.max().item() < ATOL_LOG, err.max().item()
    assert err.mean().item() < 2e-5, err.mean().item()


def test_golden_ragged_and_degenerate(golden):
    g = golden("frontend_passt.npz")
    ext = _ext()
    for n in (1025, 1345, 3201, 32001):
        w = synth.synth_wav(1, n, seed=100 + n).cuda()
        out = ext.logmel(w)
        assert out.shape == (1, 128, 1 + (n - 1) // 320)
        assert (out.cpu() - torch.from_numpy(g[f"c{n}_logmel"])).abs().max().item() < ATOL_LOG, n
    w = torch.zeros(3, 16000)
    w[1] += 0.25
    w[2, 5000] = 1.0
    out = ext.logmel(w.cuda())
    assert torch.isfinite(out).all()
    assert (out.cpu() - torch.from_numpy(g["d_logmel"])).abs().max().item() < ATOL_LOG


def test_vs_oracle_random_batches():
    ext = _ext()
    for seed, (b, n) in enumerate([(3, 48000), (5, 320000), (1, 515), (2, 100001)]):
        w = synth.synth_wav(b, n, seed=40 + seed)
        ref = OF.passt_logmel(w)
        out = ext.logmel(w.cuda()).cpu()
        assert out.shape == ref.shape
        assert (out - ref).abs().max().item() < ATOL_LOG, (b, n)


def test_train_mode_band_jitter_matches_reference_draws():
    ext = _ext().train()
    w = synth.synth_wav(2, 32000, seed=7)
    torch.manual_seed(123)
    out = ext.logmel(w.cuda()).cpu()
    torch.manual_seed(123)
    fmin = 0.0 + torch.randint(10, (1,)).item()
    fmax = 15000 + 1000 - torch.randint(2000, (1,)).item()
    ref = OF.passt_logmel(w, fmin=float(fmin), fmax=float(fmax))
    assert (out - ref).abs().max().item() < ATOL_LOG


def test_bf16_output_and_size_independent_properties():
    ext = _ext(wav_norm=True)
    w = synth.synth_wav(4, 320000, seed=9).cuda()
    a = ext.logmel(w)
    # peak normalisation makes the front end scale invariant: a size-independent property at full shape
    b = ext.logmel(3.7 * w)
    assert (a - b).abs().max().item() < 2e-4
    # batch independence / determinism
    c = ext.logmel(w[1:3])
    assert torch.equal(c, a[1:3])
    h = ext.logmel(w, out_dtype=torch.bfloat16)
    assert h.dtype == torch.bfloat16
    assert (h.float() - a).abs().max().item() < 2e-2


def test_rejects_cpu_tensor_and_bad_shapes():
    from transformer4sed_b200 import _lib
    ext = _ext()
    with pytest.raises(_lib.T4sError):
        ext.logmel(torch.zeros(1, 32000))
    with pytest.raises(_lib.T4sError):
        ext.logmel(torch.zeros(1, 400).cuda())  # too short for reflect padding


def test_dcase16k_front_end(golden):
    """SURVEY §8 a1': the 16 kHz DCASE-style parametrisation (n_fft = win = 2048, hop 256, hamming, magnitude, HTK mel, dB clamp) on the
    generic-n_fft kernel vs torchaudio golden vectors of the reference's `setmelspectrogram` + `take_log`, fused and un-fused."""
    from oracle import frontend as OF
    from transformer4sed_b200.src_preprocess.feats_extraction import setmelspectrogram, take_log
    g = golden("frontend_dcase16k.npz")
    w = synth.synth_wav(2, 48000, seed=21)
    np.testing.assert_allclose(checksum(w), g["in_ck"], rtol=1e-12)
    ms = setmelspectrogram(dict(sample_rate=16000, n_window=2048, hop_length=256, f_min=0, f_max=8000, n_mels=128)).cuda()
    mel = ms(w.cuda())
    db_fused = ms.logmel(w.cuda())
    db = take_log(mel)
    assert db.shape == tuple(g["db"].shape) == (2, 128, 1 + 48000 // 256)
    assert (db.cpu() - torch.from_numpy(g["db"])).abs().max().item() < 2e-3           # dB units (fp32 FFT noise near the -50 dB floor)
    assert (db_fused - db).abs().max().item() < 1e-5
    # full 10 s clips at 16 kHz and a ragged length, against the CPU oracle
    for n in (160000, 40001):
        w = synth.synth_wav(3, n, seed=30 + n % 7)
        ours = ms.logmel(w.cuda()).cpu()
        ref = OF.dcase_logmel(w)
        assert ours.shape == ref.shape and (ours - ref).abs().max().item() < 2e-3
    # another power-of-two frame (n_fft 512) goes through the same kernel
    ms2 = setmelspectrogram(dict(sample_rate=16000, n_window=512, hop_length=160, f_min=0, f_max=8000, n_mels=64)).cuda()
    w = synth.synth_wav(2, 32000, seed=33)
    ref = OF.dcase_logmel(w, n_fft=512, hop=160, n_mels=64)
    assert (ms2.logmel(w.cuda()).cpu() - ref).abs().max().item() < 2e-3
