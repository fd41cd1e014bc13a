"""CPU: oracle front end vs the golden vectors produced by the unmodified reference."""
import numpy as np
import torch

from conftest import checksum
from oracle import frontend as F
from transformer4sed_b200.utils import synth


def test_passt_logmel_bit_close(golden):
    g = golden("frontend_passt.npz")
    wav_a = synth.synth_wav(2, 64000, seed=11)
    np.testing.assert_allclose(checksum(wav_a), g["a_in_ck"], rtol=1e-12)
    np.testing.assert_allclose(F.passt_power_mel(wav_a).numpy(), g["a_power"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(F.passt_logmel(wav_a).numpy(), g["a_logmel"], rtol=0, atol=1e-6)
    wav_b = synth.synth_wav(1, 320000, seed=12)
    np.testing.assert_allclose(checksum(wav_b), g["b_in_ck"], rtol=1e-12)
    out = F.passt_logmel(wav_b)
    assert out.shape == (1, 128, 1000)
    np.testing.assert_allclose(out.numpy(), g["b_logmel"], rtol=0, atol=1e-6)


def test_passt_logmel_ragged_and_degenerate(golden):
    g = golden("frontend_passt.npz")
    for n in (1025, 1345, 3201, 32001):
        w = synth.synth_wav(1, n, seed=100 + n)
        np.testing.assert_allclose(checksum(w), g[f"c{n}_in_ck"], rtol=1e-12)
        out = F.passt_logmel(w)
        assert out.shape[-1] == 1 + (n - 1) // 320
        np.testing.assert_allclose(out.numpy(), g[f"c{n}_logmel"], rtol=0, atol=1e-6)
    w = torch.zeros(3, 16000)
    w[1] += 0.25
    w[2, 5000] = 1.0
    np.testing.assert_allclose(F.passt_logmel(w).numpy(), g["d_logmel"], rtol=0, atol=1e-6)


def test_mel_basis_jitter(golden):
    g = golden("frontend_passt.npz")
    mb = F.kaldi_mel_banks(128, 1024, 32000, 3.0, 15583.0)
    np.testing.assert_allclose(mb.sum(1).numpy(), g["jit_basis_rowsum"], rtol=1e-6)
    assert ((mb > 0).float().argmax(1).numpy() == g["jit_basis_first_nz"]).all()
    assert int((mb > 0).sum(1).max()) <= 32  # kernel limit: 32 taps per mel row


def test_dcase16k(golden):
    g = golden("frontend_dcase16k.npz")
    w = synth.synth_wav(2, 48000, seed=21)
    np.testing.assert_allclose(checksum(w), g["in_ck"], rtol=1e-12)
    np.testing.assert_allclose(F.dcase_logmel(w).numpy(), g["db"], rtol=0, atol=2e-4)


def test_fp64_oracle_brackets_fp32():
    """fp32 evaluation noise of the log-mel (what a CUDA fp32 FFT may legitimately differ by)."""
    w = synth.synth_wav(1, 64000, seed=5)
    d = (F.passt_logmel(w, dtype=torch.float64) - F.passt_logmel(w)).abs().max().item()
    assert d < 5e-4
