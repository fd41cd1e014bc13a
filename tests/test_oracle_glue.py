"""CPU: oracle restatement of the train-loop glue (frame_shift, mixup, class-wise median filter) vs golden vectors recorded from the
unmodified reference (oracle/make_golden.py glue)."""
import numpy as np
import torch

from conftest import checksum
from oracle import glue as G
from transformer4sed_b200.utils import synth


def _inputs(g):
    B = 6
    mel = synth.synth_tensor(31, "glue_mel", (B, 128, 1000))
    np.testing.assert_allclose(checksum(mel), g["mel_ck"], rtol=1e-12)
    label = (synth.synth_tensor(31, "glue_label", (B, 10, 1000)) > 0.6).float()
    label4 = (synth.synth_tensor(31, "glue_label4", (B, 10, 250)) > 0.6).float()
    probs = torch.sigmoid(2.0 * synth.synth_tensor(31, "glue_probs", (4, 1000, 10)))
    np.testing.assert_allclose(checksum(probs), g["probs_ck"], rtol=1e-12)
    return mel, label, label4, probs


def test_glue_oracle_matches_reference(golden):
    g = golden("glue.npz")
    mel, label, label4, probs = _inputs(g)
    f, lab = G.frame_shift(mel, label, 1, [int(s) for s in g["shifts"]])
    np.testing.assert_array_equal(f[:, ::8, ::5].numpy(), g["fs_mel"])
    np.testing.assert_array_equal(lab.numpy(), g["fs_label"])
    f, lab = G.frame_shift(mel, label4, 4, [int(s) for s in g["shifts4"]])
    np.testing.assert_array_equal(f[:, ::8, ::5].numpy(), g["fs4_mel"])
    np.testing.assert_array_equal(lab.numpy(), g["fs4_label"])
    mf, ml = G.mixup(mel, label, torch.from_numpy(g["perm"]), float(g["c"]))
    np.testing.assert_allclose(mf[:, ::8, ::5].numpy(), g["mx_mel"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ml.numpy(), g["mx_label"], rtol=1e-6, atol=1e-7)
    mf, ml = G.mixup(mel, label, torch.from_numpy(g["perm_h"]), float(g["c_h"]), "hard")
    np.testing.assert_allclose(mf[:, ::8, ::5].numpy(), g["mh_mel"], rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(ml.numpy(), g["mh_label"])
    np.testing.assert_array_equal(G.median_filter(probs, [int(k) for k in g["med_sizes"]]).numpy(), g["med"])


def test_freq_nonlinear_and_filt_aug_oracle_and_table(golden):
    """Oracle vs the reference, and the host-side (source bin, weight) table the CUDA kernel consumes vs np.interp itself."""
    import random
    from transformer4sed_b200.src_preprocess.data_aug import freq_warp_table
    g = golden("glue.npz")
    mel, _, _, _ = _inputs(g)
    small = mel[:3, :, :200].contiguous()
    phase = float(g["fn_phase"])
    np.testing.assert_array_equal(G.freq_nonlinear(small, phase, bias=0.03 * 0.7).numpy(), g["fn"])
    for F_, f_, bias_, ph in ((128, 1, 0.021, phase), (128, 1, 0.03, 0.93), (64, 1, 0.0, 0.5), (128, 1, 0.03, 0.25)):
        j, w = freq_warp_table(F_, f_, bias_, ph)
        rows = np.random.default_rng(0).standard_normal((5, F_))
        ind = np.arange(F_)
        x = ind / F_
        ind_t = F_ * (x + bias_ * np.sin(2 * np.pi * (f_ * x + ph)))
        for r in rows:
            ours = r[j] + w * (r[np.minimum(j + 1, F_ - 1)] - r[j])
            np.testing.assert_allclose(ours, np.interp(ind, ind_t, r), rtol=1e-12, atol=1e-12)
    random.seed(11)
    assert random.random() == phase
