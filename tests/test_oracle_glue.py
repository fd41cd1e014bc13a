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


def test_feature_transformation_host_logic(golden, monkeypatch):
    """The mirror's `feature_transformation` / `filt_aug` / `freq_nonlinear` host side (RNG draw order, filter assembly, warp table)
    against the reference's stack on the shipped settings.  The two kernel launches are replaced by CPU doubles that compute what
    the kernels are specified to compute (csrc/aug.cu: gather-lerp, row bias) -- the kernels themselves are checked in
    tests/test_glue_gpu.py."""
    import random
    from transformer4sed_b200.src_preprocess import data_aug as A

    def warp_double(mel, j, w):
        jt = torch.from_numpy(j)
        wt = torch.from_numpy(w.astype(np.float32)).view(1, -1, 1)
        a = mel[:, jt]
        return torch.where(wt == 0, a, a + wt * (mel[:, torch.clamp(jt + 1, max=mel.shape[1] - 1)] - a))

    def bias_double(features, bias_host):
        return features + bias_host.view(features.shape[0], features.shape[1], 1)

    monkeypatch.setattr(A, "_launch_freq_warp", warp_double)
    monkeypatch.setattr(A, "_launch_add_rowbias", bias_double)
    g = golden("glue.npz")
    mel, _, _, _ = _inputs(g)
    small = mel[:3, :, :200].contiguous()
    keep = small.clone()
    random.seed(31)
    torch.manual_seed(31)
    a, b = A.feature_transformation(small, n_transform=2, choice=[1, 0, 0, 1], filter_db_range=[-26, 26], filter_bands=[2, 5],
                                    filter_minimum_bandwidth=4, filter_type="step", log=True, norm_std=5.0)
    np.testing.assert_allclose(a.numpy(), g["ft_a"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(b.numpy(), g["ft_b"], rtol=0, atol=2e-5)
    assert torch.equal(small, keep)                                   # the input batch is not modified
    one = A.feature_transformation(small, n_transform=1, choice=[0, 0, 0, 0], filter_db_range=[-26, 26], filter_bands=[2, 5],
                                   filter_minimum_bandwidth=4, filter_type="step", log=True)
    assert torch.equal(one, small) and one.data_ptr() != small.data_ptr()
    import pytest
    with pytest.raises(NotImplementedError):
        A.feature_transformation(small, 1, [0, 1, 0, 0], [-26, 26], [2, 5], 4, "step")
    # single transforms against their own golden entries (same doubles)
    random.seed(11)
    np.testing.assert_allclose(A.freq_nonlinear(small, bias=0.03 * 0.7).numpy(), g["fn"], rtol=0, atol=2e-5)
    torch.manual_seed(21)
    np.testing.assert_allclose(A.filt_aug(small, db_range=[-6, 6], n_band=[3, 6], min_bw=6, filter_type="step", log=True, norm_std=5.0).numpy(),
                               g["fa_step"], rtol=0, atol=1e-6)
