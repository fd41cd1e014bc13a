"""GPU: frame_shift / mixup / class-wise median filter kernels through the C ABI against golden vectors of the
unmodified reference, with the reference's host RNG replayed (python `random`, `torch.randperm`, `np.random.beta`)."""
import random

import numpy as np
import pytest
import torch

from transformer4sed_b200.utils import synth

pytestmark = pytest.mark.gpu


def test_glue_kernels_match_reference(golden):
    from oracle import glue as G
    from transformer4sed_b200.src_postprocess.filter import median_filter_torch
    from transformer4sed_b200.src_preprocess.data_aug import frame_shift, mixup
    g = golden("glue.npz")
    B = 6
    mel = synth.synth_tensor(31, "glue_mel", (B, 128, 1000)).cuda()
    label = (synth.synth_tensor(31, "glue_label", (B, 10, 1000)) > 0.6).float().cuda()
    label4 = (synth.synth_tensor(31, "glue_label4", (B, 10, 250)) > 0.6).float().cuda()
    probs = torch.sigmoid(2.0 * synth.synth_tensor(31, "glue_probs", (4, 1000, 10))).cuda()
    random.seed(123)
    f, lab = frame_shift(mel, label, net_pooling=1)
    assert np.array_equal(f[:, ::8, ::5].cpu().numpy(), g["fs_mel"]) and np.array_equal(lab.cpu().numpy(), g["fs_label"])
    random.seed(7)
    f, lab = frame_shift(mel, label4, net_pooling=4)
    assert np.array_equal(f[:, ::8, ::5].cpu().numpy(), g["fs4_mel"]) and np.array_equal(lab.cpu().numpy(), g["fs4_label"])
    random.seed(7)
    assert np.array_equal(frame_shift(mel)[:, ::8, ::5].cpu().numpy(), g["fs4_mel"])      # feature-only form draws the same shifts
    torch.manual_seed(5)
    np.random.seed(5)
    mf, ml = mixup(mel, label, c=np.random.beta(10, 0.5))
    np.testing.assert_allclose(mf[:, ::8, ::5].cpu().numpy(), g["mx_mel"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ml.cpu().numpy(), g["mx_label"], rtol=1e-6, atol=1e-6)
    torch.manual_seed(6)
    np.random.seed(6)
    mf, ml = mixup(mel, label, mixup_label_type="hard")
    np.testing.assert_allclose(mf[:, ::8, ::5].cpu().numpy(), g["mh_mel"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(ml.cpu().numpy(), g["mh_label"])
    sizes = [int(k) for k in g["med_sizes"]]
    assert np.array_equal(median_filter_torch(probs, sizes).cpu().numpy(), g["med"])          # bit-exact: medians are selected, not computed
    # more than 10 classes: upstream leaves classes >= 10 at zero; strict_upstream=False filters all of them
    p12 = torch.sigmoid(torch.randn(2, 300, 12, generator=torch.Generator().manual_seed(1))).cuda()
    s12 = [5, 8, 3, 21, 1, 9, 33, 7, 11, 15, 9, 4]
    out = median_filter_torch(p12, s12)
    assert np.array_equal(out.cpu().numpy(), G.median_filter(p12.cpu(), s12).numpy()) and out[:, :, 10:].abs().max().item() == 0.0
    assert np.array_equal(median_filter_torch(p12, s12, strict_upstream=False).cpu().numpy(), G.median_filter(p12.cpu(), s12, 12).numpy())
    with pytest.raises(IndexError):
        median_filter_torch(p12[:, :, :4].contiguous(), s12[:4])


def test_feature_transforms_match_reference(golden):
    """freq_nonlinear (upstream: np.interp per (clip, frame) on the host) and filt_aug on log-mel features against golden vectors of
    the unmodified reference, the host RNG replayed draw for draw."""
    from transformer4sed_b200.src_preprocess.data_aug import filt_aug, freq_nonlinear
    g = golden("glue.npz")
    small = synth.synth_tensor(31, "glue_mel", (6, 128, 1000))[:3, :, :200].contiguous().cuda()
    random.seed(11)
    fn = freq_nonlinear(small, bias=0.03 * 0.7)
    np.testing.assert_allclose(fn.cpu().numpy(), g["fn"], rtol=1e-5, atol=2e-6)      # fp32 lerp vs numpy's float64 interp
    torch.manual_seed(21)
    fa = filt_aug(small, db_range=[-6, 6], n_band=[3, 6], min_bw=6, filter_type="step", log=True, norm_std=5.0)
    np.testing.assert_allclose(fa.cpu().numpy(), g["fa_step"], rtol=1e-6, atol=1e-6)
    torch.manual_seed(22)
    fa = filt_aug(small, db_range=[-6, 6], n_band=[3, 6], min_bw=6, filter_type="linear", log=True, norm_std=5.0)
    np.testing.assert_allclose(fa.cpu().numpy(), g["fa_lin"], rtol=1e-6, atol=1e-6)
    with pytest.raises(NotImplementedError):
        filt_aug(small, log=False)
    # the stack the recipes call (two independently augmented copies, shipped DESED settings): draw order as upstream
    from transformer4sed_b200.src_preprocess.data_aug import feature_transformation
    random.seed(31)
    torch.manual_seed(31)
    a, b = feature_transformation(small, n_transform=2, choice=[1, 0, 0, 1], filter_db_range=[-26, 26], filter_bands=[2, 5],
                                  filter_minimum_bandwidth=4, filter_type="step", log=True, norm_std=5.0)
    np.testing.assert_allclose(a.cpu().numpy(), g["ft_a"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(b.cpu().numpy(), g["ft_b"], rtol=1e-5, atol=2e-5)
