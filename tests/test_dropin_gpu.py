"""GPU: the drop-in promises around the module surface that the recipes rely on beyond forward():
  * forward hooks on ``decoder.encoder_blocks[i]`` see the reference's (T, B, C) layout (recipes/desed/pmam/extractor_feature.py:83-89);
  * the mirror survives ``copy.deepcopy`` (EMA teachers: src/utils/scheduler.py) and single-device ``nn.DataParallel`` wrapping
    (src/utils/__init__.py:11-21, every main.py) although its parameters carry arena views and cached bf16 operand copies, and
    refuses multi-device replication loudly (the path is one process per GPU);
  * `load_reference_checkpoint` restores a DataParallel-prefixed reference state dict with the stage filters of the recipes."""
import copy

import pytest
import torch

from transformer4sed_b200.utils import synth

pytestmark = pytest.mark.gpu

KW = dict(embed_dim=192, decoder_dim=192, decoder="transformerXL", decoder_layer_num=2, at_adapter=True, f_pool="mean_pool", mlm=False)


def _net(seed=3):
    from transformer4sed_b200.src_models.passt.passt_sed import PaSST_SED
    net = PaSST_SED(load_pretrained_model=False, **KW)
    net.load_state_dict(synth.synth_state_dict_like(net, seed))
    return net.cuda().eval()


def _mel(net, B=2):
    ext = net.get_feature_extractor().eval()
    return ext.logmel(synth.synth_wav(B, 320000, seed=5).cuda())


@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_decoder_block_hooks_see_time_major_layout(mode):
    from transformer4sed_b200 import functional as F
    F.set_precision(mode)
    try:
        net = _net()
        mel = _mel(net)
        seen = []
        h = net.decoder.encoder_blocks[1].register_forward_hook(lambda m, i, o: seen.append(o))
        with torch.no_grad():
            strong, weak, other = net(mel)
        h.remove()
        (out,) = seen
        assert out.shape == (1000, 2, 192)                                           # (T, B, C) as upstream
        rows = out.detach().transpose(0, 1).reshape(-1, out.shape[-1])               # the upstream consumer's flattening
        assert rows.shape == (2000, 192)
        assert torch.equal(rows.view(2, 1000, 192)[1], out[:, 1])                    # clip-major rows, no row permutation
    finally:
        F.set_precision("bf16")


def test_deepcopy_dataparallel_and_replicate_keep_working():
    from transformer4sed_b200 import functional as F
    from transformer4sed_b200.training import ParamArena, passt_param_groups
    F.set_precision("bf16")
    net = _net().train()
    mel = _mel(net)
    opt = dict(encoder=dict(lr=1e-5, weight_decay=1e-4, freeze_layer=0, step_lr=0), decoder=dict(lr=1e-4, weight_decay=1e-4),
               head=dict(lr=1e-4, weight_decay=1e-4))
    arena = ParamArena(net, passt_param_groups(net, opt))        # parameters become arena views with bf16 shadows
    with torch.no_grad():
        ref = net(mel)[0].float()
    twin = copy.deepcopy(net)
    assert all(p.data_ptr() != q.data_ptr() for p, q in zip(net.parameters(), twin.parameters()))
    with torch.no_grad():
        assert torch.equal(twin(mel)[0].float(), ref)
        twin.classifier.weight.mul_(2.0)                          # the copy is independent of the arena ...
        assert torch.equal(net(mel)[0].float(), ref)              # ... and the original does not see its update
        assert not torch.equal(twin(mel)[0].float(), ref)         # while the copy's cached bf16 operand follows its own weights
    dp = torch.nn.DataParallel(net, device_ids=[0])
    with torch.no_grad():
        assert torch.equal(dp(mel)[0].float(), ref)
    from transformer4sed_b200._lib import T4sError
    with pytest.raises(T4sError):                                 # what DataParallel does with more than one device: refused loudly
        torch.nn.parallel.replicate(net, [0, 0])
    # a training step through the arena, then every wrapper sees the updated weights
    s, w, o = net(mel)
    (s.float().mean() + w.float().mean()).backward()
    arena.step()
    with torch.no_grad():
        new = net(mel)[0].float()
        assert not torch.equal(new, ref)
        assert torch.equal(dp(mel)[0].float(), new)


def test_load_reference_checkpoint_into_mirror(tmp_path):
    from transformer4sed_b200.utils.checkpoint import load_reference_checkpoint
    src = _net(seed=9)
    sd = {f"module.{k}": v.detach().cpu() for k, v in src.state_dict().items()}      # as saved from an nn.DataParallel-wrapped reference net
    path = tmp_path / "best_student.pt"
    torch.save(sd, path)
    dst = _net(seed=3)
    res = load_reference_checkpoint(dst, str(path))
    assert not res.missing_keys and not res.unexpected_keys
    mel = _mel(src)
    with torch.no_grad():
        assert torch.equal(dst(mel)[0], src(mel)[0])
    # MLM -> fine-tuning hand-off: the classifier and the second AT adapter layer keep their fresh initialisation
    dst2 = _net(seed=3)
    before = dst2.classifier.weight.detach().clone()
    res = load_reference_checkpoint(torch.nn.DataParallel(dst2, device_ids=[0]), sd, stage="finetune_from_mlm")
    assert any("classifier." in k for k in res.missing_keys)
    assert torch.equal(dst2.classifier.weight, before)
    assert torch.equal(dst2.backbone.blocks[0].attn.qkv.weight, src.backbone.blocks[0].attn.qkv.weight)
