"""GPU: fused AdamW / gradient pack / EMA kernels vs torch, and __graft_entry__.smoke()."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_param_arena_adamw_matches_torch():
    from transformer4sed_b200.training import ParamArena
    torch.manual_seed(0)
    m1 = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.LayerNorm(17), torch.nn.Linear(17, 5)).cuda()
    m2 = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.LayerNorm(17), torch.nn.Linear(17, 5)).cuda()
    m2.load_state_dict(m1.state_dict())
    g1 = [dict(name="a", params=list(m1[0].parameters()) + list(m1[1].parameters()), lr=1e-2, weight_decay=1e-2),
          dict(name="b", params=list(m1[2].parameters()), lr=3e-3, weight_decay=0.0)]
    arena = ParamArena(m1, g1, shadow_bf16=True)
    opt = torch.optim.AdamW([dict(params=list(m2[0].parameters()) + list(m2[1].parameters()), lr=1e-2, weight_decay=1e-2),
                             dict(params=list(m2[2].parameters()), lr=3e-3, weight_decay=0.0)], betas=(0.9, 0.999), eps=1e-8)
    for it in range(4):
        x = torch.randn(8, 33, device="cuda")
        for m in (m1, m2):
            m(x).square().mean().backward()
        arena.step()
        opt.step()
        opt.zero_grad()
        for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            assert (a - b).abs().max().item() < 2e-6, (it, n)
            assert a.grad is None
            assert (a._t4s_shadow.float() - a).abs().max().item() <= a.abs().max().item() * 2 ** -8


def test_ema_update():
    import ctypes
    from transformer4sed_b200 import _lib
    t, s = torch.randn(1000, device="cuda"), torch.randn(1000, device="cuda")
    ref = 0.99 * t + 0.01 * s
    _lib.check(_lib.load().t4s_ema_update(_lib.ptr(t), _lib.ptr(s), ctypes.c_void_p(0), 1000, 0.99, _lib.stream_ptr()), "ema")
    assert (t - ref).abs().max().item() < 1e-6
