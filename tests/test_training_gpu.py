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


def test_mean_teacher_matches_reference_ema():
    """MeanTeacher.update == the reference's update_ema loop (src/utils/scheduler.py:125-130) on every parameter, arena-managed or not,
    and the teacher's bf16 GEMM shadow follows."""
    import copy
    from transformer4sed_b200.training import MeanTeacher, ParamArena
    torch.manual_seed(1)
    student = torch.nn.Sequential(torch.nn.Linear(24, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 8)).cuda()
    student[2].bias.requires_grad_(False)                       # a frozen tensor: outside the arena, still averaged
    ref_teacher = copy.deepcopy(student)
    arena = ParamArena(student, [dict(name="all", params=list(student.parameters()), lr=1e-2, weight_decay=0.0)], shadow_bf16=True)
    mt = MeanTeacher(student, arena)
    assert all(not p.requires_grad for p in mt.teacher.parameters())
    assert student[0].weight.data_ptr() != mt.teacher[0].weight.data_ptr()
    for step in range(1, 5):
        student(torch.randn(4, 24, device="cuda")).square().mean().backward()
        arena.step()
        with torch.no_grad():
            student[2].bias.add_(0.1)
        alpha = mt.update(step, ema_factor=0.999)
        assert abs(alpha - min(1 - 1 / step, 0.999)) < 1e-12
        with torch.no_grad():
            for tp, sp in zip(ref_teacher.parameters(), student.parameters()):
                tp.data.mul_(alpha).add_(sp.data, alpha=1 - alpha)
        for (n, a), (_, b) in zip(mt.teacher.named_parameters(), ref_teacher.named_parameters()):
            assert (a - b).abs().max().item() < 1e-6, (step, n)
        w = mt.teacher[0].weight
        assert (w._t4s_shadow.float() - w).abs().max().item() <= w.abs().max().item() * 2 ** -8


def test_param_without_gradient_is_left_untouched_like_torch_adamw():
    """A trainable parameter that received no gradient this step (e.g. the mask token under the upstream no-op masking) is skipped by
    torch.optim.AdamW: no weight decay, no moment update.  The arena's flat-range kernel must leave it (and its bf16 shadow) alone."""
    from transformer4sed_b200.training import ParamArena
    torch.manual_seed(2)
    m1 = torch.nn.ModuleDict(dict(a=torch.nn.Linear(16, 16), unused=torch.nn.Linear(16, 4))).cuda()
    import copy
    m2 = copy.deepcopy(m1)
    arena = ParamArena(m1, [dict(name="all", params=list(m1.parameters()), lr=1e-2, weight_decay=0.1)], shadow_bf16=True)
    opt = torch.optim.AdamW(m2.parameters(), lr=1e-2, weight_decay=0.1)
    w0 = m1["unused"].weight.detach().clone()
    for _ in range(3):
        x = torch.randn(8, 16, device="cuda")
        for m in (m1, m2):
            m["a"](x).square().mean().backward()
        arena.step()
        opt.step()
        opt.zero_grad()
    assert torch.equal(m1["unused"].weight, w0) and torch.equal(m2["unused"].weight, w0)
    assert (m1["unused"].weight._t4s_shadow.float() - w0).abs().max().item() <= w0.abs().max().item() * 2 ** -8
    assert (m1["a"].weight - m2["a"].weight).abs().max().item() < 2e-6
