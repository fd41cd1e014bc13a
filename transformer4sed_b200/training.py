"""Parameter arena, fused AdamW and the data-parallel gradient exchange (one process per GPU).

Replaces the reference's single-process `nn.DataParallel` + `torch.optim.AdamW` (recipes/desed/setting.py:254-258):
parameters live in one flat fp32 arena (grouped by LR group), gradients are packed into one flat buffer by a libt4s kernel,
all-reduced ONCE over NCCL/NVLink, and a fused AdamW kernel updates the master weights and refreshes the bf16 GEMM-operand
shadow in the same pass.  The path shards by clip with no activation exchange, so the all-reduce is the only collective.
"""
import ctypes

import torch

from . import _lib
from . import functional as F


def _align(n, a=8):  # 16 bytes for the bf16 shadow (TMA base alignment)
    return (n + a - 1) // a * a


def flat_layout(groups):
    """Deterministic arena layout shared by every rank: [(param, element offset)], per-group [start, end) ranges, total size.
    Offsets are 8-element (16-byte in bf16) aligned; frozen parameters are skipped."""
    total, layout, ranges = 0, [], []
    for g in groups:
        start = total
        for p in g["params"]:
            if not p.requires_grad:
                continue
            layout.append((p, total))
            total = _align(total + p.numel())
        ranges.append(dict(name=g["name"], lr=float(g["lr"]), weight_decay=float(g["weight_decay"]), start=start, end=total))
    return layout, ranges, total


def shard_for_rank(n_items, rank, world):
    """Contiguous shard [lo, hi) of a global batch for `rank` (clips are independent: SURVEY §8e)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def _world_size():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def all_reduce_flat(buf):
    """The path's single collective: in-place sum of the packed gradient buffer over ranks; returns the world size."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
        return dist.get_world_size()
    return 1


def bucket_ranges(layout, total, bucket_elems):
    """Cut the arena layout into contiguous buckets of about `bucket_elems` elements: [(first entry, last entry + 1, start, end)]
    (entry indices into `layout`, element offsets into the flat buffers).  Same on every rank: it only depends on the layout."""
    out, i0 = [], 0
    for i, (p, off) in enumerate(layout):
        nxt = layout[i + 1][1] if i + 1 < len(layout) else total
        if nxt - layout[i0][1] >= bucket_elems or i + 1 == len(layout):
            out.append((i0, i + 1, layout[i0][1], nxt))
            i0 = i + 1
    return out


class GradBuckets:
    """Overlap of the gradient exchange with the backward pass: the arena is cut into a few contiguous buckets; a
    post-accumulate-grad hook on every parameter counts the bucket's gradients in, and the moment the last one lands the bucket is
    packed into the flat buffer (`pack(first, last)`: the t4s_grad_pack kernel on the product path) and its slice is all-reduced
    asynchronously (NCCL runs on its own stream while autograd keeps producing the gradients of the earlier layers).  `finish()`
    packs and reduces whatever did not fire (parameters without a gradient this step) and waits for every handle."""

    def __init__(self, layout, total, flat_grad, pack, bucket_elems=1 << 24):
        import torch.distributed as dist
        self.dist = dist
        self.layout, self.flat, self.pack = layout, flat_grad, pack
        self.buckets = bucket_ranges(layout, total, bucket_elems)
        self.bucket_of = {}
        for b, (i0, i1, _, _) in enumerate(self.buckets):
            for i in range(i0, i1):
                self.bucket_of[id(layout[i][0])] = b
        self.pending = [i1 - i0 for i0, i1, _, _ in self.buckets]
        self.handles = [None] * len(self.buckets)
        self.dirty = [False] * len(self.buckets)
        self.hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p, _ in layout]

    def _launch(self, b):
        i0, i1, start, end = self.buckets[b]
        self.pack(i0, i1)
        self.handles[b] = self.dist.all_reduce(self.flat[start:end], async_op=True)

    def _on_grad(self, p):
        b = self.bucket_of[id(p)]
        if self.handles[b] is not None:
            self.dirty[b] = True           # a second backward before the step (gradient accumulation): reduce again in finish()
            return
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def finish(self):
        for b in range(len(self.buckets)):
            if self.handles[b] is None:
                self._launch(b)
            elif self.dirty[b]:
                self.handles[b].wait()
                self._launch(b)
        for h in self.handles:
            h.wait()
        self.pending = [i1 - i0 for i0, i1, _, _ in self.buckets]
        self.handles = [None] * len(self.buckets)
        self.dirty = [False] * len(self.buckets)

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


class ParamArena:
    def __init__(self, module: torch.nn.Module, groups, shadow_bf16=True, betas=(0.9, 0.999), eps=1e-8, overlap=None,
                 bucket_elems=1 << 24):
        """groups: list of dicts {name, params (list of nn.Parameter), lr, weight_decay}.  Parameters not listed are frozen.
        overlap=True (or T4S_OVERLAP_ALLREDUCE=1): bucketed all-reduce launched from autograd hooks while the backward pass is still
        running (`GradBuckets`); default: one all-reduce over the whole buffer after backward, which measured faster on NVSwitch."""
        dev = next(module.parameters()).device
        _lib.ensure_device(next(module.parameters()))
        layout, self.groups, total = flat_layout(groups)
        self.n = total
        self.device = dev
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.shadow = torch.empty(total, dtype=torch.bfloat16, device=dev) if shadow_bf16 else None
        self.layout = layout
        with torch.no_grad():
            for p, off in layout:
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                if self.shadow is not None:
                    p._t4s_shadow = self.shadow[off:off + p.numel()].view(p.shape)
        if self.shadow is not None:
            F.convert(self.flat, self.shadow)
            for p, _ in layout:
                p._t4s_shadow_version = p._version      # functional._arena_shadow re-converts after load_state_dict / copy_
        self.betas, self.eps, self.step_count = betas, eps, 0
        # pointer tables of the pack kernel: the host never rewrites a pinned table whose upload may still be queued (the step has
        # no host sync, so the host can run more than a step ahead of the GPU): two tables, each guarded by the event of its copy
        self._tables = [(torch.empty(len(layout), 3, dtype=torch.int64).pin_memory(),
                         torch.empty(len(layout), 3, dtype=torch.int64, device=dev), torch.cuda.Event()) for _ in range(2)]
        self._table_used = [False, False]
        self._table_turn = 0
        self._layout_off = [off for _, off in layout]
        self._layout_numel = [p.numel() for p, _ in layout]
        self._agreed = self._agreed_local = None
        self.buckets = None
        if overlap is None:
            import os
            # measured on 2 and 8 B200 (64 and 32 clips per GPU): the bucketed exchange is 0.3-1.2 ms per step SLOWER than one
            # all-reduce after backward (NCCL's kernels take SMs from the persistent GEMM / attention CTAs and every bucket costs a
            # pointer-table upload and a launch), so it is opt-in (profiles/r2_multi_gpu.md)
            overlap = _world_size() > 1 and os.environ.get("T4S_OVERLAP_ALLREDUCE", "0") == "1"
        if overlap and _world_size() > 1:
            self.buckets = GradBuckets(layout, total, self.grad, self._pack_range, bucket_elems)
            # one pinned + device pointer table per bucket and parity (see pack_grads for the hazard)
            self._btables = [[(torch.empty(i1 - i0, 3, dtype=torch.int64).pin_memory(), torch.empty(i1 - i0, 3, dtype=torch.int64, device=dev),
                               torch.cuda.Event(), [False]) for _ in range(2)] for i0, i1, _, _ in self.buckets.buckets]
            self._bturn = [0] * len(self.buckets.buckets)

    def _pack_range(self, i0, i1):
        """Pack the gradients of layout entries [i0, i1) (one bucket) into the flat buffer; called from autograd hooks."""
        b = next(k for k, r in enumerate(self.buckets.buckets) if r[0] == i0)
        turn = self._bturn[b]
        self._bturn[b] ^= 1
        t, t_dev, copied, used = self._btables[b][turn]
        if used[0]:
            copied.synchronize()
        for k in range(i0, i1):
            p, off = self.layout[k]
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                g = g.float().contiguous()
                p.grad = g
            t[k - i0, 0] = g.data_ptr() if g is not None else 0
            t[k - i0, 1] = off
            t[k - i0, 2] = p.numel()
        with torch.cuda.device(self.device):
            t_dev.copy_(t, non_blocking=True)
            copied.record()
            used[0] = True
            _lib.check(_lib.load().t4s_grad_pack(_lib.ptr(t_dev), i1 - i0, _lib.ptr(self.grad), _lib.stream_ptr()), "t4s_grad_pack")

    def pack_grads(self):
        """Gather every parameter's .grad into the flat buffer (missing grads -> zeros) with one kernel."""
        turn = self._table_turn
        self._table_turn ^= 1
        t, t_dev, copied = self._tables[turn]
        if self._table_used[turn]:
            copied.synchronize()            # the upload that last read this pinned table has finished
        # one vectorised store per column through the numpy view of the pinned table (an element-wise `t[i, j] = ...` costs a
        # torch dispatch each: 3 ms of host time per step for the ~700 tensors of the PMAM / DASM models)
        ptrs = []
        for p, _ in self.layout:
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                g = g.float().contiguous()
                p.grad = g
            ptrs.append(g.data_ptr() if g is not None else 0)
        tn = t.numpy()
        tn[:, 0] = ptrs
        tn[:, 1] = self._layout_off
        tn[:, 2] = self._layout_numel
        with torch.cuda.device(self.device):
            t_dev.copy_(t, non_blocking=True)
            copied.record()
            self._table_used[turn] = True
            _lib.check(_lib.load().t4s_grad_pack(_lib.ptr(t_dev), len(self.layout), _lib.ptr(self.grad), _lib.stream_ptr()),
                       "t4s_grad_pack")

    def all_reduce(self):
        """Sum of the packed gradients over ranks (NCCL over NVLink / NVSwitch)."""
        return all_reduce_flat(self.grad)

    def step(self, lr_scale=1.0):
        """pack -> all-reduce -> fused AdamW.  Differences from torch.optim.AdamW (documented): bias correction uses one
        global step count (torch keeps one per tensor: identical unless a tensor first receives a gradient late), and under data
        parallelism a tensor counts as "without gradient" only when it has none on every rank.  Per-rank mean losses are averaged
        with weight 1/world: shard batches so that every rank holds the same number of strong / weak / unlabelled clips
        (`shard_for_rank` on each subset) if the reference's global-batch means are to be reproduced exactly."""
        # torch.optim.AdamW leaves a parameter whose grad is None completely untouched (no decay, no moment update); the fused
        # kernel runs over whole flat ranges, so such (rare, e.g. the mask token under the upstream no-op masking) tensors are
        # put back after the update
        has_grad = [p.grad is not None for p, _ in self.layout]
        if _world_size() > 1:
            # every rank must skip the same tensors or the replicas drift apart: a tensor is skipped only if NO rank has a gradient.
            # The set is agreed on the first step and re-checked every 100 steps (the agreement needs a device -> host read, which
            # would otherwise stall the launch queue every step); in between a rank whose own set changed fails loudly.
            import torch.distributed as dist
            if self._agreed is None or self.step_count % 100 == 0:
                flags = torch.tensor(has_grad, dtype=torch.int32, device=self.device)
                dist.all_reduce(flags, op=dist.ReduceOp.MAX)
                self._agreed, self._agreed_local = flags.bool().tolist(), has_grad
            elif has_grad != self._agreed_local:
                raise _lib.T4sError("the set of parameters without a gradient changed on this rank between two agreement points; "
                                    "under data parallelism every rank must run the same graph (ParamArena.step)")
            has_grad = self._agreed
        skipped = [(off, p.numel()) for (p, off), h in zip(self.layout, has_grad) if not h]
        keep = [(off, n, self.flat[off:off + n].clone(), self.exp_avg[off:off + n].clone(), self.exp_avg_sq[off:off + n].clone())
                for off, n in skipped]
        if self.buckets is not None:
            self.buckets.finish()          # every bucket was packed and all-reduced from the autograd hooks (stragglers: here)
            world = _world_size()
        else:
            self.pack_grads()
            world = self.all_reduce()
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device(self.device):
            for g in self.groups:
                n = g["end"] - g["start"]
                if n == 0:
                    continue
                o4, o2 = g["start"] * 4, g["start"] * 2
                _lib.check(lib.t4s_adamw_step(
                    ctypes.c_void_p(self.flat.data_ptr() + o4), ctypes.c_void_p(self.grad.data_ptr() + o4),
                    ctypes.c_void_p(self.exp_avg.data_ptr() + o4), ctypes.c_void_p(self.exp_avg_sq.data_ptr() + o4),
                    ctypes.c_void_p(self.shadow.data_ptr() + o2) if self.shadow is not None else ctypes.c_void_p(0), n,
                    g["lr"] * lr_scale, self.betas[0], self.betas[1], self.eps, g["weight_decay"], self.step_count, 1.0 / world,
                    _lib.stream_ptr()), "t4s_adamw_step")
        for off, n, w, m1, m2 in keep:
            self.flat[off:off + n].copy_(w)
            self.exp_avg[off:off + n].copy_(m1)
            self.exp_avg_sq[off:off + n].copy_(m2)
            if self.shadow is not None:
                F.convert(w, self.shadow[off:off + n])
        for p, _ in self.layout:
            p.grad = None


class MeanTeacher:
    """EMA teacher of the mean-teacher recipes (reference src/utils/scheduler.py:125-130, recipes/desed/finetune/train.py:129-213):
    teacher = alpha * teacher + (1 - alpha) * student with alpha = min(1 - 1/step, ema_factor), over EVERY parameter.

    The teacher is a deep copy of the student whose arena-managed parameters are re-homed into one flat fp32 buffer with the
    student arena's layout (plus a bf16 shadow for its GEMMs), so the whole update is ONE kernel launch over that buffer; the few
    parameters outside the arena (frozen tensors, unused heads) are updated tensor by tensor.  Buffers (e.g. BatchNorm running
    statistics) are not averaged, as upstream."""

    def __init__(self, student: torch.nn.Module, arena: ParamArena):
        import copy
        # deep-copy with the arena views detached from the flat buffer (a view would drag the whole arena into every copy)
        saved = [(p, p.data, getattr(p, "_t4s_shadow", None)) for p, _ in arena.layout]
        for p, off in arena.layout:
            p.data = p.data.clone()
            if hasattr(p, "_t4s_shadow"):
                del p._t4s_shadow
        try:
            self.teacher = copy.deepcopy(student)
        finally:
            for p, data, shadow in saved:
                p.data = data
                if shadow is not None:
                    p._t4s_shadow = shadow          # (`_t4s_shadow_version` stays on the parameter; `.data =` does not bump the version)
        for p in self.teacher.parameters():
            p.requires_grad_(False)
        self.arena = arena
        dev = arena.device
        names = {id(p): n for n, p in student.named_parameters()}
        tparams = dict(self.teacher.named_parameters())
        self.flat = torch.zeros(arena.n, dtype=torch.float32, device=dev)
        self.shadow = torch.empty(arena.n, dtype=torch.bfloat16, device=dev) if arena.shadow is not None else None
        managed = set()
        with torch.no_grad():
            for p, off in arena.layout:
                tp = tparams[names[id(p)]]
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(tp.data)
                tp.data = view
                if self.shadow is not None:
                    tp._t4s_shadow = self.shadow[off:off + p.numel()].view(p.shape)
                managed.add(names[id(p)])
        if self.shadow is not None:
            F.convert(self.flat, self.shadow)
            for p, _ in arena.layout:
                tp = tparams[names[id(p)]]
                tp._t4s_shadow_version = tp._version
        sparams = dict(student.named_parameters())
        self.rest = [(tparams[n], sparams[n]) for n in tparams if n not in managed]

    def update(self, step: int, ema_factor: float = 0.999):
        alpha = min(1.0 - 1.0 / max(step, 1), ema_factor)
        lib = _lib.load()
        with torch.cuda.device(self.arena.device):
            _lib.check(lib.t4s_ema_update(_lib.ptr(self.flat), _lib.ptr(self.arena.flat), _lib.ptr(self.shadow), self.arena.n, alpha,
                                          _lib.stream_ptr()), "t4s_ema_update")
            for tp, sp in self.rest:
                if tp.dtype == torch.float32 and tp.is_contiguous() and sp.is_contiguous():
                    _lib.check(lib.t4s_ema_update(_lib.ptr(tp.data), _lib.ptr(sp.data), ctypes.c_void_p(0), tp.numel(), alpha, _lib.stream_ptr()),
                               "t4s_ema_update")
                    F.invalidate_weight_cache(tp)
        return alpha


def check_tensor_name_decoder(tensor_name) -> bool:
    """reference recipes/desed/finetune/passt/setting.py:18-25."""
    return any(kw in tensor_name for kw in ("decoder", "f_pool_module", "transformer_projector"))


def passt_param_groups(net, lr_dict):
    """The optimizer groups of the DESED fine-tuning recipes, as reference recipes/desed/finetune/passt/setting.py:28-103 (`get_params`)
    builds them from ``configs["opt"]["param_groups"]`` -- including its side effects on ``requires_grad``:

      * encoder: every backbone parameter; with ``step_lr`` > 0 the blocks ``12 - idx <= step_lr`` and every key containing "norm."
        (that is also the LayerNorms of the LOWER blocks: upstream's `elif "norm." in k`) train at 2 x lr, the rest at lr;
        ``lr <= 0`` freezes everything but "norm." keys; ``freeze_layer`` > 0 freezes the blocks below it;
      * decoder: names containing "decoder" / "f_pool_module" / "transformer_projector" (frozen when lr <= 0);
      * head: everything else.

    Returns ParamArena groups (dicts with name / params / lr / weight_decay).  Parameters that can never receive a gradient on
    the SED path (the PaSST classification heads) are left out, which is what torch.optim.AdamW does with `grad is None`."""
    import re
    enc, dec = lr_dict["encoder"], lr_dict["decoder"]
    head = lr_dict.get("head", dec)
    backbone = [(k, p) for k, p in net.backbone.named_parameters()]
    passt_ids = {id(p) for _, p in backbone}
    unused = lambda k: k.startswith("head.") or k.startswith("head_dist.")   # noqa: E731
    if not enc.get("step_lr"):
        groups = [dict(name="encoder", params=[p for k, p in backbone if not unused(k)], lr=enc["lr"], weight_decay=enc["weight_decay"])]
    else:
        low, high = [], []
        for k, p in backbone:
            if unused(k):
                continue
            m = re.search(r"blocks.(\d+)", k)
            if m and (12 - int(m.group(1)) <= enc["step_lr"]):
                high.append(p)
            elif "norm." in k:
                high.append(p)
            else:
                low.append(p)
        groups = [dict(name="encoder_low", params=low, lr=enc["lr"], weight_decay=enc["weight_decay"]),
                  dict(name="encoder_high", params=high, lr=enc["lr"] * 2, weight_decay=enc["weight_decay"])]
    if enc["lr"] <= 0:
        for k, p in backbone:
            if "norm." not in k:
                p.requires_grad = False
    if enc.get("freeze_layer", 0) > 0:
        for k, p in backbone:
            m = re.search(r"blocks.(\d+)", k)
            p.requires_grad = bool((m and int(m.group(1)) + 1 > enc["freeze_layer"]) or "norm." in k)
    decoder_params = [p for k, p in net.named_parameters() if check_tensor_name_decoder(k)]
    decoder_ids = {id(p) for p in decoder_params}
    if dec["lr"] <= 0:
        for p in decoder_params:
            p.requires_grad = False
    head_params = [p for p in net.parameters() if id(p) not in passt_ids and id(p) not in decoder_ids]
    groups.append(dict(name="decoder", params=decoder_params, lr=dec["lr"], weight_decay=dec["weight_decay"]))
    groups.append(dict(name="head", params=head_params, lr=head["lr"], weight_decay=head["weight_decay"]))
    return groups


class _SedLosses(torch.autograd.Function):
    """The six losses of the mean-teacher step and their weighted total in one pass (csrc/post.cu `t4s_sed_losses_*`)."""

    @staticmethod
    def forward(ctx, strong, weak, at, t_strong, t_at, y, yw, rows, weights):
        _lib.ensure_device(strong)
        dev = strong.device
        f = lambda t: t.detach().contiguous().float()  # noqa: E731
        strong_c, weak_c, at_c, ts_c, ta_c, y_c, yw_c = map(f, (strong, weak, at, t_strong, t_at, y, yw))
        d = _lib.SedLosses()
        d.strong, d.weak, d.at, d.t_strong, d.t_at, d.y, d.yw = (t.data_ptr() for t in (strong_c, weak_c, at_c, ts_c, ta_c, y_c, yw_c))
        d.batch, d.classes = weak_c.shape[0], weak_c.shape[1]
        d.strong_inner = strong_c[0].numel()
        d.s0, d.s1, d.w0, d.w1 = rows
        d.w_weak, d.w_at, d.w_cons, d.w_weak_cons = weights
        with torch.cuda.device(dev):
            ws = torch.empty(768, dtype=torch.float32, device=dev)
            out = torch.empty(7, dtype=torch.float32, device=dev)
            _lib.check(_lib.load().t4s_sed_losses_fwd(ctypes.byref(d), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr()), "t4s_sed_losses_fwd")
        ctx.desc, ctx.keep = d, (strong_c, weak_c, at_c, ts_c, ta_c, y_c, yw_c)
        ctx.shapes = (strong.shape, weak.shape, at.shape)
        parts = out[:6].clone()
        ctx.mark_non_differentiable(parts)
        return out[6], parts

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        strong_c, weak_c, at_c = ctx.keep[:3]
        dev = strong_c.device
        with torch.cuda.device(dev):
            g = g_total.detach().reshape(1).float().contiguous()
            ds, dw, da = torch.empty_like(strong_c), torch.empty_like(weak_c), torch.empty_like(at_c)
            _lib.check(_lib.load().t4s_sed_losses_bwd(ctypes.byref(ctx.desc), _lib.ptr(g), _lib.ptr(ds), _lib.ptr(dw), _lib.ptr(da), _lib.stream_ptr()),
                       "t4s_sed_losses_bwd")
        return ds.view(ctx.shapes[0]), dw.view(ctx.shapes[1]), da.view(ctx.shapes[2]), None, None, None, None, None, None


def sed_losses(stu_strong, stu_weak, stu_at, tch_strong, tch_at, labels, labels_weak, strong_rows, weak_rows, w_weak=1.0, w_at=1.0, w_cons=1.0,
               w_weak_cons=1.0):
    """Reference recipes/desed/finetune/train.py:166-188 as ONE forward and ONE backward kernel pair:

        total = BCE(strong[S], y[S]) + w_weak BCE(weak[W], yw[W]) + w_at BCE(at[W], yw[W])
                + w_cons (MSE(strong, tch_strong) + w_weak_cons MSE(weak, tch_at) + w_at MSE(at, tch_at))

    S = rows [strong_rows[0], strong_rows[1]), W = rows [weak_rows[0], weak_rows[1]) (the contiguous masks of `get_mask`).  Returns
    (total, parts) with parts = the six un-weighted losses in the order (class_strong, class_weak, class_at, cons_strong, cons_weak,
    cons_at).  Gradients flow to the three student tensors only (the teacher side is detached upstream)."""
    return _SedLosses.apply(stu_strong, stu_weak, stu_at, tch_strong, tch_at, labels, labels_weak, (*strong_rows, *weak_rows),
                            (float(w_weak), float(w_at), float(w_cons), float(w_weak_cons)))


def mean_teacher_step(student, teacher: MeanTeacher, arena: ParamArena, stu_feat, tch_feat, labels, labels_weak, strong_rows, weak_rows,
                      stu_kwargs=None, tch_kwargs=None, loss_weights=None, step_num=2, ema_factor=0.999, lr_scale=1.0):
    """One mean-teacher fine-tuning step (reference recipes/desed/finetune/train.py:129-199): student forward, teacher forward without
    gradient (its `train_tch_kwargs`, e.g. the sliding-window fusion of finetune2.yaml:72-78), the fused six-loss kernel, backward,
    packed all-reduce + fused AdamW, and the one-kernel EMA update.  `step_num` is the reference's `scheduler.step_num` AFTER
    `scheduler.step()` (2 on the first update).  Returns (total, parts)."""
    stu = student(stu_feat, **(stu_kwargs or {}))
    with torch.no_grad():
        tch = teacher.teacher(tch_feat, **(tch_kwargs or {}))
    total, parts = sed_losses(stu[0], stu[1], stu[2]["at_out"], tch[0], tch[2]["at_out"], labels, labels_weak, strong_rows, weak_rows,
                              **(loss_weights or {}))
    total.backward()
    arena.step(lr_scale)
    teacher.update(step_num, ema_factor)
    return total.detach(), parts
