"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path — identical arena layout on every rank, batch
sharding by rank, ONE all-reduce over the packed gradient buffer, 1/world scaling — reproduces the single-process gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.GELU(), torch.nn.LayerNorm(16), torch.nn.Linear(16, 3))


def _groups(m):
    return [dict(name="enc", params=list(m[0].parameters()) + list(m[2].parameters()), lr=1e-3, weight_decay=0.0),
            dict(name="head", params=list(m[3].parameters()), lr=1e-2, weight_decay=0.0)]


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(8, 12, generator=g), torch.randn(8, 3, generator=g)


def _worker(rank, world, port, out):
    from transformer4sed_b200.training import all_reduce_flat, flat_layout, shard_for_rank
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _model()
        layout, ranges, total = flat_layout(_groups(m))
        x, y = _data()
        lo, hi = shard_for_rank(x.shape[0], rank, world)
        torch.nn.functional.mse_loss(m(x[lo:hi]), y[lo:hi]).backward()
        flat = torch.zeros(total)
        for p, off in layout:                      # (the product packs with the t4s_grad_pack kernel; same layout contract)
            flat[off:off + p.numel()] = p.grad.flatten()
        w = all_reduce_flat(flat)
        flat /= w
        if rank == 0:
            torch.save(dict(flat=flat, offsets=[o for _, o in layout], ranges=ranges, total=total, shard=(lo, hi)), out)
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_exchange_matches_single_process(tmp_path):
    from transformer4sed_b200.training import flat_layout
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    m = _model()
    layout, ranges, total = flat_layout(_groups(m))
    assert total == r["total"] and [o for _, o in layout] == r["offsets"] and all(o % 8 == 0 for o in r["offsets"])
    assert r["shard"] == (0, 4)
    x, y = _data()
    torch.nn.functional.mse_loss(m(x), y).backward()
    for p, off in layout:
        assert torch.allclose(r["flat"][off:off + p.numel()], p.grad.flatten(), atol=1e-6)
    assert ranges[0]["start"] == 0 and ranges[0]["end"] == ranges[1]["start"] and ranges[1]["end"] == total


def test_shards_cover_batch_without_overlap():
    from transformer4sed_b200.training import shard_for_rank
    for n in (1, 7, 64, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_for_rank(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_sliding_window_placement_matches_reference_loop():
    """Host logic of the batched sliding window (no GPU): window enumeration, grouping by patch grid and output offsets equal the
    reference's Python loop (encoder_slide_window.py:26-34) for the shipped train / validation parameters."""
    from oracle import model as OM
    from transformer4sed_b200.src_models.encoder_slide_window import EncoderSlideWindow

    class Probe(EncoderSlideWindow):
        def encode(self, x):
            raise AssertionError("not used")

        def frames_per_window(self, width):
            return min((width - 16) // 10 + 1, 99)

    for win in ([512, 49], [512, 31], [512, 29], [1000, 100], [256, 256]):
        ours = Probe(None, win).window_starts(1000)
        ref = [(a, b - a) for a, b in OM.window_starts(1000, tuple(win))]
        assert ours == ref, (win, ours, ref)
    w31 = Probe(None, [512, 31]).window_starts(1000)
    assert len(w31) == 17 and w31[-1] == (496, 504) and Probe(None, [512, 31]).frames_per_window(504) == 49
    w49 = Probe(None, [512, 49]).window_starts(1000)
    assert len(w49) == 11 and w49[-1] == (490, 510) and Probe(None, [512, 49]).frames_per_window(510) == 50
    assert Probe(None, [2000, 10]).window_starts(1000) == []      # window longer than the clip + step: no window, embedding is all zeros


def test_pmam_dasm_schemas_and_lora_merge_roundtrip_cpu():
    """State-dict schemas of the PMAM / DASM mirrors match the golden key lists recorded from the reference; LoRA merge / un-merge on
    eval() / train() is exact bookkeeping (host-side torch ops, no kernel)."""
    import numpy as np
    import os
    import torch
    from transformer4sed_b200 import schema
    from transformer4sed_b200.src_models import lora
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for name, shapes in (("pmam_base.npz", schema.passt_cnn_shapes()), ("dasm_base.npz", schema.dasm_shapes())):
        keys = [str(k) for k in np.load(os.path.join(gdir, name))["sd_keys"] if not str(k).endswith("num_batches_tracked")]
        assert sorted(shapes) == keys
    lin = lora.Linear(32, 48, r=8, lora_alpha=1)
    with torch.no_grad():
        lin.lora_B.normal_(0, 0.3)
    w0 = lin.weight.detach().clone()
    lin.eval()
    assert lin.merged and torch.allclose(lin.weight, w0 + (lin.lora_B @ lin.lora_A) / 8, atol=1e-7)
    lin.eval()
    assert torch.allclose(lin.weight, w0 + (lin.lora_B @ lin.lora_A) / 8, atol=1e-7)     # idempotent
    lin.train()
    assert not lin.merged and torch.allclose(lin.weight, w0, atol=1e-6)
    assert not lin.weight.requires_grad and lin.lora_A.requires_grad


def test_passt_param_groups_match_reference_get_params():
    """`training.passt_param_groups` builds the same optimizer groups (membership, lr, weight decay) and leaves the same requires_grad
    flags as the reference's `get_params` for the shipped finetune2 settings and three variants (golden: oracle/make_golden.py groups).
    The PaSST classification heads (never on the SED path: no gradient, skipped by torch's AdamW) are the only keys left out."""
    import json
    import os
    from transformer4sed_b200.src_models.passt.passt_sed import PaSST_SED
    from transformer4sed_b200.training import passt_param_groups
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "param_groups.json")))
    base = dict(passt_feature_layer=10, f_pool="mean_pool", decode_ratio=10, at_adapter=True, decoder="transformerXL", decoder_layer_num=3,
                decoder_pos_emd_len=1000, mlm=False)
    unused = lambda n: n.startswith("backbone.head.") or n.startswith("backbone.head_dist.")   # noqa: E731
    for tag, g in gold.items():
        net = PaSST_SED(load_pretrained_model=False, **base)
        names = {id(p): n for n, p in net.named_parameters()}
        groups = passt_param_groups(net, g["lr_dict"])
        assert len(groups) == len(g["groups"]), tag
        for ours, ref in zip(groups, g["groups"]):
            assert sorted(names[id(p)] for p in ours["params"]) == [n for n in ref["names"] if not unused(n)], (tag, ours["name"])
            assert abs(ours["lr"] - ref["lr"]) < 1e-15 and abs(ours["weight_decay"] - ref["weight_decay"]) < 1e-15, (tag, ours["name"])
        assert sorted(n for n, p in net.named_parameters() if p.requires_grad) == g["trainable"], tag


def test_dropin_install_registers_reference_import_paths():
    """`dropin.install()` makes the reference's import statements (recipes/**/setting.py, main.py) resolve to the CUDA mirrors; run in
    a subprocess so the registration does not leak into the other tests."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import transformer4sed_b200.dropin as d; d.install()\n"
        "from src.models.passt.passt_sed import PaSST_SED\n"
        "from src.models.cnn_transformer.passt_cnn import PaSST_CNN\n"
        "from src.models.detect_any_sound.detect_any_sound import DASM\n"
        "from src.models.passt.passt_win import PasstWithSlide\n"
        "from src.postprocess.filter import median_filter_torch\n"
        "import src.models.lora as lora\n"
        "from src.preprocess.data_aug import mixup, frame_shift, feature_transformation\n"
        "assert feature_transformation.__module__.startswith('transformer4sed_b200.')\n"
        "mods = {PaSST_SED.__module__, PaSST_CNN.__module__, DASM.__module__, PasstWithSlide.__module__, lora.Linear.__module__, median_filter_torch.__module__}\n"
        "assert all(m.startswith('transformer4sed_b200.') for m in mods), mods\n"
        "print('ok')\n") % __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_arena_shadow_follows_torch_side_parameter_updates(monkeypatch):
    """Host logic of `functional._arena_shadow`: the bf16 shadow of an arena parameter is re-converted when the parameter was
    modified through torch (load_state_dict / copy_ bump the version counter) and left alone otherwise.  `convert` is replaced by
    a CPU double; the kernel itself is covered by tests/test_training_gpu.py."""
    import torch
    from transformer4sed_b200 import functional as F
    calls = []

    def convert_double(src, dst):
        calls.append(src.data_ptr())
        dst.copy_(src.reshape(dst.shape))
        return dst

    monkeypatch.setattr(F, "convert", convert_double)
    monkeypatch.setattr(F, "_MODE", "bf16")
    flat, shadow = torch.zeros(16), torch.zeros(16, dtype=torch.bfloat16)
    lin = torch.nn.Linear(4, 3, bias=False)
    p = lin.weight
    with torch.no_grad():                                     # what ParamArena.__init__ does with every managed parameter
        v = flat[2:14].view(3, 4)
        v.copy_(p.data)
        p.data = v
    p._t4s_shadow = shadow[2:14].view(3, 4)
    shadow.copy_(flat)
    p._t4s_shadow_version = p._version
    assert F.cast_weight(p) is p._t4s_shadow and calls == []
    flat.mul_(2.0)                                            # kernel-style update of master (+ shadow): no version bump, no refresh
    shadow.copy_(flat)
    assert F.cast_weight(p) is p._t4s_shadow and calls == []
    lin.load_state_dict({"weight": torch.full((3, 4), 0.5)})  # checkpoint resume after the arena was built
    assert torch.equal(flat[2:14], torch.full((12,), 0.5))    # written through the view into the flat master ...
    out = F.cast_weight(p)
    assert len(calls) == 1 and torch.equal(out.float(), torch.full((3, 4), 0.5))   # ... and the shadow caught up exactly once
    assert F.cast_weight(p) is out and len(calls) == 1
    with torch.no_grad():
        p.copy_(torch.full((3, 4), 0.25))
    assert torch.equal(F.to_plain(p, torch.bfloat16).float(), torch.full((3, 4), 0.25)) and len(calls) == 2
    q = torch.nn.Parameter(torch.ones(2, 2))                  # a shadow without version bookkeeping keeps the old behaviour
    q._t4s_shadow = torch.zeros(2, 2, dtype=torch.bfloat16)
    assert F.cast_weight(q) is q._t4s_shadow and len(calls) == 2


def _bucket_worker(rank, world, port, out):
    """GradBuckets with a torch copy standing in for the pack kernel: hooks fire during backward, buckets are reduced asynchronously."""
    from transformer4sed_b200.training import GradBuckets, flat_layout, shard_for_rank
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _model()
        extra = torch.nn.Parameter(torch.zeros(5))           # a parameter that never receives a gradient (e.g. an unused mask token)
        groups = _groups(m)
        groups[1]["params"].append(extra)
        layout, ranges, total = flat_layout(groups)
        flat = torch.zeros(total)
        launched = []

        def pack(i0, i1):
            launched.append((i0, i1))
            for p, off in layout[i0:i1]:
                flat[off:off + p.numel()] = p.grad.flatten() if p.grad is not None else 0.0

        gb = GradBuckets(layout, total, flat, pack, bucket_elems=64)
        x, y = _data()
        lo, hi = shard_for_rank(x.shape[0], rank, world)
        for step in range(2):                                # two steps: the counters must re-arm
            for p, _ in layout:
                p.grad = None
            launched.clear()
            torch.nn.functional.mse_loss(m(x[lo:hi]), y[lo:hi]).backward()
            in_backward = list(launched)
            gb.finish()
        if rank == 0:
            torch.save(dict(flat=flat / world, buckets=gb.buckets, in_backward=in_backward, all=list(launched)), out)
        gb.remove()
    finally:
        dist.destroy_process_group()


def test_bucketed_overlapped_all_reduce_matches_single_process(tmp_path):
    from transformer4sed_b200.training import bucket_ranges, flat_layout
    out = str(tmp_path / "r0.pt")
    mp.spawn(_bucket_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    m = _model()
    layout, ranges, total = flat_layout(_groups(m))
    x, y = _data()
    torch.nn.functional.mse_loss(m(x), y).backward()
    for p, off in layout:
        assert torch.allclose(r["flat"][off:off + p.numel()], p.grad.flatten(), atol=1e-6)
    b = r["buckets"]
    assert len(b) >= 3 and b[0][0] == 0 and all(x[1] == y[0] and x[3] == y[2] for x, y in zip(b, b[1:]))
    assert len(r["in_backward"]) >= len(b) - 1          # every bucket but the one holding the gradient-less parameter fired from a hook
    assert sorted(r["all"]) == sorted((x[0], x[1]) for x in b)
    # bucket boundaries only depend on the layout
    assert bucket_ranges(layout, total, 64)[:2] == [tuple(x) for x in b[:2]]
