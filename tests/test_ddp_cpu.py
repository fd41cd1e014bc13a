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
