"""world_size-2 gloo test of the shard / gather / reduce host logic (the N>1 path of bench.py and
fastdiffsr_b200.parallel), on CPU tensors."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastdiffsr_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.randn(n_items, 3, 4, 4, generator=g)
        local, n_valid = P.shard_batch(full, rank, world)
        start, stop, per = P.shard_bounds(n_items, rank, world)
        assert local.shape[0] == per and n_valid == stop - start
        processed = local * 2 + 1   # stand-in for the per-shard sampling loop (no cross-sample op)
        out = P.gather_batch(processed, n_items)
        ok = torch.equal(out, full * 2 + 1)
        acc = torch.tensor([float(n_valid), float(local[:n_valid].double().sum())], dtype=torch.float64)
        P.reduce_sums(acc)
        ok = ok and acc[0].item() == n_items and abs(acc[1].item() - full.double().sum().item()) < 1e-9
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run(n_items, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_even_shards():
    _run(8)


def test_ragged_shards_are_padded_and_dropped():
    _run(5)


def test_single_item_leaves_a_rank_empty():
    _run(1)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 16, 64):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                a, b, per = P.shard_bounds(n, r, world)
                assert 0 <= b - a <= per
                seen += list(range(a, b))
            assert seen == list(range(n))


def test_psnr_from_sse():
    sse = torch.tensor([0.0, 3.0 * 16], dtype=torch.float64)
    ps = P.psnr_from_sse(sse, 3 * 16)
    assert torch.isinf(ps[0]) and abs(ps[1].item() - 10 * torch.log10(torch.tensor(255.0 ** 2)).item()) < 1e-4
