"""Host logic of the multi-GPU path on CPU: view partitioning and the packed single all-reduce, with a
world_size-2 gloo process group (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from geosplatting_b200.parallel import (FlatAllReduce, GradientBucket, allreduce_gradients, common_flat, shard_views,
                                        view_counts)


def test_shard_views_partitions_the_batch():
    views = list(range(13))
    for world in (1, 2, 4, 8):
        parts = [shard_views(views, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == views
        assert [len(p) for p in parts] == view_counts(13, world)
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        shard_views(views, 2, 2)


def test_bucket_pack_unpack_roundtrip_single_process():
    shapes = [(5, 3), (5, 4), (7,), (2, 2, 2)]
    ts = [torch.randn(s) for s in shapes]
    ts[2] = None
    out = allreduce_gradients(ts, GradientBucket(shapes, "cpu"))
    for a, b in zip(ts, out):
        assert torch.equal(b, torch.zeros_like(b) if a is None else a)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 1000
        # per-view "gradients" are a deterministic function of the view id: every rank sums its own views
        def view_grad(v):
            g = torch.Generator().manual_seed(v)
            return [torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g), torch.randn(6, 8, 8, 3, generator=g),
                    torch.randn(1, generator=g)]
        mine = shard_views(list(range(n_views)), rank, world)
        local = [sum(x) for x in zip(*[view_grad(v) for v in mine])]
        bucket = GradientBucket([t.shape for t in local], "cpu")
        summed = allreduce_gradients(local, bucket, average_over=n_views)
        expect = [sum(x) / n_views for x in zip(*[view_grad(v) for v in range(n_views)])]
        ok = all(torch.allclose(a, b, atol=1e-5) for a, b in zip(summed, expect))
        q.put((rank, ok, bucket.nbytes))
    finally:
        dist.destroy_process_group()


def test_view_sharded_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == (1000 * 3 + 1000 * 4 + 6 * 8 * 8 * 3 + 1) * 4


def test_common_flat_covers_views_of_one_buffer():
    buf = torch.arange(100, dtype=torch.float32)
    a, b, c = buf[4:16].view(4, 3), buf[16:24].view(2, 4), buf[30:31]
    flat = common_flat([a, None, b, c])
    assert flat is not None and flat.data_ptr() == buf[4:].data_ptr() and flat.numel() == 27
    flat.mul_(2.0)
    assert float(a[0, 0]) == 8.0 and float(c[0]) == 60.0 and float(buf[3]) == 3.0 and float(buf[31]) == 31.0
    assert common_flat([a, torch.zeros(3)]) is None                      # different buffers
    assert common_flat([buf[0:10:2]]) is None                            # not contiguous
    FlatAllReduce([a, b]).wait()                                         # no process group: a no-op, in place
    packed = FlatAllReduce([a, torch.ones(3)])
    assert packed.packed
    packed.wait()


def _flat_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, T, n = 500, 96, 3
        # the layout of the batched backward: [env 4T | quats 4N | ks 2N | means 3N | ... | exposure n | spare]
        def local(r):
            g = torch.Generator().manual_seed(100 + r)
            return torch.randn(4 * T + 19 * N + n + 1, generator=g)
        flat = local(rank).clone()
        views = [flat[:4 * T].view(T, 4), flat[4 * T:4 * T + 4 * N].view(N, 4), flat[-1:]]
        red = FlatAllReduce(views + [flat[4 * T + 4 * N:-1]], async_op=True)
        assert not red.packed and red.nbytes == flat.numel() * 4
        red.wait()
        q.put((rank, bool(torch.allclose(flat, local(0) + local(1), atol=1e-6))))
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_in_place_world2_gloo():
    """The N > 1 exchange as bench.py issues it: ONE all-reduce on the backward's own flat buffer, no pack."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
