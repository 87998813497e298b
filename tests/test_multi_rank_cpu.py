"""world_size-2 `gloo` tests (CPU) of the N > 1 host logic: image sharding and the head-gradient
bucket all-reduce.  No kernel of the hot path runs here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slenderobjdet_b200.dist import GradBucket, rank_strided_indices, shard_range

SHAPES = {"cls_dcn.weight": (8, 4, 3, 3), "refine_dcn.weight": (8, 4, 3, 3), "cls_out.bias": (5,)}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)  # identical on both ranks
        base = {k: torch.randn(*s, generator=g) for k, s in SHAPES.items()}
        grads = {k: v * (rank + 1) for k, v in base.items()}           # rank r contributes (r+1) * base
        bucket = GradBucket(SHAPES, "cpu", pad_to=1000)
        assert bucket.numel == 1000 and bucket.used == 8 * 4 * 9 * 2 + 5
        bucket.pack(grads)
        # kernels accumulate into the views directly: emulate a second accumulation
        bucket.views["cls_out.bias"].add_(1.0)
        bucket.all_reduce(average=True)
        scale = sum(r + 1 for r in range(world)) / world
        for k in SHAPES:
            exp = base[k] * scale + (1.0 if k == "cls_out.bias" else 0.0)
            assert torch.allclose(bucket.views[k], exp, atol=1e-6), k
        assert float(bucket.flat[bucket.used:].abs().sum()) == 0.0
        # async + average: waiting on the handle leaves the MEAN in the buffer (the division is not skipped)
        b2 = GradBucket(SHAPES, "cpu")
        b2.pack(grads)
        b2.all_reduce(average=True, async_op=True).wait()
        assert torch.allclose(b2.views["cls_dcn.weight"], base["cls_dcn.weight"] * scale, atol=1e-6)
        # producer already scaled by 1/world (sdb_dcn_backward's `scale`): no second division
        b3 = GradBucket(SHAPES, "cpu")
        b3.pack({k: v / world for k, v in grads.items()})
        b3.all_reduce(average=True, prescaled=True)
        assert torch.allclose(b3.views["cls_dcn.weight"], base["cls_dcn.weight"] * scale, atol=1e-6)
        params = {k: torch.nn.Parameter(torch.zeros(*s, dtype=torch.bfloat16)) for k, s in SHAPES.items()}
        bucket.unpack(params)
        assert all(p.grad.dtype == torch.bfloat16 for p in params.values())
        assert torch.allclose(params["cls_dcn.weight"].grad.float(), base["cls_dcn.weight"] * scale, atol=2e-2)
        # sharding: every image exactly once, by contiguous range and by stride
        lo, hi = shard_range(16, rank, world)
        owned = torch.zeros(16)
        owned[lo:hi] = 1
        dist.all_reduce(owned)
        assert bool((owned == 1).all())
        strided = torch.zeros(17)
        strided[rank_strided_indices(17, rank, world)] = 1
        dist.all_reduce(strided)
        assert bool((strided == 1).all())
        out.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_bucket_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


def test_shard_range_rejects_uneven_batches():
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
    assert shard_range(16, 3, 8) == (6, 8)


def test_bucket_without_process_group_is_local():
    b = GradBucket({"w": (2, 3)}, "cpu")
    b.views["w"].fill_(2.0)
    assert b.all_reduce() is None and float(b.flat.sum()) == 12.0
