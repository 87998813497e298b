"""Data-parallel plumbing of the hot path: the path shards by image, one process per
GPU, and the only exchange is the all-reduce of the head's weight gradients.

Mirrors what the reference gets from detectron2: images split evenly over ranks
(detectron2/data/build.py:268-275 asserts ``IMS_PER_BATCH % world_size == 0``; the training sampler
strides an index stream by rank, detectron2/data/samplers/distributed_sampler.py:40-45) and DDP
averaging gradients over ranks (detectron2/engine/defaults.py:280-283).  Instead of DDP's per-bucket
hooks the head's gradients live in ONE flat fp32 buffer (21.4 MB for the RepPoints head) that the
weight-gradient kernels can accumulate into directly and a single all-reduce (NCCL over NVLink on the
GPU box, gloo in the CPU tests) averages in place.  torch.distributed is plumbing only: no compute
happens here.
"""
import torch
import torch.distributed as dist


def shard_range(num_images, rank, world_size):
    """Contiguous slice [start, stop) of a global batch owned by `rank`; the batch must divide evenly
    (detectron2/data/build.py:268-275)."""
    if num_images % world_size != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (num_images, world_size))
    per = num_images // world_size
    return rank * per, (rank + 1) * per


def rank_strided_indices(num_indices, rank, world_size):
    """Indices rank, rank+world, ... of an index stream (InfiniteSampler, distributed_sampler.py:40-45)."""
    return list(range(rank, num_indices, world_size))


class GradBucket:
    """Flat fp32 gradient buffer with named, shaped views.

    ``views[name]`` is a tensor view into ``flat``; kernels that accumulate (sdb_dcn_backward_weight
    adds into grad_weight) can be pointed at the view directly, so no pack copy is needed.
    ``pad_to`` lets a benchmark size the bucket like the whole head (the DCN weights are part of it).
    """

    def __init__(self, shapes, device, pad_to=0):
        self.shapes = dict(shapes)
        total = sum(int(torch.Size(s).numel()) for s in self.shapes.values())
        self.numel = max(total, int(pad_to))
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.views, o = {}, 0
        for name, shp in self.shapes.items():
            n = int(torch.Size(shp).numel())
            self.views[name] = self.flat[o:o + n].view(shp)
            o += n
        self.used = o

    def zero_(self):
        self.flat.zero_()

    def pack(self, grads):
        """Copy ``grads[name]`` (any float dtype) into the bucket."""
        for name, g in grads.items():
            self.views[name].copy_(g)

    def all_reduce(self, group=None, average=True, async_op=False, prescaled=False):
        """Sum over ranks in place (one collective for the whole head) with DDP's averaging semantics.

        ``average``: the result is the mean over ranks.  The division is applied BEFORE the collective (the buffer
        is multiplied by 1/world, then summed), so it also holds for ``async_op=True``: waiting on the returned
        handle leaves the averaged gradients in ``flat``.  ``prescaled=True`` says the producer already folded
        1/world into what it accumulated (``sdb_dcn_backward`` takes a ``scale``), so no extra pass is made.
        No-op (returns None) without an initialised process group or with a single rank."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if average and not prescaled:
            self.flat.mul_(1.0 / dist.get_world_size(group))
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def unpack(self, params):
        """Write the bucket's views into ``params[name].grad`` (cast to the parameter dtype)."""
        for name, p in params.items():
            g = self.views[name]
            p.grad = g.to(p.dtype).clone() if p.grad is None else p.grad.copy_(g)
