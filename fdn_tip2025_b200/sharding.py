"""Multi-GPU inference = image-level sharding, one process per GPU, no collective on the data path.

The global FFTs forbid spatial tiling, so each GPU runs whole frames.  Image i of the global batch
goes to rank i mod world - the rule the reference's validation loop uses (image_restoration_model.py:731).  The only
collective is the max-reduction of the measured time in bench.py, outside the timed region.
"""
import torch
import torch.distributed as dist


def shard_indices(rank, world, images_per_gpu):
    """Global image indices processed by `rank` in one step (weak scaling: images_per_gpu is fixed as world grows)."""
    return [rank + world * j for j in range(images_per_gpu)]


def global_images_per_step(world, images_per_gpu):
    return world * images_per_gpu


def max_over_ranks(t):
    """Max over ranks of a 1-element tensor (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t
