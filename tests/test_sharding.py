"""Host-side multi-GPU logic: image i -> rank i mod world (reference rule, image_restoration_model.py:731), no collective
on the data path, max-over-ranks timing reduction.  Runs with the gloo backend, world_size 2, on CPU."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fdn_tip2025_b200 import sharding
    per_gpu = 3
    idx = sharding.shard_indices(rank, world, per_gpu)
    # every rank times its own shard; the job time is the max over ranks (bench.py contract)
    ms = torch.tensor([10.0 + 5.0 * rank])
    job_ms = sharding.max_over_ranks(ms)
    gathered = [None] * world
    dist.all_gather_object(gathered, idx)      # test-only collective to inspect the partition
    if rank == 0:
        out.put((gathered, job_ms.item(), sharding.global_images_per_step(world, per_gpu)))
    dist.barrier()
    dist.destroy_process_group()


def test_image_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, job_ms, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gathered == [[0, 2, 4], [1, 3, 5]]            # image i -> rank i mod world, disjoint and complete
    assert sorted(sum(gathered, [])) == list(range(6))
    assert job_ms == 15.0                                 # max over ranks
    assert total == 6


def test_sharding_single_process():
    from fdn_tip2025_b200 import sharding
    assert sharding.shard_indices(0, 1, 4) == [0, 1, 2, 3]
    assert sharding.shard_indices(3, 8, 8) == [3 + 8 * j for j in range(8)]
    assert sharding.max_over_ranks(torch.tensor([7.0])).item() == 7.0
