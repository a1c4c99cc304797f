"""CPU suite: the N > 1 host logic (image/RoI sharding, max-over-ranks timing,
variable-size result gather) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from arfe_b200 import shard
    from arfe_b200 import workload as wl
    num_images, K = 5, 203
    rois = wl.synthetic_rois(K, batch=num_images, seed=0)
    imgs = shard.shard_images(num_images, rank, world)
    local, keep = shard.shard_rois(rois, imgs)
    # local batch indices index the rank's own images
    assert local.shape[0] == keep.numel()
    assert local.numel() == 0 or (0 <= int(local[:, 0].min()) and int(local[:, 0].max()) < len(imgs))
    assert torch.equal(local[:, 1:], rois[keep][:, 1:])
    assert all(imgs[int(b)] == int(g) for b, g in zip(local[:, 0], rois[keep][:, 0]))
    t = shard.max_over_ranks(1.0 + rank)
    assert t == float(world)
    parts = shard.gather_variable(keep)
    if rank == 0:
        allkeep = torch.cat(parts).sort().values
        out["ok"] = bool(torch.equal(allkeep, torch.arange(K)))        # partition: every RoI exactly once
        out["sizes"] = [int(p.numel()) for p in parts]
        out["imgs"] = [shard.shard_images(num_images, r, world) for r in range(world)]
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partition_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out["ok"]
    assert sum(out["sizes"]) == 203
    assert out["imgs"] == [[0, 1, 2], [3, 4]]


def test_shard_helpers_single_process():
    from arfe_b200 import shard
    assert shard.shard_images(8, 3, 4) == [6, 7]
    assert [len(shard.shard_images(7, r, 3)) for r in range(3)] == [3, 2, 2]
    rois = torch.tensor([[0, 1, 1, 5, 5], [2, 0, 0, 3, 3], [1, 2, 2, 9, 9], [2, 4, 4, 8, 8.]])
    loc, keep = shard.shard_rois(rois, [2])
    assert keep.tolist() == [1, 3] and loc[:, 0].tolist() == [0.0, 0.0]
    loc, keep = shard.shard_rois(rois, [])
    assert loc.shape == (0, 5)
    assert shard.max_over_ranks(2.5) == 2.5
