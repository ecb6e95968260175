"""Multi-GPU host logic on the CPU: world_size-2 gloo.  Captures are sharded in contiguous
ranges (all blocks of a capture on one rank); the per-rank partial normal-equation pieces
that the CUDA path allreduces (tag blocks, focal terms, cost) must sum to the full ones."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials(po, m, cap_idx, tag_idx, obs):
    cost, r, jc, jp, ja = po.evaluate(cap_idx, tag_idx, obs, m.cam0, m.cap0, m.tag0)
    Ht = np.zeros((m.n_tag, 6, 6))
    gt = np.zeros((m.n_tag, 6))
    np.add.at(Ht, tag_idx, np.einsum("bki,bkj->bij", ja, ja))
    np.add.at(gt, tag_idx, np.einsum("bki,bk->bi", ja, r))
    cam = np.array([np.sum(jc[:, :, 0] ** 2), np.sum(jc[:, :, 0] * r), cost])
    return np.concatenate([Ht.ravel(), gt.ravel(), cam])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench
    from ar_slam_b200 import synth
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = synth.make_map(400, 60, seed=13)
    ci, ti, ob = bench.shard(m, rank, world)
    local = torch.from_numpy(_partials(po, m, ci, ti, ob))
    # every capture lives on exactly one rank
    owner = torch.zeros(m.n_cap, dtype=torch.int64)
    owner[np.unique(ci)] = 1
    dist.all_reduce(owner)
    dist.all_reduce(local)
    nblk = torch.tensor([len(ci)])
    dist.all_reduce(nblk)
    if rank == 0:
        full = _partials(po, m, m.cap_idx, m.tag_idx, m.obs)
        q.put((bool((owner == 1).all()), int(nblk.item()) == len(m.cap_idx),
               float(np.abs(local.numpy() - full).max() / np.abs(full).max())))
    dist.destroy_process_group()


def test_capture_sharding_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    one_owner, all_blocks, err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert one_owner and all_blocks and err < 1e-12
