"""world_size-2 `gloo` test (CPU) of the data-parallel host logic: shard the global batch, weight
each rank's mean-loss gradient by n_local/n_global, SUM all-reduce -> equals the single-process gradient of
the full batch.  The oracle stands in for the CUDA replica (BatchNorm off: per-replica statistics are a
documented deviation)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from oracle import keras_unet as ko
    from s2s_ismr_unet_b200.parallel import allreduce_mean_grads, shard_batch, shard_weight
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = ko.UnetConfig(H=16, W=16, Cin=3, filters=2, n_blocks=2, ct_kernel=3, bn=False)
    w = ko.random_init(cfg, 0)
    rng = np.random.default_rng(0)
    N = 5                                                     # uneven shards: 3 + 2
    x = rng.normal(size=(N, 16, 16, 3)).astype(np.float32)
    y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (N, 16, 16))]
    sl = shard_batch(N, rank, world)
    net = ko.UnetOracle(cfg, w)
    _, _, g = net.backward(x[sl], y[sl], grad_scale=shard_weight(sl.stop - sl.start, N))
    flat = torch.cat([g[n].reshape(-1) for n in net.trainable])
    allreduce_mean_grads(flat, sl.stop - sl.start, N, dist)
    if rank == 0:
        _, _, gfull = ko.UnetOracle(cfg, w).backward(x, y)
        ref = torch.cat([gfull[n].reshape(-1) for n in net.trainable])
        ret["err"] = float((flat - ref).abs().max() / ref.abs().max())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_allreduce_to_the_full_batch_gradient():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["err"] < 1e-12, ret["err"]


def _handle_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from s2s_ismr_unet_b200.parallel import exchange_ipc_handles
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank + 1]) * 64                     # stands in for this rank's cudaIpcMemHandle_t
    blob = exchange_ipc_handles(mine, dist)
    ret[rank] = blob
    try:
        exchange_ipc_handles(b"short", dist)
        ret[f"bad{rank}"] = False
    except ValueError:
        ret[f"bad{rank}"] = True
    dist.barrier()
    dist.destroy_process_group()


def test_peer_dp_rendezvous_orders_the_ipc_handles_by_rank():
    """The only host-side exchange of the peer-memory data-parallel path (csrc/dp.cuh): every rank must hand
    s2s_dp_connect the same world x 64-byte table in rank order."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_handle_worker, args=(2, port, ret), nprocs=2, join=True)
    want = bytes([1]) * 64 + bytes([2]) * 64
    assert ret[0] == want and ret[1] == want
    assert ret["bad0"] and ret["bad1"]
