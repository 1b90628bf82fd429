"""2-GPU NCCL test of DataParallelTrainer (skipped with fewer than 2 GPUs): the all-reduced, shard-weighted
gradients and the post-Adam weights of every replica equal a single-GPU step on the concatenated batch
(BatchNorm off; with BatchNorm on the statistics are per replica by design)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from oracle import keras_unet as ko
    from s2s_ismr_unet_b200.model import Model
    from s2s_ismr_unet_b200.parallel import DataParallelTrainer, shard_batch
    from s2s_ismr_unet_b200.runtime import set_device
    torch.cuda.set_device(rank)
    set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = ko.UnetConfig(H=32, W=32, Cin=3, filters=2, n_blocks=3, ct_kernel=3, bn=False)
    w = ko.random_init(cfg, 0)
    rng = np.random.default_rng(0)
    N = 8
    x = rng.normal(size=(N, 32, 32, 3)).astype(np.float32)
    y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (N, 32, 32))]
    m = Model((32, 32, 3), bn=False, max_batch=N, weights=w)
    m.compile(loss="categorical_crossentropy")
    tr = DataParallelTrainer(m)
    tr.broadcast_weights(0)
    sl = shard_batch(N, rank, world)
    for _ in range(2):
        tr.train_on_batch(x[sl], y[sl], n_global=N)
    got = m.get_weights()
    if rank == 0:
        ref = Model((32, 32, 3), bn=False, max_batch=N, weights=w)
        ref.compile(loss="categorical_crossentropy")
        for _ in range(2):
            ref.train_on_batch(x, y)
        rw = ref.get_weights()
        ret["err"] = max(float(np.abs(got[k] - rw[k]).max()) for k in got)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_data_parallel_matches_single_gpu():
    from s2s_ismr_unet_b200.runtime import device_count
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29600 + (os.getpid() % 2000), ret), nprocs=2, join=True)
    assert ret["err"] < 2e-6, ret["err"]
