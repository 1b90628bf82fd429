"""Peer-memory data parallelism (csrc/dp.cuh): the gradient all-reduce fused with Adam and the sync-BN exchange
are this library's own kernels over CUDA-IPC mapped peer buffers.

* world 1 (any GPU box): the DP step runs the same kernels with no peer and must equal the plain train step.
* world 2 (skipped with fewer than 2 GPUs): 2 ranks x 4 samples with sync-BN == one GPU on the batch of 8
  (BatchNormalization ON: statistics of the global batch), and the replicas stay bit-identical."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data(N=8, seed=0):
    rng = np.random.default_rng(seed)
    x = (rng.gamma(2.0, 3.0, size=(N, 32, 32, 3)) / 6.0).astype(np.float32)
    y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (N, 32, 32))]
    return x, y


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("sync_bn", [False, True])
def test_world1_dp_step_equals_plain_step(sync_bn, pinned):
    from oracle import keras_unet as ko
    from s2s_ismr_unet_b200.model import Model
    from s2s_ismr_unet_b200.parallel import PeerDataParallelTrainer
    cfg = ko.UnetConfig(H=32, W=32, Cin=3, filters=2, n_blocks=3, ct_kernel=3)
    w = ko.random_init(cfg, 0)
    x, y = _data()
    if pinned:                                   # pinned host batches take the single-call end-to-end entry points
        from s2s_ismr_unet_b200.runtime import pinned_empty
        px, py = pinned_empty(x.shape), pinned_empty(y.shape)
        px[...], py[...] = x, y
        x, y = px, py
    a = Model((32, 32, 3), max_batch=8, weights=w)
    a.compile(loss="categorical_crossentropy")
    b = Model((32, 32, 3), max_batch=8, weights=w)
    b.compile(loss="categorical_crossentropy")
    tr = PeerDataParallelTrainer(b, sync_bn=sync_bn)
    for _ in range(3):
        la, _ = a.train_on_batch(x, y)
        lb, _ = tr.train_on_batch(x, y)
        assert abs(la - lb) <= 1e-6 * abs(la), (la, lb)
    tr.check()
    wa, wb = a.get_weights(), b.get_weights()
    err = max(float(np.abs(wa[k] - wb[k]).max()) for k in wa)
    assert err <= (1e-6 if sync_bn else 0.0), err
    tr.close()
    a.close(), b.close()


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from oracle import keras_unet as ko
    from s2s_ismr_unet_b200.model import Model
    from s2s_ismr_unet_b200.parallel import PeerDataParallelTrainer, shard_batch
    from s2s_ismr_unet_b200.runtime import set_device
    set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # rendezvous only: no NCCL on the data path
    cfg = ko.UnetConfig(H=32, W=32, Cin=3, filters=2, n_blocks=3, ct_kernel=3)
    w = ko.random_init(cfg, 0)
    N = 8
    x, y = _data(N)
    m = Model((32, 32, 3), max_batch=N, weights=w)
    m.compile(loss="categorical_crossentropy")
    tr = PeerDataParallelTrainer(m, sync_bn=True)
    sl = shard_batch(N, rank, world)
    from s2s_ismr_unet_b200.runtime import pinned_empty
    xs, ys = pinned_empty(x[sl].shape), pinned_empty(y[sl].shape)      # pinned shards: s2s_unet_dp_train_step_host
    xs[...], ys[...] = x[sl], y[sl]
    losses = [tr.train_on_batch(x[sl], y[sl], n_global=N)[0]] + [tr.train_on_batch(xs, ys, n_global=N)[0] for _ in range(2)]
    # two more steps through the streamed entry point (s2s_unet_dp_train_steps_host: shard i + 1 staged while step i computes)
    losses += [float(v) for v in tr.train_on_batches([xs, xs], [ys, ys], n_global=N)[:, 0]]
    tr.check()
    got = m.get_weights()
    flat = np.concatenate([got[k].ravel() for k in sorted(got)])
    gathered = [None] * world
    dist.all_gather_object(gathered, flat.tobytes())
    if rank == 0:
        ret["replicas_identical"] = all(g == gathered[0] for g in gathered)
        ref = Model((32, 32, 3), max_batch=N, weights=w)
        ref.compile(loss="categorical_crossentropy")
        ref_losses = [ref.train_on_batch(x, y)[0] for _ in range(5)]
        rw = ref.get_weights()
        ret["err"] = max(float(np.abs(got[k] - rw[k]).max()) for k in got)
        ret["loss_err"] = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
    dist.barrier()
    tr.close()
    dist.destroy_process_group()


def test_two_gpu_sync_bn_matches_single_gpu_batch():
    from s2s_ismr_unet_b200.runtime import device_count
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29700 + (os.getpid() % 2000), ret), nprocs=2, join=True)
    assert ret["replicas_identical"]
    assert ret["loss_err"] < 1e-5, ret["loss_err"]
    assert ret["err"] < 2e-5, ret["err"]
