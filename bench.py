#!/usr/bin/env python
"""bench.py — U-Net train samples/s (fwd + loss + bwd + Adam) on B200, with roofline, CPU baseline and e2e.

    python bench.py --gpus 1 --steps 50 --warmup 5                 # this framework (CUDA, sm_100a)
    python bench.py --impl reference --steps 20 --warmup 3         # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W  # batch-sharded DP, NCCL all-reduce

Workload (BASELINE.json metric config): MME U-Net, stacked predictor channels C=3,
64x64 India grid, filters=2, n_blocks=3, ct_kernel=3, batch 16 per GPU (weak scaling), fp32, synthetic
Gamma(2,3) fields with tercile one-hot labels.  A "step" is one optimiser step on one batch.
  value : on-device throughput; the batch is gathered from a device-resident data set that is larger
          than L2 (402 MB vs 126 MB), timed with CUDA events on the launch stream, max over ranks.
  e2e   : the same step through Model.train_on_batch with HOST numpy batches: pinned staging, H2D,
          step, D2H of {loss, accuracy}; wall-clock bracketed by stream synchronisation.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(H=64, W=64, Cin=3, filters=2, n_blocks=3, ct_kernel=3)
WORKLOAD_NAME = "MME U-Net 64x64 C=3 filters=2 n_blocks=3 ct=3 (tune_MME.py, SURVEY C3)"
# profiler tag -> __global__ function that ran it (csrc/): the roofline line names the dominant KERNEL
KERNEL_OF_TAG = {"conv3x3_fwd": "gconv_kernel", "conv3x3_dgrad": "gconv_kernel", "convT_dgrad": "gconv_kernel",
                 "conv3x3_wgrad": "wgrad_kernel", "convT_wgrad": "wgrad_kernel", "convT_fwd": "convt_fwd_kernel",
                 "conv3x3_fwd_tf32": "tc3conv_kernel", "conv3x3_dgrad_tf32": "tc3conv_kernel",
                 "conv3x3_fwd_tcgen05": "tcconv_kernel", "bn_apply": "bn_apply_kernel", "bn_apply_pool": "bn_apply_kernel",
                 "bn_bwd_reduce": "bn_bwd_reduce_kernel", "bn_bwd_apply": "bn_bwd_apply_kernel", "head": "head_kernel",
                 "grad_reduce_adam": "grad_reduce_adam_kernel", "grad_reduce": "grad_reduce_adam_kernel",
                 "wprep_dgrad": "wprep_kernel", "wprep_tf32": "tc3_wprep_kernel", "chansum": "chansum_kernel"}


def line_config(batch, n_gpus, dataset_samples):
    """The `config` object of the JSON line: IDENTICAL for the product arm and the reference arm (same workload, batch,
    data set), so that the driver's same-config check holds; arm-specific details go to the line's `setup` key."""
    T = max(dataset_samples, batch * 8)
    nbytes = T * WORKLOAD["H"] * WORKLOAD["W"] * (WORKLOAD["Cin"] + 3) * 4
    return {"workload": WORKLOAD_NAME, "batch_per_gpu": batch, "global_batch": batch * n_gpus, **WORKLOAD,
            "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single",
            "l2": f"inputs larger than L2: batches drawn from a {nbytes / 1e6:.0f} MB synthetic data set of {T} samples"}


# --------------------------------------------------------------------------------------------
def synth_dataset(T, H, W, Cin, seed=1234):
    """Deterministic synthetic hindcast set: x ~ Gamma(2,3), y = 0.5*mean_C(x) + 0.5*noise,
    elliptical land mask (~40 % land, ocean -> 0 as preprocessing.py:342-343), tercile one-hot labels."""
    rng = np.random.default_rng(seed)
    x = rng.gamma(2.0, 3.0, size=(T, H, W, Cin)).astype(np.float32)
    yf = 0.5 * x.mean(-1) + 0.5 * rng.gamma(2.0, 3.0, size=(T, H, W)).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    land = (((yy - H / 2) / (0.42 * H)) ** 2 + ((xx - W / 2) / (0.30 * W)) ** 2) <= 1.0
    x *= land[None, :, :, None]
    yf *= land[None]
    e = np.quantile(yf[: min(T, 512)], [1 / 3, 2 / 3], axis=0)
    lab = np.where(yf < e[0], 0, np.where(yf > e[1], 2, 1))
    y = np.eye(3, dtype=np.float32)[lab]
    return x, y, land


def algorithmic_work(cfg, train=True):
    """(flops, bytes) per sample, SURVEY §8d definition: conv/convT/pool layers, (in+out) elements * 4 B,
    x3 for fwd+dgrad+wgrad."""
    H, W, nb, f, k, cin = cfg["H"], cfg["W"], cfg["n_blocks"], cfg["filters"], cfg["ct_kernel"], cfg["Cin"]
    fl = by = 0.0

    def conv(h, w, ci, co, kk=3):
        nonlocal fl, by
        fl += 2.0 * kk * kk * ci * co * h * w
        by += 4.0 * h * w * (ci + co)

    c_prev = cin
    for b in range(nb):
        h, w, c = H >> b, W >> b, f * 4 * 2 ** b
        conv(h, w, c_prev, c), conv(h, w, c, c)
        by += 4.0 * (h * w * c + h * w * c / 4)      # pool
        c_prev = c
    h, w, c = H >> nb, W >> nb, f * 4 * 2 ** nb
    conv(h, w, c_prev, c), conv(h, w, c, c)
    for b in range(nb - 1, -1, -1):
        h, w, c = H >> b, W >> b, f * 4 * 2 ** b
        fl += 2.0 * k * k * 2 * c * c * (h // 2) * (w // 2)
        by += 4.0 * ((h // 2) * (w // 2) * 2 * c + h * w * c)
        conv(h, w, 2 * c, c), conv(h, w, c, c)
    c0 = f * 4
    fl += 2.0 * c0 * 3 * H * W                       # 1x1 head
    by += 4.0 * H * W * (c0 + 3)
    m = 3.0 if train else 1.0
    return fl * m, by * m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started before the warm-up (nvidia-smi needs ~100 ms
    to come up); rows are stamped on arrival and `mark()`/`summary()` keep those that fell inside the timed region
    (all rows taken under load if the region was shorter than the sampling period)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, enabled=True):
        self.rows, self.proc, self.th, self.idx = [], None, None, gpu_index
        self.t0 = self.t1 = None
        self.enabled = enabled
        self.t_load_end = None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=10.0):
        """nvidia-smi needs 0.1-2 s to print its first row (longer when 8 ranks start together): block until it has."""
        t = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.02)
        return bool(self.rows)

    def rows_since(self, t):
        return sum(1 for ts, _ in self.rows if ts >= t)

    def mark(self, start: bool):
        if start:
            self.t0 = time.perf_counter()
        else:
            self.t1 = time.perf_counter()

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= t <= self.t1 + 0.06]
        # a timed region shorter than the 50 ms sampling period is followed by untimed steps of the SAME work until two rows
        # have landed (t_load_end): those rows are taken under the same load
        end = self.t_load_end if self.t_load_end is not None else (self.t1 or 0) + 0.06
        rows = inside if len(inside) >= 2 else [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= end + 0.06]
        if not rows:
            rows = [r for t, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside),
                "sampling": "nvidia-smi -lms 50 on rank 0's GPU; rows inside the timed region, else rows taken while the same "
                            "steps kept running right after it"}


# --------------------------------------------------------------------------------------------
def cpu_port_throughput(batch, budget_s=20.0, max_steps=10_000, warmup=2, threads=None, seed=42, dataset_samples=None, cfg_kw=None):
    """The reference's CPU path as restated by oracle/ (torch-CPU fp32, all host threads), timed on a
    bounded sample of the same workload.  Returns (samples_per_s, steps, threads, seconds)."""
    import torch
    from oracle import keras_unet as ko
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = ko.UnetConfig(**(cfg_kw or WORKLOAD))
    net = ko.UnetOracle(cfg, ko.glorot_uniform_init(cfg, seed), dtype=torch.float32)
    net.compile(lr=1e-3)
    nb = max(4, (dataset_samples or 0) // batch)
    x, y, _ = synth_dataset(batch * nb, cfg.H, cfg.W, cfg.Cin)
    for i in range(warmup):
        net.train_step(x[:batch], y[:batch])
    t0 = time.perf_counter()
    steps = 0
    while steps < max_steps:
        j = (steps % nb) * batch
        net.train_step(x[j:j + batch], y[j:j + batch])
        steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return steps * batch / dt, steps, threads, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    Wm = max(args.warmup, 3)                      # the same warm-up count as the product arm
    sps, steps, threads, dt = cpu_port_throughput(args.batch, budget_s=min(150.0, 4.0 * args.steps), max_steps=args.steps,
                                                  warmup=Wm, dataset_samples=max(args.dataset_samples, args.batch * 8))
    line = {
        "impl": "reference", "metric": "U-Net train samples/s (fwd+bwd)", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": Wm, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": line_config(args.batch, args.gpus, args.dataset_samples),
        "setup": {"what": "oracle/keras_unet.py (torch-CPU fp32 restatement of the reference's Keras fit step) on rank 0's host cores; "
                          "under torchrun the other ranks exit without work"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} train steps of batch {args.batch} (torch-CPU fp32 restatement of the Keras path; "
                                   "TensorFlow/Keras are not installable here)"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU per step (reference batch_size=16)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="STRONG scaling (SURVEY C3): fix the global batch (e.g. 128) and give every GPU global/N samples; implies "
                         "--sync-bn so that N GPUs train exactly like one device on the global batch.  0 = weak scaling (--batch per GPU)")
    ap.add_argument("--dataset-samples", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl"],
                    help="N>1 exchange: 'peer' = this library's fused all-reduce+Adam / sync-BN kernels over NVLink peer memory "
                         "(CUDA IPC; torch.distributed/gloo is only the rendezvous), 'nccl' = ncclAllReduce between backward and Adam")
    ap.add_argument("--sync-bn", action="store_true",
                    help="peer DP with BatchNormalization statistics of the GLOBAL batch (exchanged over peer memory): N GPUs x B "
                         "samples train exactly like one device on N*B; default = per-replica statistics (each replica sees the "
                         "reference's batch of 16)")
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--large-batch", type=int, default=128,
                    help="secondary measurement: the same step at the strong-scaling global batch (SURVEY C3); 0 = skip")
    ap.add_argument("--inference-c5", type=int, default=1,
                    help="secondary measurement: config 5 inference (256x256, C=3, batch 64), fp32 vs bf16 tensor-core mode; 0 = skip")
    ap.add_argument("--extras", type=int, default=1,
                    help="secondary measurements promised by BASELINE.md: one C1 fit epoch, predict at batch 32, the largest tuning-grid "
                         "point (fp32 and tf32), each next to the CPU port; 0 = skip")
    ap.add_argument("--concurrent-models", type=int, default=8,
                    help="secondary measurement: K independent U-Net fits (sweep trials) on K streams of one GPU; 0 = skip")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    strong = args.global_batch > 0
    if strong:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} GPUs")
        args.batch = args.global_batch // world
        args.sync_bn = world > 1
        args.dp = "peer"

    from s2s_ismr_unet_b200 import _lib, model as s2s_model
    from s2s_ismr_unet_b200.keras_api.optimizers import Adam
    from s2s_ismr_unet_b200.runtime import DeviceBuffer, Event, set_device

    set_device(local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        if args.dp == "nccl":
            os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the one JSON line
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("gloo")

    cfg = WORKLOAD
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    s2s_model.set_seed(42)
    m = s2s_model.Model((cfg["H"], cfg["W"], cfg["Cin"]), filters=cfg["filters"], n_blocks=cfg["n_blocks"],
                        ct_kernel=cfg["ct_kernel"], max_batch=B)
    m.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy", metrics=["accuracy"])
    m.set_graphs(not args.no_graphs)
    call = _lib.call
    st = m.stream

    T = max(args.dataset_samples, B * 8)
    x, y, _ = synth_dataset(T, cfg["H"], cfg["W"], cfg["Cin"], seed=1234 + rank)
    dx, dy = DeviceBuffer.from_array(x, st), DeviceBuffer.from_array(y, st)
    dataset_bytes = x.nbytes + y.nbytes
    xrow, yrow = x[0].nbytes, y[0].nbytes

    # ---- data-parallel plumbing: NCCL all-reduce of the dense grad arena between backward and Adam
    peer = None
    nccl_pg = None                      # None = default group
    timing_on_cuda = world > 1 and args.dp == "nccl"
    if world > 1 and args.dp == "peer":
        from s2s_ismr_unet_b200.parallel import PeerDataParallelTrainer
        try:
            peer = PeerDataParallelTrainer(m, sync_bn=args.sync_bn)
            peer.broadcast_weights(0)
        except RuntimeError as e:       # raised on EVERY rank (the set-up agrees on its outcome): fall back to NCCL
            if rank == 0:
                print(f"bench: {e}; falling back to ncclAllReduce + Adam", file=sys.stderr, flush=True)
            peer = None
            args.dp = "nccl"
            os.environ.setdefault("NCCL_DEBUG", "WARN")
            nccl_pg = dist.new_group(backend="nccl")
    if world > 1 and args.dp == "nccl":
        import torch

        class _Ptr:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        grad_t = torch.as_tensor(_Ptr(m._grads_ptr, m.n_params_padded), device=f"cuda:{local_rank}")
        ext = torch.cuda.ExternalStream(st.ptr, device=local_rank)

    def dev_step(i):
        """One optimiser step on batch i of the device-resident data set."""
        j = (i * B) % (T - B + 1)
        xp, yp = C.c_void_p(dx.ptr + j * xrow), C.c_void_p(dy.ptr + j * yrow)
        if world == 1:
            call("s2s_unet_train_step", m._h, xp, yp, None, B, None, m.sp)
        elif peer is not None:
            call("s2s_unet_dp_train_step", m._h, xp, yp, B, B * world, None, m.sp)
        else:
            call("s2s_unet_backward_only", m._h, xp, yp, None, B, C.c_float(1.0 / world), None, m.sp)
            with torch.cuda.stream(ext):
                dist.all_reduce(grad_t, group=nccl_pg)
            call("s2s_unet_apply_adam", m._h, m.sp)

    def barrier():
        st.synchronize()
        if world > 1:
            dist.barrier()
            st.synchronize()

    def max_over_ranks(v):
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}" if timing_on_cuda else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- `value`: device-resident inputs, CUDA events on the launch stream
    e0, e1 = Event(), Event()
    with ClockSampler(local_rank, enabled=rank == 0) as clk:
        for i in range(Wm):
            dev_step(i)
        clk.wait_first()                # nvidia-smi is up and printing before the timed region starts
        barrier()
        l0 = m.launch_count()
        clk.mark(True)
        e0.record(st)
        for i in range(K):
            dev_step(Wm + i)
        e1.record(st)
        barrier()
        clk.mark(False)
        launches = m.launch_count() - l0
        # a timed region shorter than the sampling period: keep the same steps running (untimed, all ranks in lockstep by
        # a fixed count) so that the clock rows are taken under this load
        extra = 0 if K * 0.45e-3 > 0.25 else int(0.3 / 0.45e-3)
        for i in range(extra):
            dev_step(Wm + K + i)
        barrier()
        clk.t_load_end = time.perf_counter()
    ms = e0.elapsed_ms(e1)
    if world > 1:
        ms = max_over_ranks(ms)
        if peer is not None:
            peer.check()
    value = world * B * K / (ms * 1e-3)

    # ---- `e2e`: host numpy batches through the public API (pinned staging + H2D + step + D2H of the loss)
    # host batches live in pinned memory (the e2e contract: H2D of each step's inputs from pinned host memory)
    from s2s_ismr_unet_b200.runtime import pinned_empty
    hx, hy = [], []
    for i in range(8):
        j = (i * B) % (T - B + 1)
        bx, by_ = pinned_empty((B,) + x.shape[1:]), pinned_empty((B,) + y.shape[1:])
        bx[...] = x[j:j + B]
        by_[...] = y[j:j + B]
        hx.append(bx), hy.append(by_)
    e2e_sps = e2e_pageable_sps = e2e_call_sps = e2e_pageable_stream_sps = None
    if world == 1:
        for i in range(Wm):
            m.train_on_batch(hx[i % 8], hy[i % 8])
        st.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            m.train_on_batch(hx[i % 8], hy[i % 8])
        st.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_call_sps = B * K / e2e_s
        # the streaming form of the same public API (Model.train_on_batches -> s2s_unet_train_steps_host; what model.fit does with
        # host arrays): every step still copies its own batch H2D from pinned memory and its own {loss, accuracy} D2H, but the
        # copy of batch i + 1 is staged while step i computes
        m.train_on_batches([hx[i % 8] for i in range(Wm)], [hy[i % 8] for i in range(Wm)])
        st.synchronize()
        t0 = time.perf_counter()
        losses_stream = m.train_on_batches([hx[i % 8] for i in range(K)], [hy[i % 8] for i in range(K)])
        st.synchronize()
        e2e_sps = B * K / (time.perf_counter() - t0)
        assert losses_stream.shape == (K, 2) and np.isfinite(losses_stream).all()
        # the same call on PAGEABLE NumPy batches (what the reference passes to model.fit): staged through the model's pinned
        # buffers by the host layer before the H2D copy
        for i in range(3):
            m.train_on_batch(x[i * B:(i + 1) * B], y[i * B:(i + 1) * B])
        st.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            j = (i * B) % (T - B + 1)
            m.train_on_batch(x[j:j + B], y[j:j + B])
        st.synchronize()
        e2e_pageable_sps = B * K / (time.perf_counter() - t0)
        # ... and streamed: the C call copies every pageable batch into a ring of pinned slots while the GPU computes
        pg = [(x[(i * B) % (T - B + 1):(i * B) % (T - B + 1) + B], y[(i * B) % (T - B + 1):(i * B) % (T - B + 1) + B]) for i in range(K)]
        m.train_on_batches([p[0] for p in pg[:Wm]], [p[1] for p in pg[:Wm]])
        st.synchronize()
        t0 = time.perf_counter()
        m.train_on_batches([p[0] for p in pg], [p[1] for p in pg])
        st.synchronize()
        e2e_pageable_stream_sps = B * K / (time.perf_counter() - t0)
    else:
        def e2e_step(i):
            if peer is not None:
                return peer.train_on_batch(hx[i % 8], hy[i % 8], n_global=B * world)
            loss = m.backward_on_batch(hx[i % 8], hy[i % 8], grad_scale=1.0 / world)
            with torch.cuda.stream(ext):
                dist.all_reduce(grad_t, group=nccl_pg)
            m.apply_adam()
            return loss
        for i in range(Wm):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            e2e_step(i)
        barrier()
        e2e_sps = world * B * K / max_over_ranks(time.perf_counter() - t0)
        if peer is not None:
            # streamed form (PeerDataParallelTrainer.train_on_batches): every step copies its own shard H2D and its global
            # {loss, accuracy} D2H, the copy of shard i + 1 is staged while step i computes
            e2e_call_sps = e2e_sps
            peer.train_on_batches([hx[i % 8] for i in range(Wm)], [hy[i % 8] for i in range(Wm)], n_global=B * world)
            barrier()
            t0 = time.perf_counter()
            peer.train_on_batches([hx[i % 8] for i in range(K)], [hy[i % 8] for i in range(K)], n_global=B * world)
            barrier()
            e2e_sps = world * B * K / max_over_ranks(time.perf_counter() - t0)

    # ---- N > 1: the parity-exact variant (sync-BN: N GPUs x B samples == one device on N*B, training.py:102) measured
    # beside `value`, and one sync-BN step on a fixed global batch checked against the fp64 oracle on rank 0
    dp_parity = sync_bn_sec = None
    if world > 1 and peer is not None:
        import hashlib
        from s2s_ismr_unet_b200.parallel import PeerDataParallelTrainer
        ms2 = s2s_model.Model((cfg["H"], cfg["W"], cfg["Cin"]), filters=cfg["filters"], n_blocks=cfg["n_blocks"],
                              ct_kernel=cfg["ct_kernel"], max_batch=B)
        ms2.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
        ms2.set_graphs(not args.no_graphs)
        peer2 = PeerDataParallelTrainer(ms2, sync_bn=True)
        peer2.broadcast_weights(0)

        def sstep(i):
            j = (i * B) % (T - B + 1)
            call("s2s_unet_dp_train_step", ms2._h, C.c_void_p(dx.ptr + j * xrow), C.c_void_p(dy.ptr + j * yrow), B, B * world, None, ms2.sp)
        for i in range(Wm):
            sstep(i)
        ms2.stream.synchronize(); dist.barrier()
        s0, s1 = Event(), Event()
        s0.record(ms2.stream)
        for i in range(K):
            sstep(Wm + i)
        s1.record(ms2.stream)
        ms2.stream.synchronize(); dist.barrier()
        peer2.check()
        ms_s = max_over_ranks(s0.elapsed_ms(s1))
        sync_bn_sec = {"value": world * B * K / (ms_s * 1e-3), "unit": "samples/s", "ms_per_step": ms_s / K,
                       "what": "the same step with BatchNormalization statistics of the GLOBAL batch exchanged over peer memory "
                               "(12 in-kernel exchanges per step): exactly the single-device step on the global batch"}
        # parity: fresh optimiser state, fixed global batch, every rank trains on its shard
        ms2.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
        w0 = ms2.get_weights()
        xg, yg, _ = synth_dataset(B * world, cfg["H"], cfg["W"], cfg["Cin"], seed=4321)
        loss_g, _ = peer2.train_on_batch(xg[rank * B:(rank + 1) * B], yg[rank * B:(rank + 1) * B], n_global=B * world)
        w1 = ms2.get_weights()
        digest = hashlib.sha256(b"".join(np.ascontiguousarray(w1[k]).tobytes() for k in sorted(w1))).hexdigest()
        digests = [None] * world
        dist.all_gather_object(digests, digest)
        if rank == 0:
            import torch
            from oracle import keras_unet as ko
            ocfg = ko.UnetConfig(**cfg)
            onet = ko.UnetOracle(ocfg, w0, dtype=torch.float64)
            onet.compile(lr=1e-3)
            loss_o, _ = onet.train_step(xg, yg)
            wo = onet.get_weights()
            rl2 = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64).ravel() - np.asarray(b, np.float64).ravel()) /
                                     max(np.linalg.norm(np.asarray(b, np.float64).ravel()), 1e-300))
            dp_parity = {"global_batch": B * world, "loss": float(loss_g), "oracle_loss": float(loss_o),
                         "rel_err": abs(float(loss_g) - float(loss_o)) / abs(float(loss_o)),
                         "weights_rel_l2_after_step": max(rl2(w1[k], wo[k]) for k in w1),
                         "replicas_identical": len(set(digests)) == 1,
                         "what": "one sync-BN data-parallel step on a fixed global batch vs the fp64 oracle's single-device step on the "
                                 "same batch (oracle/keras_unet.py, rank 0 host); replicas compared by the SHA-256 of their weights"}
        dist.barrier()
        peer2.close()
        ms2.close()

    # ---- roofline of the dominant kernel: per-launch CUDA events (eager replay of the same step)
    roof, table = None, {}
    if rank == 0:
        null_us = C.c_float(0)
        call("s2s_prof_null_us", C.byref(null_us), m.sp)
        bracket_us = max(0.0, null_us.value - 1.1)        # an empty kernel costs ~1.1 us in-stream (tools/graph_floor.cu)
        call("s2s_prof_enable", 1)
        for i in range(args.profile_steps):
            call("s2s_unet_train_step" if world == 1 else "s2s_unet_backward_only", m._h, C.c_void_p(dx.ptr), C.c_void_p(dy.ptr), None, B,
                 *([None, m.sp] if world == 1 else [C.c_float(1.0), None, m.sp]))
        st.synchronize()
        buf = C.create_string_buffer(1 << 16)
        call("s2s_prof_report", buf, C.c_size_t(len(buf)))
        call("s2s_prof_enable", 0)
        tot_ms = 0.0
        for line in buf.value.decode().strip().splitlines():
            tag, n, tms, by, fl = line.split(",")
            nl = int(n) // args.profile_steps
            kms = max(float(tms) / args.profile_steps - nl * bracket_us * 1e-3, 1e-6)     # minus the event-bracket overhead
            table[tag] = dict(launches=nl, ms=kms, bytes=float(by) / args.profile_steps, flops=float(fl) / args.profile_steps)
            tot_ms += kms
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        # fp32 CUDA-core peak MEASURED here (csrc/micro.cuh): the fp32 path is FFMA-bound (AI 27.6 flop/B > ridge ~11)
        f_s, f_p = C.c_float(0), C.c_float(0)
        call("s2s_ffma_peak", C.byref(f_s), C.byref(f_p), m.sp)
        ffma_peak = max(f_s.value, f_p.value)
        # the dominant KERNEL (not profiler tag): one __global__ function may serve several operators
        fam = {}
        for tag, t in table.items():
            k = KERNEL_OF_TAG.get(tag, tag)
            f = fam.setdefault(k, dict(launches=0, ms=0.0, bytes=0.0, flops=0.0, tags=[]))
            f["launches"] += t["launches"]; f["ms"] += t["ms"]; f["bytes"] += t["bytes"]; f["flops"] += t["flops"]; f["tags"].append(tag)
        top = max(fam, key=lambda k: fam[k]["ms"])
        tk = fam[top]
        ach_gbs = tk["bytes"] / (tk["ms"] * 1e-3) / 1e9
        ach_tf = tk["flops"] / (tk["ms"] * 1e-3) / 1e12
        traffic, tsrc = None, None
        for tj in (ROOT / "profiles" / "r2_traffic.json", ROOT / "profiles" / "r1_traffic.json"):
            if tj.exists():   # dram__bytes_read+write per launch from the committed ncu --set full capture of that kernel
                tt = json.loads(tj.read_text())
                ent = tt.get(top) or (tt.get("conv3x3_fwd_or_dgrad") if top == "gconv_kernel" else None)
                if ent:
                    traffic, tsrc = ent["dram_bytes_per_launch"], tt.get("_source")
                    break
        roof = {"bound": "ffma", "kernel": top, "tags": tk["tags"], "achieved": ach_tf, "peak": ffma_peak, "unit": "TFLOP/s",
                "frac": ach_tf / ffma_peak, "peak_source": "measured in this run (s2s_ffma_peak: scalar FFMA %.1f, packed FFMA2 %.1f TFLOP/s; "
                                                           "nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.5)" % (f_s.value, f_p.value),
                "ffma_peak_measured": ffma_peak,
                "traffic": traffic, "traffic_source": tsrc,
                "algorithmic_flops_per_launch": tk["flops"] / tk["launches"], "algorithmic_bytes_per_launch": tk["bytes"] / tk["launches"],
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": peak_src},
                "launches_per_step": tk["launches"], "avg_launch_us": 1e3 * tk["ms"] / tk["launches"],
                "share_of_kernel_time": tk["ms"] / tot_ms,
                "how": f"CUDA events around every launch on its own stream (minus the {bracket_us:.2f} us bracket overhead calibrated "
                       f"on an empty kernel), eager replay of {args.profile_steps} steps after the timed region; kernels grouped by "
                       "__global__ function, the family with the largest summed time is reported"}
        flops_s, bytes_s = algorithmic_work(cfg)
        roof["step"] = {"kernel_ms_sum": tot_ms, "graph_ms_per_step": ms / K,
                        "algorithmic_gb_per_step": (bytes_s * B + 28.0 * m.count_params()) / 1e9,
                        "hbm_frac_whole_step": (bytes_s * B + 28.0 * m.count_params()) / (ms / K * 1e-3) / 1e9 / hbm_peak,
                        "ffma_frac_whole_step": flops_s * B / (ms / K * 1e-3) / 1e12 / ffma_peak}
        roof["kernel_families"] = {k: dict(launches=v["launches"], ms=v["ms"], tflops=v["flops"] / (v["ms"] * 1e-3) / 1e12,
                                           gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9) for k, v in fam.items()}

    # ---- secondary: the same model at the strong-scaling global batch of SURVEY C3 (128) on one GPU, with its own
    # per-kernel roofline: at batch 16 every layer is latency-bound, this shows the kernels at a throughput size.
    large = None
    if rank == 0 and world == 1 and args.large_batch > B:
        LB = args.large_batch
        ml = s2s_model.Model((cfg["H"], cfg["W"], cfg["Cin"]), filters=cfg["filters"], n_blocks=cfg["n_blocks"],
                             ct_kernel=cfg["ct_kernel"], max_batch=LB)
        ml.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
        ml.set_graphs(not args.no_graphs)

        def lstep(i):
            j = (i * LB) % (T - LB + 1)
            call("s2s_unet_train_step", ml._h, C.c_void_p(dx.ptr + j * xrow), C.c_void_p(dy.ptr + j * yrow), None, LB, None, ml.sp)
        Kl = max(20, K // 4)
        for i in range(Wm):
            lstep(i)
        ml.stream.synchronize()
        a0, a1 = Event(), Event()
        a0.record(ml.stream)
        for i in range(Kl):
            lstep(Wm + i)
        a1.record(ml.stream)
        ml.stream.synchronize()
        lms = a0.elapsed_ms(a1) / Kl
        call("s2s_prof_enable", 1)
        lstep(0)
        ml.stream.synchronize()
        buf = C.create_string_buffer(1 << 16)
        call("s2s_prof_report", buf, C.c_size_t(len(buf)))
        call("s2s_prof_enable", 0)
        ltab = {}
        for line in buf.value.decode().strip().splitlines():
            tag, n, tms, by, fl = line.split(",")
            kms = max(float(tms) - int(n) * bracket_us * 1e-3, 1e-6)
            ltab[tag] = dict(launches=int(n), ms=kms, gbs=float(by) / kms / 1e6, tflops=float(fl) / kms / 1e9)
        ltop = max(ltab, key=lambda k: ltab[k]["ms"])
        fl_s, by_s = algorithmic_work(cfg)
        large = {"batch": LB, "value": LB / (lms * 1e-3), "unit": "samples/s", "ms_per_step": lms,
                 "hbm_frac_whole_step": (by_s * LB + 28.0 * m.count_params()) / (lms * 1e-3) / 1e9 / hbm_peak,
                 "ffma_frac_whole_step": fl_s * LB / (lms * 1e-3) / 1e12 / ffma_peak,
                 "top_kernel": ltop, "top_kernel_gbs": ltab[ltop]["gbs"], "top_kernel_tflops": ltab[ltop]["tflops"],
                 "top_kernel_hbm_frac": ltab[ltop]["gbs"] / hbm_peak, "kernels": ltab}
        ml.close()

        # the tensor-core training mode (precision="tf32": tcgen05 forward / dgrad / wgrad / transposed conv where the cost model
        # picks them) on the headline workload and on the large batch — reduced precision (tolerance 1e-2), reported beside fp32
        def tf32_point(bsz):
            mt = s2s_model.Model((cfg["H"], cfg["W"], cfg["Cin"]), filters=cfg["filters"], n_blocks=cfg["n_blocks"],
                                 ct_kernel=cfg["ct_kernel"], max_batch=bsz, precision="tf32")
            mt.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
            mt.set_graphs(not args.no_graphs)

            def tstep(i):
                j = (i * bsz) % (T - bsz + 1)
                call("s2s_unet_train_step", mt._h, C.c_void_p(dx.ptr + j * xrow), C.c_void_p(dy.ptr + j * yrow), None, bsz, None, mt.sp)
            kt = max(20, K // 4)
            for i in range(Wm):
                tstep(i)
            mt.stream.synchronize()
            t0_, t1_ = Event(), Event()
            t0_.record(mt.stream)
            for i in range(kt):
                tstep(Wm + i)
            t1_.record(mt.stream)
            mt.stream.synchronize()
            tms = t0_.elapsed_ms(t1_) / kt
            mt.close()
            return {"batch": bsz, "ms_per_step": tms, "samples_per_s": bsz / (tms * 1e-3)}
        large["tf32"] = tf32_point(LB)
        large["tf32"]["speedup_vs_fp32"] = large["tf32"]["samples_per_s"] / large["value"]
        large["tf32_headline_batch"] = tf32_point(B)
        large["tf32_headline_batch"]["speedup_vs_fp32"] = large["tf32_headline_batch"]["samples_per_s"] / (B * K / (ms * 1e-3))

    # ---- secondary: config 5 inference (real-time MME at 0.25 deg, 256x256): fp32 path vs tcgen05 bf16 mode
    infer = None
    if rank == 0 and world == 1 and args.inference_c5:
        IH, IB, IT = 256, 64, 256
        rng = np.random.default_rng(5)
        xi = rng.gamma(2.0, 3.0, size=(IT, IH, IH, 3)).astype(np.float32)
        dxi = DeviceBuffer.from_array(xi, st)
        dpi = DeviceBuffer(4 * IT * IH * IH * 3)
        infer = {"grid": f"{IH}x{IH}", "C": 3, "batch": IB, "samples": IT}
        ref_out = None
        for prec in ("fp32", "bf16_tc", "tf32"):
            mi = s2s_model.Model((IH, IH, 3), filters=cfg["filters"], n_blocks=cfg["n_blocks"], ct_kernel=cfg["ct_kernel"],
                                 max_batch=IB, precision=prec, weights=None)
            for _ in range(2):
                call("s2s_unet_predict_dataset", mi._h, C.c_void_p(dxi.ptr), IT, IB, C.c_void_p(dpi.ptr), mi.sp)
            mi.stream.synchronize()
            a0, a1 = Event(), Event()
            a0.record(mi.stream)
            for _ in range(3):
                call("s2s_unet_predict_dataset", mi._h, C.c_void_p(dxi.ptr), IT, IB, C.c_void_p(dpi.ptr), mi.sp)
            a1.record(mi.stream)
            mi.stream.synchronize()
            ms_i = a0.elapsed_ms(a1) / 3
            out = dpi.download((8, IH, IH, 3), np.float32, mi.stream)
            if prec == "fp32":
                ref_out, w_keep = out, mi.get_weights()
            infer[prec] = {"samples_per_s": IT / (ms_i * 1e-3), "ms_per_batch": ms_i / (IT / IB)}
            mi.close()
        fl_i, by_i = algorithmic_work(dict(cfg, H=IH, W=IH), train=False)
        infer["fp32"]["ffma_frac"] = fl_i * infer["fp32"]["samples_per_s"] / 1e12 / ffma_peak
        infer["fp32"]["hbm_frac"] = by_i * infer["fp32"]["samples_per_s"] / 1e9 / hbm_peak
        infer["bf16_tc"]["hbm_frac_fp32_bytes"] = by_i * infer["bf16_tc"]["samples_per_s"] / 1e9 / hbm_peak
        infer["tf32"]["hbm_frac"] = by_i * infer["tf32"]["samples_per_s"] / 1e9 / hbm_peak
        infer["tf32"]["speedup_vs_fp32"] = infer["tf32"]["samples_per_s"] / infer["fp32"]["samples_per_s"]
        infer["tf32"]["what"] = "every 3x3 conv with Cin % 8 == 0 on tcgen05 kind::tf32 straight from the fp32 NHWC tensors (no cast pass)"
        dxi.free(), dpi.free()

    # ---- secondary: config 5 skill maps (ACC / CC over an archive of 4096 starts at 256x256, RPS over 1024) and Grad-CAM:
    # streaming per-gridpoint reductions, reported as achieved GB/s of their ALGORITHMIC bytes against the HBM peak
    skill = None
    if rank == 0 and world == 1 and args.inference_c5:
        from s2s_ismr_unet_b200.runtime import d2d
        SY, T0, REP, NG = 256, 512, 8, 22
        YX = SY * SY
        rng = np.random.default_rng(6)
        TA = T0 * REP
        dxa, dya = DeviceBuffer(4 * TA * YX), DeviceBuffer(4 * TA * YX)
        for dst in (dxa, dya):
            blk = DeviceBuffer.from_array(rng.random((T0, YX), dtype=np.float32), st)
            for r in range(REP):
                d2d(dst.ptr + r * blk.nbytes, blk.ptr, blk.nbytes, st)
            st.synchronize()
            blk.free()
        week = np.arange(TA) % NG
        order = np.argsort(week, kind="stable").astype(np.int32)
        gstart = np.concatenate([[0], np.cumsum(np.bincount(week, minlength=NG))]).astype(np.int32)
        do_, dg_ = DeviceBuffer.from_array(order, st), DeviceBuffer.from_array(gstart, st)
        dacc, dcc = DeviceBuffer(4 * YX), DeviceBuffer(4 * YX)

        def timed(fn, reps=5):
            fn()
            st.synchronize()
            a0, a1 = Event(), Event()
            a0.record(st)
            for _ in range(reps):
                fn()
            a1.record(st)
            st.synchronize()
            return a0.elapsed_ms(a1) / reps
        ms_acc = timed(lambda: call("s2s_acc_map", C.c_void_p(dxa.ptr), C.c_void_p(dya.ptr), C.c_void_p(do_.ptr), C.c_void_p(dg_.ptr),
                                    NG, TA, SY, SY, C.c_void_p(dacc.ptr), C.c_void_p(dcc.ptr), C.c_void_p(st.ptr)))
        by_acc = 2.0 * 4 * TA * YX
        TR = 1024                                   # RPS: forecast + one-hot observation, 3 categories each (same buffers re-read as [TR,YX,3])
        drps = DeviceBuffer(4 * YX)
        ms_rps = timed(lambda: call("s2s_rps_map", C.c_void_p(dxa.ptr), C.c_void_p(dya.ptr), TR, SY, SY, C.c_void_p(drps.ptr), C.c_void_p(st.ptr)))
        by_rps = 2.0 * 4 * TR * YX * 3
        skill = {"grid": f"{SY}x{SY}", "acc_cc_map": {"T": TA, "iso_week_groups": NG, "ms": ms_acc, "gbs": by_acc / ms_acc / 1e6,
                                                      "hbm_frac": by_acc / ms_acc / 1e6 / hbm_peak},
                 "rps_map": {"T": TR, "ms": ms_rps, "gbs": by_rps / ms_rps / 1e6, "hbm_frac": by_rps / ms_rps / 1e6 / hbm_peak}}
        for b_ in (dxa, dya, do_, dg_, dacc, dcc, drps):
            b_.free()
        mg = s2s_model.Model((SY, SY, 3), filters=cfg["filters"], n_blocks=cfg["n_blocks"], ct_kernel=cfg["ct_kernel"], max_batch=64)
        dxg = DeviceBuffer.from_array(rng.gamma(2.0, 3.0, size=(64, SY, SY, 3)).astype(np.float32), mg.stream)
        dcam = DeviceBuffer(4 * 64 * (SY // 8) * (SY // 8))
        gc = lambda: call("s2s_unet_gradcam", mg._h, C.c_void_p(dxg.ptr), 64, b"bottleneck", 2, C.c_void_p(dcam.ptr), mg.sp)
        gc()
        mg.stream.synchronize()
        a0, a1 = Event(), Event()
        a0.record(mg.stream)
        for _ in range(5):
            gc()
        a1.record(mg.stream)
        mg.stream.synchronize()
        ms_gc = a0.elapsed_ms(a1) / 5
        skill["gradcam_bottleneck_above"] = {"batch": 64, "ms_per_batch": ms_gc, "samples_per_s": 64 / (ms_gc * 1e-3)}
        dxg.free(), dcam.free()
        mg.close()

    # ---- secondary: trial batching: K independent fits share the GPU, one stream each
    trial = None
    if rank == 0 and world == 1 and args.concurrent_models > 1:
        Kc = args.concurrent_models
        models = [m]
        for j in range(1, Kc):
            mj = s2s_model.Model((cfg["H"], cfg["W"], cfg["Cin"]), filters=cfg["filters"], n_blocks=cfg["n_blocks"],
                                 ct_kernel=cfg["ct_kernel"], max_batch=B)
            mj.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
            mj.set_graphs(not args.no_graphs)
            models.append(mj)

        def multi_step(i):
            j = (i * B) % (T - B + 1)
            xp, yp = C.c_void_p(dx.ptr + j * xrow), C.c_void_p(dy.ptr + j * yrow)
            for mj in models:
                call("s2s_unet_train_step", mj._h, xp, yp, None, B, None, mj.sp)
        for i in range(Wm):
            multi_step(i)
        for mj in models:
            mj.stream.synchronize()
        t0 = time.perf_counter()
        ev = [(Event(), Event()) for _ in models]
        for (a0, _), mj in zip(ev, models):
            a0.record(mj.stream)
        for i in range(K):
            multi_step(Wm + i)
        for (_, a1), mj in zip(ev, models):
            a1.record(mj.stream)
        for mj in models:
            mj.stream.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = max(a0.elapsed_ms(a1) for a0, a1 in ev)
        trial = {"concurrent_models": Kc, "value": Kc * B * K / max(wall, dev_ms * 1e-3), "unit": "samples/s",
                 "ms_per_round": 1e3 * max(wall, dev_ms * 1e-3) / K,
                 "note": "K independent fits (one CUDA graph + stream each) driven by one host thread; wall clock"}
        for mj in models[1:]:
            mj.close()

    # ---- secondary: one full C1 epoch through model.fit, predict at Keras' default batch 32, and the
    # largest grid point of the tuning grid, each next to the CPU port on this box's host cores
    c1 = pred32 = gridmax = None
    if rank == 0 and world == 1 and args.extras:
        import torch
        from oracle import keras_unet as ko
        from s2s_ismr_unet_b200.keras_api.callbacks import EarlyStopping
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        c1cfg = dict(H=64, W=64, Cin=1, filters=2, n_blocks=3, ct_kernel=3)
        xt, yt, _ = synth_dataset(261 + 65, 64, 64, 1, seed=77)        # 16 years: 261 train / 65 validation starts
        xtr, ytr, xva, yva = xt[:261], yt[:261], xt[261:], yt[261:]
        mc = s2s_model.Model((64, 64, 1), filters=2, n_blocks=3, ct_kernel=3, max_batch=32)
        mc.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy", metrics=["accuracy"])
        mc.fit(x=xtr, y=ytr, validation_data=(xva, yva), epochs=1, batch_size=16, shuffle=True, verbose=0)     # upload + graph capture
        t0 = time.perf_counter()
        EP = 5
        mc.fit(x=xtr, y=ytr, validation_data=(xva, yva), epochs=EP, batch_size=16, shuffle=True, verbose=0,
               callbacks=[EarlyStopping(monitor="val_loss", patience=10, restore_best_weights=True)])
        gpu_epoch_ms = 1e3 * (time.perf_counter() - t0) / EP
        ocfg = ko.UnetConfig(**c1cfg)
        onet = ko.UnetOracle(ocfg, ko.glorot_uniform_init(ocfg, 42), dtype=torch.float32)
        onet.compile(lr=1e-3)
        t0 = time.perf_counter()
        onet.fit(xtr, ytr, (xva, yva), 1, 16, [np.random.default_rng(0).permutation(261)], patience=10)
        cpu_epoch_ms = 1e3 * (time.perf_counter() - t0)
        c1 = {"what": "one model.fit epoch of config C1 (261 starts = 17 steps of batch 16 incl. the last batch of 5, then the "
                      "65-start validation pass, callbacks; training.py:102-103), wall clock through the public API, host NumPy data",
              "gpu_ms_per_epoch": gpu_epoch_ms, "cpu_port_ms_per_epoch": cpu_epoch_ms, "cpu_cores": threads,
              "speedup": cpu_epoch_ms / gpu_epoch_ms, "epochs_timed": EP}
        # predict at Keras' default batch size 32 (training.py:133-135), host array in, host array out
        mc.predict(xtr, verbose=0)
        t0 = time.perf_counter()
        for _ in range(5):
            pg = mc.predict(xtr, verbose=0)
        gpu_pred = 5 * len(xtr) / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        pc = onet.predict(xtr, batch_size=32)
        cpu_pred = len(xtr) / (time.perf_counter() - t0)
        pred32 = {"what": "model.predict(X_train (261,64,64,1), batch 32) end to end (H2D, forward, D2H)", "gpu_samples_per_s": gpu_pred,
                  "cpu_port_samples_per_s": cpu_pred, "cpu_cores": threads, "speedup": gpu_pred / cpu_pred}
        mc.close()
        # the largest point of the tuning grid (tune_MME.py:115-116): filters 3, n_blocks 5, ct 5 -> 6.4 M parameters
        gridmax = {"what": "train step at the largest tuning-grid point: filters=3, n_blocks=5, ct_kernel=5, 64x64, batch 16 "
                           "(tune_GEFS_full.py:88-89), device-resident batches, CUDA events", "points": {}}
        for Cg in (1, 3):
            gcfg = dict(H=64, W=64, Cin=Cg, filters=3, n_blocks=5, ct_kernel=5)
            xg, yg, _ = synth_dataset(B * 8, 64, 64, Cg, seed=90 + Cg)
            dxg_, dyg_ = DeviceBuffer.from_array(xg, st), DeviceBuffer.from_array(yg, st)
            fl_g, by_g = algorithmic_work(gcfg)
            pt = {}
            for prec in ("fp32_ffma", "fp32", "tf32"):
                # fp32 = the library's fp32 path (thick layers on the tensor cores with the error-compensated 3xTF32 split, fp32
                # parity); fp32_ffma = the same with S2S_TC3_FP32=0 (CUDA-core kernels only); tf32 = single-pass tensor-core mode
                if prec == "fp32_ffma":
                    os.environ["S2S_TC3_FP32"] = "0"
                mg_ = s2s_model.Model((64, 64, Cg), filters=3, n_blocks=5, ct_kernel=5, max_batch=B, precision="fp32" if prec == "fp32_ffma" else prec)
                os.environ.pop("S2S_TC3_FP32", None)
                mg_.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
                def gstep(i):
                    j = (i % 7) * B
                    call("s2s_unet_train_step", mg_._h, C.c_void_p(dxg_.ptr + j * xg[0].nbytes), C.c_void_p(dyg_.ptr + j * yg[0].nbytes),
                         None, B, None, mg_.sp)
                for i in range(5):
                    gstep(i)
                mg_.stream.synchronize()
                a0, a1 = Event(), Event()
                a0.record(mg_.stream)
                KG = 30
                for i in range(KG):
                    gstep(5 + i)
                a1.record(mg_.stream)
                mg_.stream.synchronize()
                gms = a0.elapsed_ms(a1) / KG
                npar = mg_.count_params()
                pt[prec] = {"ms_per_step": gms, "samples_per_s": B / (gms * 1e-3), "params": npar,
                            "tflops": fl_g * B / (gms * 1e-3) / 1e12,
                            "hbm_frac": (by_g * B + 28.0 * npar) / (gms * 1e-3) / 1e9 / hbm_peak}
                if prec == "fp32_ffma":
                    pt[prec]["ffma_frac"] = pt[prec]["tflops"] / ffma_peak
                mg_.close()
            pt["tf32_speedup"] = pt["tf32"]["samples_per_s"] / pt["fp32"]["samples_per_s"]
            pt["tf32_speedup_vs_fp32_ffma"] = pt["tf32"]["samples_per_s"] / pt["fp32_ffma"]["samples_per_s"]
            pt["fp32_3xtf32_speedup_vs_ffma"] = pt["fp32"]["samples_per_s"] / pt["fp32_ffma"]["samples_per_s"]
            if not args.no_cpu_baseline:
                sps_c, steps_c, thr_c, dt_c = cpu_port_throughput(B, budget_s=8.0, max_steps=6, warmup=1, cfg_kw=gcfg)
                pt["cpu_port"] = {"samples_per_s": sps_c, "cores": thr_c, "steps": steps_c}
            gridmax["points"][f"C={Cg}"] = pt
            dxg_.free(), dyg_.free()

    # ---- CPU baseline (reference's CPU path, torch-CPU port) on this box's host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, steps, threads, dt = cpu_port_throughput(B, budget_s=15.0)
        cpu = {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"{steps} train steps of batch {B} in {dt:.1f} s (torch-CPU fp32 restatement, oracle/keras_unet.py)"}

    if rank == 0:
        line = {
            "metric": "U-Net train samples/s (fwd+bwd)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": line_config(B, world, args.dataset_samples),
            "setup": {"bn": ("global batch statistics (sync-BN over peer memory)" if peer is not None and args.sync_bn
                             else "per-replica batch statistics"),
                      "exchange": ("none" if world == 1 else "fused all-reduce+Adam kernel over NVLink peer memory (CUDA IPC)"
                                   if peer is not None else "ncclAllReduce + Adam kernel"),
                      "cuda_graphs": not args.no_graphs, "precision": "fp32 (the reference's)",
                      "device_dataset_mb": dataset_bytes / 1e6},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_sps, "unit": "samples/s", "h2d_bytes_per_step": int(hx[0].nbytes + hy[0].nbytes),
                    "d2h_bytes_per_step": 8,
                    "host_buffers": "pinned host batches prepared before the timed region (the pageable -> pinned staging the "
                                    "reference's pageable NumPy arrays would need is NOT inside it; see value_pageable)",
                    "value_pageable": e2e_pageable_sps,
                    "value_pageable_streamed": e2e_pageable_stream_sps,
                    "value_per_call": e2e_call_sps,
                    "api": ("Model.train_on_batches (s2s_unet_train_steps_host): a stream of steps, each copying its own batch H2D and "
                            "its own {loss, accuracy} D2H inside the timed region, the copy of batch i+1 staged while step i computes; "
                            "value_per_call = one synchronous Model.train_on_batch call per step") if world == 1 else
                           ("PeerDataParallelTrainer.train_on_batches (s2s_unet_dp_train_steps_host) on every rank: a stream of "
                            "data-parallel steps, each copying its own shard H2D and the global {loss, accuracy} D2H, the copy of shard "
                            "i+1 staged while step i computes; value_per_call = one synchronous train_on_batch call per step"
                            if peer is not None else "one synchronous data-parallel train_on_batch call per step on every rank")},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "trial_batching": trial, "large_batch": large, "inference_c5": infer, "skill_c5": skill,
            "c1_epoch": c1, "predict_b32": pred32, "grid_max": gridmax, "dp_parity": dp_parity, "sync_bn": sync_bn_sec, "kernels": table,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if peer is not None:
            peer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
