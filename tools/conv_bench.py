"""Per-shape timing of the conv3x3 forward operator through the C ABI (CUDA events, L2-cold via rotating buffers).
    python tools/conv_bench.py            # env S2S_GCONVC=0 / S2S_BIGTILE_MIN_CTAS=n select the kernel"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from s2s_ismr_unet_b200._lib import call  # noqa: E402
from s2s_ismr_unet_b200.runtime import DeviceBuffer, Event, Stream  # noqa: E402

SHAPES = [(128, 64, 64, 8, 8), (128, 64, 64, 16, 8), (128, 32, 32, 8, 16), (128, 32, 32, 16, 16), (128, 32, 32, 32, 16),
          (64, 256, 256, 8, 8), (64, 128, 128, 16, 16), (64, 64, 64, 32, 32), (16, 64, 64, 8, 8)]
st = Stream()
rng = np.random.default_rng(0)
for (N, H, W, Ci, Co) in SHAPES:
    nx, ny = N * H * W * Ci, N * H * W * Co
    R = max(2, int(300e6 // (4 * (nx + ny))))          # rotate over > 2 x L2 worth of buffers
    xs = [DeviceBuffer.from_array(rng.normal(size=nx).astype(np.float32), st) for _ in range(R)]
    ys = [DeviceBuffer(4 * ny) for _ in range(R)]
    w = DeviceBuffer.from_array((rng.normal(size=9 * Ci * Co) / np.sqrt(9 * Ci)).astype(np.float32), st)
    b = DeviceBuffer.from_array(np.zeros(Co, np.float32), st)
    run = lambda i: call("s2s_op_conv3x3_fwd", C.c_void_p(xs[i % R].ptr), C.c_void_p(w.ptr), C.c_void_p(b.ptr), C.c_void_p(ys[i % R].ptr),
                         N, H, W, Ci, Co, 1, C.c_void_p(st.ptr))
    for i in range(3):
        run(i)
    st.synchronize()
    e0, e1 = Event(), Event()
    K = 20
    e0.record(st)
    for i in range(K):
        run(i)
    e1.record(st)
    st.synchronize()
    us = 1e3 * e0.elapsed_ms(e1) / K
    fl = 18.0 * Ci * Co * N * H * W
    by = 4.0 * (nx + ny)
    print(f"{N:4d}x{H}x{W} {Ci:3d}->{Co:3d}: {us:8.1f} us  {fl / us / 1e6:6.1f} TFLOP/s  {by / us / 1e3:7.0f} GB/s", flush=True)
    for bf in xs + ys + [w, b]:
        bf.free()
