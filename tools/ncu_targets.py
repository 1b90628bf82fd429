"""Short single-process workload for `ncu --set full` captures of the tensor-core / CUDA-core convolution kernels
(one launch list per kernel family; see tools/ncu_r2.sh).  Every operator goes through the C ABI."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from s2s_ismr_unet_b200._lib import call  # noqa: E402
from s2s_ismr_unet_b200.runtime import DeviceBuffer, Stream  # noqa: E402

st = Stream()
rng = np.random.default_rng(0)
P = lambda b: C.c_void_p(b.ptr)
sp = C.c_void_p(st.ptr)


def bufs(N, H, W, Ci, Co):
    x = DeviceBuffer.from_array(rng.normal(size=N * H * W * Ci).astype(np.float32), st)
    w = DeviceBuffer.from_array((rng.normal(size=9 * Ci * Co) / np.sqrt(9 * Ci)).astype(np.float32), st)
    b = DeviceBuffer.from_array(np.zeros(Co, np.float32), st)
    y = DeviceBuffer(4 * N * H * W * Co)
    return x, w, b, y


reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
# tcgen05 tf32 conv (tc3conv_kernel): thin training layers at batch 16 / 128, the C5 inference layer, thick layers
for (N, H, W, Ci, Co) in [(16, 64, 64, 8, 8), (128, 64, 64, 8, 8), (128, 32, 32, 16, 16), (128, 16, 16, 32, 32), (64, 256, 256, 8, 8),
                          (16, 64, 64, 96, 96), (16, 16, 16, 192, 192)]:
    x, w, b, y = bufs(N, H, W, Ci, Co)
    for npass in (1, 3):
        for _ in range(reps):
            call("s2s_op_conv3x3_fwd_tf32", P(x), P(w), P(b), P(y), N, H, W, Ci, Co, 1, npass, sp)
    # the CUDA-core kernel on the same shape (gconv_kernel)
    for _ in range(reps):
        call("s2s_op_conv3x3_fwd", P(x), P(w), P(b), P(y), N, H, W, Ci, Co, 1, sp)
    st.synchronize()
    for t in (x, w, b, y):
        t.free()
# tcgen05 bf16 conv of round 1 (tcconv_kernel): thick layers
for (N, H, W, Ci, Co) in [(64, 32, 32, 64, 64), (16, 64, 64, 96, 96), (16, 16, 16, 192, 192)]:
    x, w, b, y = bufs(N, H, W, Ci, Co)
    for _ in range(reps):
        call("s2s_op_conv3x3_fwd_tc", P(x), P(w), P(b), P(y), N, H, W, Ci, Co, 1, sp)
    st.synchronize()
    for t in (x, w, b, y):
        t.free()
print("ncu_targets done")
