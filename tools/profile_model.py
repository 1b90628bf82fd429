"""Per-kernel-tag time table of one train step (eager replay under the library's per-launch event profiler).
    python tools/profile_model.py --filters 3 --n-blocks 5 --ct 5 --cin 1 --batch 16 [--precision tf32] [--hw 64]"""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from s2s_ismr_unet_b200 import _lib, model as s2s_model  # noqa: E402
from s2s_ismr_unet_b200.keras_api.optimizers import Adam  # noqa: E402
from s2s_ismr_unet_b200.runtime import DeviceBuffer, Event  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--filters", type=int, default=2)
ap.add_argument("--n-blocks", type=int, default=3)
ap.add_argument("--ct", type=int, default=3)
ap.add_argument("--cin", type=int, default=3)
ap.add_argument("--hw", type=int, default=64)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
call = _lib.call
m = s2s_model.Model((a.hw, a.hw, a.cin), filters=a.filters, n_blocks=a.n_blocks, ct_kernel=a.ct, max_batch=a.batch, precision=a.precision)
m.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy")
rng = np.random.default_rng(0)
x = rng.gamma(2.0, 3.0, size=(a.batch, a.hw, a.hw, a.cin)).astype(np.float32)
y = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=(a.batch, a.hw, a.hw))]
dx, dy = DeviceBuffer.from_array(x, m.stream), DeviceBuffer.from_array(y, m.stream)
step = lambda: call("s2s_unet_train_step", m._h, C.c_void_p(dx.ptr), C.c_void_p(dy.ptr), None, a.batch, None, m.sp)
for _ in range(5):
    step()
m.stream.synchronize()
e0, e1 = Event(), Event()
e0.record(m.stream)
for _ in range(a.steps):
    step()
e1.record(m.stream)
m.stream.synchronize()
ms = e0.elapsed_ms(e1) / a.steps
print(f"{a}: graph replay {ms * 1e3:.1f} us/step = {a.batch / ms * 1e3:.0f} samples/s, params {m.count_params()}")
call("s2s_prof_enable", 1)
for _ in range(3):
    step()
m.stream.synchronize()
buf = C.create_string_buffer(1 << 16)
call("s2s_prof_report", buf, C.c_size_t(len(buf)))
call("s2s_prof_enable", 0)
rows = []
for line in buf.value.decode().strip().splitlines():
    tag, n, tms, by, fl = line.split(",")
    rows.append((float(tms) / 3, tag, int(n) // 3, float(by) / 3, float(fl) / 3))
tot = sum(r[0] for r in rows)
for tms, tag, n, by, fl in sorted(rows, reverse=True):
    print(f"  {tag:22s} x{n:3d} {tms * 1e3:9.1f} us ({100 * tms / tot:4.1f} %)  {fl / tms / 1e9 if tms else 0:7.2f} TFLOP/s {by / tms / 1e6 if tms else 0:8.1f} GB/s")
print(f"  sum {tot * 1e3:.1f} us")
