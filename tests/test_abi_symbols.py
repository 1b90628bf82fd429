"""The C-ABI library loads here (no GPU) and exports every symbol include/s2s_unet.h declares; the
ctypes struct layouts match the C compiler's.  No compute call is made."""
import ctypes as C
import subprocess
import textwrap

from s2s_ismr_unet_b200 import _lib
from conftest import ROOT


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    protos = _lib.parse_header()
    names = [n for n, _, _ in protos]
    assert len(names) >= 50 and len(set(names)) == len(names)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in s2s_unet.h but not exported: {missing}"
    assert lib.s2s_version() == 2          # S2S_ABI_VERSION: 2 added s2s_unet_cfg.act and the tf32 precision


def test_error_channel_without_gpu_is_loud():
    """No CPU fallback: with no device the model constructor fails with a CUDA error, it does not compute."""
    from s2s_ismr_unet_b200.runtime import device_count
    if device_count() > 0:
        return
    import pytest
    from s2s_ismr_unet_b200.model import Model
    with pytest.raises(_lib.S2SError):
        Model((32, 32, 1))


def test_invalid_arguments_are_rejected_before_any_cuda_call():
    lib = _lib.load()
    cfg = _lib.UnetCfg(24, 24, 1, 2, 4, 3, 0, 1, 0, 16, 1e-3, 0.99, 0)      # 24 not divisible by 2^4
    h = C.c_void_p()
    rc = lib.s2s_unet_create(C.byref(cfg), C.byref(h))
    assert rc == -1 and b"not divisible" in lib.s2s_last_error()
    cfg = _lib.UnetCfg(32, 32, 1, 2, 3, 4, 0, 1, 0, 16, 1e-3, 0.99, 0)      # ct_kernel 4 unsupported
    assert lib.s2s_unet_create(C.byref(cfg), C.byref(h)) == -1 and b"ct_kernel" in lib.s2s_last_error()


def test_struct_layouts_match_the_c_compiler(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text(textwrap.dedent('''
        #include <stdio.h>
        #include <stddef.h>
        #include "s2s_unet.h"
        int main(void) {
            printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(s2s_unet_cfg), sizeof(s2s_tensor_desc), sizeof(s2s_adam_cfg),
                   offsetof(s2s_tensor_desc, offset), offsetof(s2s_unet_cfg, bn_eps), offsetof(s2s_tensor_desc, shape));
            return 0;
        }'''))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.UnetCfg), C.sizeof(_lib.TensorDesc), C.sizeof(_lib.AdamCfg), _lib.TensorDesc.offset.offset,
            _lib.UnetCfg.bn_eps.offset, _lib.TensorDesc.shape.offset]
    assert got == want
