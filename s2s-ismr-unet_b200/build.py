"""Build libs2s_unet.so (sm_100a) in-tree with nvcc.

    python s2s-ismr-unet_b200/build.py [--force] [--verbose]

The library is a plain C-ABI shared object (include/s2s_unet.h): no torch, no pybind.  It is
rebuilt only when a source under csrc/ or the public header is newer than the .so.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libs2s_unet.so"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: cannot build libs2s_unet.so (there is no CPU fallback)")
    return cand


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for d in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh"))) + [ROOT / "include" / "s2s_unet.h"]:
        h.update(d.name.encode())
        h.update(d.read_bytes())
    return h.hexdigest()


def _stale() -> bool:
    """Content hash, not mtime: the .so travels to the GPU box in a snapshot that resets mtimes."""
    stamp = LIB.with_suffix(".so.sha256")
    if not LIB.exists() or not stamp.exists():
        return True
    return stamp.read_text().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = [
        _nvcc(), "-O3", "-std=c++17", "-lineinfo", *ARCH_FLAGS,
        "-shared", "-Xcompiler", "-fPIC,-fvisibility=default",
        "-I", str(ROOT / "include"),
        "-o", str(LIB),
        *[str(s) for s in sources()],
    ]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libs2s_unet.so")
    LIB.with_suffix(".so.sha256").write_text(_source_hash())
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
