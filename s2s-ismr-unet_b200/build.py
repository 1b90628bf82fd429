"""Build libs2s_unet.so (sm_100a) in-tree with nvcc.

    python s2s-ismr-unet_b200/build.py [--force] [--verbose]

The library is a plain C-ABI shared object (include/s2s_unet.h): no torch, no pybind.  Every .cu under
csrc/ is one translation unit, compiled in parallel and cached by content hash.
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
OBJ = PKG / "lib" / "obj"
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found: cannot build libs2s_unet.so (there is no CPU fallback)")
    return cand


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps(src: Path, seen: "set[Path] | None" = None) -> "set[Path]":
    """Transitive #include "..." closure of one source file (csrc/ headers and the public header)."""
    import re
    seen = set() if seen is None else seen
    for inc in re.findall(r'#include\s+"([^"]+)"', src.read_text()):
        d = (src.parent / inc).resolve()
        if d.exists() and d not in seen:
            seen.add(d)
            _deps(d, seen)
    return seen


def _unit_hash(src: Path) -> str:
    """Content hash (not mtime: the .so travels to the GPU box in a snapshot that resets mtimes) of one
    translation unit: its .cu plus every header it includes, transitively."""
    import hashlib
    h = hashlib.sha256()
    for d in [src] + sorted(_deps(src)):
        h.update(d.name.encode())
        h.update(d.read_bytes())
    h.update(" ".join(ARCH_FLAGS).encode())
    return h.hexdigest()


def _source_hash(unit_hashes: "list[str] | None" = None) -> str:
    import hashlib
    h = hashlib.sha256()
    for u in (unit_hashes if unit_hashes is not None else [_unit_hash(s) for s in sources()]):
        h.update(u.encode())
    return h.hexdigest()


def _stale() -> bool:
    stamp = LIB.with_suffix(".so.sha256")
    if not LIB.exists() or not stamp.exists():
        return True
    return stamp.read_text().strip() != _source_hash()


def _compile_unit(src: Path, verbose: bool) -> "tuple[Path, str]":
    """Returns the object and the content hash it was compiled FROM (taken before nvcc starts: a source edited while the
    compiler runs must leave the library stale, not stamped with the hash of text it never saw)."""
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".o.sha256")
    want = _unit_hash(src)
    if obj.exists() and stamp.exists() and stamp.read_text().strip() == want:
        return obj, want
    cmd = [_nvcc(), "-O3", "-std=c++17", "-lineinfo", *ARCH_FLAGS, "-Xcompiler", "-fPIC,-fvisibility=default",
           "-I", str(ROOT / "include"), "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}")
    stamp.write_text(want)
    return obj, want


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per .cu (compiled in parallel, cached by content hash), linked into libs2s_unet.so."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    LIB.parent.mkdir(parents=True, exist_ok=True)
    OBJ.mkdir(parents=True, exist_ok=True)
    if force:
        for f in OBJ.glob("*.sha256"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        units = list(ex.map(lambda s: _compile_unit(s, verbose), sources()))
    objs = [o for o, _ in units]
    cmd = [_nvcc(), *ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB), *[str(o) for o in objs]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        sys.stderr.write(r.stderr)
        raise RuntimeError("nvcc failed linking libs2s_unet.so")
    LIB.with_suffix(".so.sha256").write_text(_source_hash([u for _, u in units]))
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
