"""Build libvgen_b200.so (and the oracle's C pieces) in-tree with nvcc for sm_100a.

`python -m vgen_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU; the
resulting .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libvgen_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _stamp(src: Path) -> str:
    h = hashlib.sha1()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    h.update(" ".join(CFLAGS + ARCH).encode())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    OBJ.mkdir(exist_ok=True)
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".stamp")
    want = _stamp(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [NVCC, *CFLAGS, *ARCH, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = OBJ / (src.stem + ".log")
    log.write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for line in (r.stdout + r.stderr).splitlines():
            if "spill" in line and "0 bytes spill stores, 0 bytes spill loads" not in line:
                print(f"[build] {src.name}: {line.strip()}")
    stamp.write_text(want)
    return obj


def build(verbose: bool = True) -> Path:
    srcs = _sources()
    if not srcs:
        raise RuntimeError("no CUDA sources found")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if (not LIB.exists()) or LIB.stat().st_mtime < newest:
        tmp = LIB.with_suffix(".so.tmp")          # link aside, then rename: a reader never sees a half-written library
        cmd = [NVCC, *ARCH, "-shared", "-o", str(tmp), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
    if verbose:
        print(f"[build] {LIB} ({LIB.stat().st_size // 1024} KiB, {len(objs)} objects)")
    return LIB


if __name__ == "__main__":
    build(verbose=True)
    sys.exit(0)
