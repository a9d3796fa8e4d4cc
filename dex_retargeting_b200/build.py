"""In-tree build of libdexr.so (nvcc, sm_100a only).  `python -m dex_retargeting_b200.build [--force]`"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc" / "dexr.cu"
DEPS = [SRC, PKG / "csrc" / "dexr_kernels.cuh", PKG.parent / "include" / "dexr.h"]
OUT = PKG / "libdexr.so"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xptxas", "-v", "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def build_library(force: bool = False, verbose: bool = True) -> Path:
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in DEPS):
        return OUT
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", str(OUT), str(SRC)]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = (res.stdout or "") + (res.stderr or "")
    (PKG / "csrc" / "build.log").write_text(log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libdexr.so")
    if verbose:
        for line in log.splitlines():
            if "registers" in line or "spill" in line:
                print(line)
    return OUT


if __name__ == "__main__":
    build_library(force="--force" in sys.argv)
