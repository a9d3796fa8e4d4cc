"""Host emulation of the CUDA solver source (TEST INFRASTRUCTURE): builds tests/emu/emu_driver.cpp -- which compiles
dex_retargeting_b200/csrc/dexr_kernels.cuh itself through tests/emu/warp_shim.h -- with g++ and calls it through ctypes.
Used by tests/test_solver_host_emulation.py; never by the product (the product path is the CUDA library only)."""
import ctypes as C
import subprocess
from functools import lru_cache
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
EMU = ROOT / "tests" / "emu"
OUT = EMU / "_build"
SOURCES = [EMU / "emu_driver.cpp", EMU / "warp_shim.h", ROOT / "dex_retargeting_b200" / "csrc" / "dexr_kernels.cuh",
           ROOT / "include" / "dexr.h"]


@lru_cache(maxsize=None)
def load(defines: tuple = ()):
    """defines: compile-time experiment switches, e.g. ("DEXR_EXP_MERGEDRES",)."""
    OUT.mkdir(exist_ok=True)
    tag = "_".join(d.replace("DEXR_EXP_", "").lower() for d in defines) or "default"
    so = OUT / f"libdexr_emu_{tag}.so"
    if not so.exists() or any(so.stat().st_mtime < p.stat().st_mtime for p in SOURCES):
        # -O0: the rendezvous protocol compares the call sites of the lanes, and optimising host compilers duplicate calls
        # (jump threading) -- they do not know that a warp collective is convergent
        cmd = ["g++", "-O0", "-std=c++17", "-fPIC", "-shared", f"-I{EMU / 'stub'}", *[f"-D{d}" for d in defines],
               "-o", str(so), str(EMU / "emu_driver.cpp")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("g++ failed building the host emulation:\n" + res.stderr[-4000:])
    lib = C.CDLL(str(so))
    lib.emu_solve_frames.restype = C.c_int
    lib.emu_rounds.restype = C.c_longlong
    return lib


def solve_frames(opt, last_qpos, keypoints=None, ref_value=None, fixed_qpos=None, projected=None, defines=(), use_arrow=True,
                 clip_init=False):
    """Emulated dexr_solve_frames for an Optimizer of the host mirror.  Returns (qpos [B,n], status [B], cost [B])."""
    from dex_retargeting_b200 import _native as N

    lib = load(tuple(defines))
    table, prm = opt.build_table(), opt.params(clip_init=clip_init)
    B, n = last_qpos.shape[0], table.n_var

    def f32(a):
        return None if a is None else np.ascontiguousarray(a, dtype=np.float32)

    kp, ref, last, fixed = f32(keypoints), f32(ref_value), f32(last_qpos), f32(fixed_qpos)
    assert (kp is None) != (ref is None)
    out = np.full((B, n), np.nan, np.float32)
    status = np.zeros(B, np.int32)
    cost = np.zeros(B, np.float32)
    err = C.create_string_buffer(600)

    def ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    rc = lib.emu_solve_frames(C.byref(table), C.byref(prm), C.c_int(int(use_arrow)), ptr(kp), ptr(ref), ptr(last), ptr(fixed),
                              ptr(projected), C.c_longlong(B), ptr(out), None, ptr(status), ptr(cost), err, C.c_int(600))
    if rc != 0:
        raise RuntimeError(f"host emulation failed ({rc}): {err.value.decode()}")
    return out, status, cost
