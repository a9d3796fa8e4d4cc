"""C-ABI library: loads, exports every symbol include/dexr.h declares, struct layouts agree with the
binding, argument validation works without a GPU (no compute calls here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from dex_retargeting_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "dexr.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dexr_[a-z_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = N.load()
    names = declared_symbols()
    assert set(names) == set(N.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_struct_layouts():
    lib = N.load()
    assert lib.dexr_version() == 1
    assert lib.dexr_table_sizeof() == C.sizeof(N.DexrTable) == 8192
    assert lib.dexr_params_sizeof() == C.sizeof(N.DexrParams) == 52
    assert lib.dexr_frames_sizeof() == C.sizeof(N.DexrFrames) == 80
    assert lib.dexr_sequences_sizeof() == C.sizeof(N.DexrSequences) == 72
    p = N.default_params()
    assert (p.huber_delta, p.norm_delta, p.max_iters, p.clip_init) == (pytest.approx(0.02), pytest.approx(4e-3), 64, 0)
    assert p.tol == pytest.approx(1e-5) and p.lambda0 == pytest.approx(1e-2) and p.lp_alpha < 0


def test_header_constants_match_binding():
    text = (ROOT / "include" / "dexr.h").read_text()
    consts = dict(re.findall(r"#define\s+(DEXR_[A-Z_]+)\s+\(?(-?\d+)\)?", text))
    assert int(consts["DEXR_MAX_LANES"]) == N.MAX_LANES
    assert int(consts["DEXR_MAX_LINKS"]) == N.MAX_LINKS
    assert int(consts["DEXR_MAX_RES"]) == N.MAX_RES
    assert int(consts["DEXR_MAX_GROUP"]) == N.MAX_GROUP
    assert int(consts["DEXR_MAX_LINKS_PER_LANE"]) == N.MAX_LINKS_PER_LANE
    assert int(consts["DEXR_NUM_KEYPOINTS"]) == N.NUM_KEYPOINTS
    assert (int(consts["DEXR_LOSS_POSITION"]), int(consts["DEXR_LOSS_VECTOR"]), int(consts["DEXR_LOSS_DEXPILOT"])) == (0, 1, 2)


def test_invalid_arguments_are_rejected_without_gpu():
    lib = N.load()
    h = C.c_void_p()
    assert lib.dexr_robot_create(None, 0, C.byref(h)) == -1
    assert b"null" in lib.dexr_last_error()
    t = N.DexrTable()  # zeroed: bad magic
    assert lib.dexr_robot_create(C.byref(t), 0, C.byref(h)) == -1
    assert b"magic" in lib.dexr_last_error()
    t.magic, t.nbytes = N.TABLE_MAGIC, 17
    assert lib.dexr_robot_create(C.byref(t), 0, C.byref(h)) == -1
    assert b"size" in lib.dexr_last_error()
    t.nbytes = C.sizeof(N.DexrTable)
    t.dof = 40
    assert lib.dexr_robot_create(C.byref(t), 0, C.byref(h)) == -1
    assert b"dof" in lib.dexr_last_error()
    p = N.default_params()
    io = N.DexrFrames()
    assert lib.dexr_solve_frames(None, C.byref(p), C.byref(io), 1, None) == -1
    assert lib.dexr_solve_sequences(None, C.byref(p), None, 1, 1, None) == -1
    assert lib.dexr_solve_frames_host(None, C.byref(p), C.byref(io), 1) == -1
    lib.dexr_robot_destroy(None)  # no-op


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("DEXR_LIBRARY", str(tmp_path / "nope.so"))
    monkeypatch.setattr(N, "_LIB", None)
    with pytest.raises(N.DexrError, match="no CPU fallback"):
        N.load()
    monkeypatch.delenv("DEXR_LIBRARY")
    monkeypatch.setattr(N, "_LIB", None)
    N.load()


def test_product_never_imports_oracle():
    for py in (ROOT / "dex_retargeting_b200").glob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), py
