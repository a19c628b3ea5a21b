"""CPU check of the PDQ kernels' schedules: the emulators in tests/emu compile the very headers the CUDA kernels are
built from (the product's csrc/pdq_systolic_core.h; the test-only tests/legacy/pdq_fused*_core.h of the round-1
tiled kernels) and execute them step by step; the output must equal the oracle's decimated 64x64 plane bit for bit
-- for every work split (frames per warp / CTA 0, 1, many; ranges starting mid-batch).  Also checks the branch-free
exact division by 3 the kernels use."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from tests import synth

EMU_DIR = Path(__file__).resolve().parent / "emu"


@pytest.fixture(scope="module")
def emu():
    so = EMU_DIR / "libpdq_fused_emu.so"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(EMU_DIR / "pdq_fused_emu.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emu_fused_a64.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
    return lib


@pytest.mark.parametrize("n_frames,grid", [(1, 1), (3, 1), (5, 2), (4, 7)])
def test_emulated_schedule_is_bit_exact(emu, n_frames, grid):
    frames = synth.synth_frames(n_frames, seed=31 + n_frames)
    a64 = np.full((n_frames, 64, 64), np.nan, np.float32)
    errors = emu.emu_fused_a64(frames.ctypes.data_as(C.c_void_p), n_frames, grid, a64.ctypes.data_as(C.c_void_p))
    assert errors == 0
    assert not np.isnan(a64).any(), "some decimated outputs were never written"
    for f in range(n_frames):
        _, _, a_ref, _ = oracle.pdq_stages(frames[f])
        assert a64[f].tobytes() == a_ref.tobytes(), f"frame {f}: decimated plane differs from the oracle"


@pytest.fixture(scope="module")
def emu2():
    so = EMU_DIR / "libpdq_fused2_emu.so"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(EMU_DIR / "pdq_fused2_emu.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emu_fused2_a64.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int]
    return lib


@pytest.mark.parametrize("n_frames,grid,channels", [(1, 1, 3), (2, 1, 3), (3, 1, 3), (6, 1, 3), (5, 2, 3), (4, 7, 3),
                                                   (9, 2, 3), (1, 1, 1), (5, 2, 1), (6, 1, 1)])
def test_emulated_pair_schedule_is_bit_exact(emu2, n_frames, grid, channels):
    """kx_fused_jarosz2 (two frames per lane, 8 main warps + the P4 warp, deferred power-of-two scaling): odd and
    even frame counts per CTA (half B one frame shorter), a single frame (half B empty), empty CTAs; RGB24 and
    8-bit gray input."""
    frames = synth.synth_frames(n_frames, seed=57 + n_frames, channels=channels)
    a64 = np.full((n_frames, 64, 64), np.nan, np.float32)
    errors = emu2.emu_fused2_a64(frames.ctypes.data_as(C.c_void_p), n_frames, grid, a64.ctypes.data_as(C.c_void_p),
                                 channels)
    assert errors == 0
    assert not np.isnan(a64).any(), "some decimated outputs were never written (or were fed poison)"
    for f in range(n_frames):
        rgb = frames[f] if channels == 3 else np.repeat(frames[f][..., None], 3, axis=2)  # gray == R = G = B
        _, _, a_ref, _ = oracle.pdq_stages(rgb)
        assert a64[f].tobytes() == a_ref.tobytes(), f"frame {f}: decimated plane differs from the oracle"


@pytest.fixture(scope="module")
def emu_sys():
    so = EMU_DIR / "libpdq_systolic_emu.so"
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(EMU_DIR / "pdq_systolic_emu.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emu_systolic_a64.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int]
    lib.emu_systolic_general_only.argtypes = [C.c_int]
    return lib


@pytest.mark.parametrize("n_frames,n_warps,channels,general_only", [
    (1, 1, 3, False), (2, 1, 3, False), (4, 1, 3, False), (3, 2, 3, False), (5, 8, 3, False), (1, 1, 1, False),
    (3, 2, 1, False), (4, 1, 1, False), (3, 1, 3, True), (2, 3, 1, True)])
def test_emulated_systolic_schedule_is_bit_exact(emu_sys, n_frames, n_warps, channels, general_only):
    """kx_systolic_jarosz (one warp per frame, chain state handed lane to lane, per-group TMA rings): several
    frames streamed back to back through one warp (the zero rows between frames must flush every column chain),
    warps without frames, RGB24 and gray.  The emulator counts a read of a ring slot that is in flight or holds
    another stream row as an error, and checks per lane what a plain iteration takes for granted.  general_only: no
    iteration is plain and every TMA event is per-group boxes (the kernel under VPDQ_B200_SYSTOLIC_3D=0)."""
    frames = synth.synth_frames(n_frames, seed=57 + n_frames, channels=channels)
    a64 = np.full((n_frames, 64, 64), np.nan, np.float32)
    emu_sys.emu_systolic_general_only(int(general_only))
    try:
        errors = emu_sys.emu_systolic_a64(frames.ctypes.data_as(C.c_void_p), n_frames, n_warps,
                                          a64.ctypes.data_as(C.c_void_p), channels)
    finally:
        emu_sys.emu_systolic_general_only(0)
    assert errors == 0
    assert not np.isnan(a64).any(), "some decimated outputs were never written"
    for f in range(n_frames):
        rgb = frames[f] if channels == 3 else np.repeat(frames[f][..., None], 3, axis=2)  # gray == R = G = B
        _, _, a_ref, _ = oracle.pdq_stages(rgb)
        assert a64[f].tobytes() == a_ref.tobytes(), f"frame {f}: decimated plane differs from the oracle"


def test_div3_is_exact(tmp_path):
    exe = tmp_path / "div3_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(EMU_DIR / "div3_check.c"), "-lm"], check=True)
    for d in ("3", "255"):
        out = subprocess.run([str(exe), "11", d], check=True, capture_output=True, text=True).stdout.split()
        assert int(out[0]) > 190_000_000 and int(out[1]) == 0, (d, out)
