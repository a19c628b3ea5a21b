"""CPU check of the fused PDQ kernel's schedule (kx_fused_p123): the emulator in tests/emu compiles the very
header the CUDA kernel is built from (csrc/pdq_fused_core.h) and executes it step by step; its output,
finished with a plain numpy column pass 2 + decimation, must equal the oracle's 64x64 plane bit for bit --
for every grid size (frames per CTA 0, 1, many; CTA ranges starting mid-batch)."""
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
    lib.emu_fused_p3t.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
    return lib


def colpass2_decimate(p3t: np.ndarray) -> np.ndarray:
    """p3t [64 cols][512 rows] f32 -> A [64][64]: running-sum column pass (window 4), outputs 8m+4."""
    s = np.zeros(64, np.float32)
    a = np.zeros((64, 64), np.float32)
    quarter = np.float32(0.25)
    for r in range(512):
        s = (s + p3t[:, r]).astype(np.float32)
        if r >= 4:
            s = (s - p3t[:, r - 4]).astype(np.float32)
        o = r - 2
        if o >= 0 and o % 8 == 4:
            a[o // 8, :] = s * quarter
    return a


@pytest.mark.parametrize("n_frames,grid", [(1, 1), (3, 1), (5, 2), (4, 7)])
def test_emulated_schedule_is_bit_exact(emu, n_frames, grid):
    frames = synth.synth_frames(n_frames, seed=31 + n_frames)
    p3t = np.full((n_frames, 64, 512), np.nan, np.float32)
    errors = emu.emu_fused_p3t(frames.ctypes.data_as(C.c_void_p), n_frames, grid, p3t.ctypes.data_as(C.c_void_p))
    assert errors == 0
    assert not np.isnan(p3t).any(), "some decimated outputs were never written"
    for f in range(n_frames):
        _, _, a_ref, _ = oracle.pdq_stages(frames[f])
        a = colpass2_decimate(p3t[f])
        assert a.tobytes() == a_ref.tobytes(), f"frame {f}: decimated plane differs from the oracle"
