"""CPU check of the frame submission service behind every VideoHasher (csrc/hash_service.h): the harness in
tests/emu compiles the very template the CUDA library instantiates, with a mock device, and drives it from several
threads.  What must hold for any interleaving: results come back per hasher in push order (the round-1 ring drained
out of order at exact multiples of its batch size), frames of concurrent hashers share launches without
cross-talk, a full ring blocks pushes instead of dropping them, the consumed watermark reaches the push count by the
time finish returns, a handle is reusable after finish, and a device error surfaces as an error -- never as a hang."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import pytest

EMU_DIR = Path(__file__).resolve().parent / "emu"


@pytest.fixture(scope="module")
def svc():
    so = EMU_DIR / "libhash_service_emu.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", str(so),
                    str(EMU_DIR / "hash_service_emu.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emu_service_run.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    lib.emu_service_failure.argtypes = [C.c_int]
    return lib


def _run(lib, arena, workers, threads, videos, counts, frame_bytes=4096):
    arr = (C.c_int * len(counts))(*counts)
    stats = (C.c_longlong * 3)()
    errors = lib.emu_service_run(arena, workers, threads, videos, arr, len(counts), frame_bytes, stats)
    return errors, list(stats)


@pytest.mark.parametrize("arena,workers,threads,videos,counts", [
    (8, 2, 1, 11, [0, 1, 7, 8, 9, 16, 31, 32, 33, 96, 128]),          # exact multiples of the ring, empty videos
    (8, 3, 4, 20, [10, 0, 1, 8, 16, 32, 96, 128, 300, 5]),            # tiny ring, 4 caller threads: back-pressure
    (32, 1, 2, 12, [96, 128, 64, 33]),                                # a single copy worker
    (64, 4, 3, 10, [300, 10, 96, 128, 64, 1]),
    (256, 4, 2, 4, [300, 1000]),                                      # the product's ring size
])
def test_results_in_push_order_for_any_interleaving(svc, arena, workers, threads, videos, counts):
    errors, (launches, frames, biggest) = _run(svc, arena, workers, threads, videos, counts)
    assert errors == 0
    expect = sum(counts[(t * videos + v) % len(counts)] for t in range(threads) for v in range(videos))
    assert frames == expect                      # every frame went through exactly one launch
    assert biggest <= arena and launches <= max(1, frames)


def test_concurrent_hashers_share_launches(svc):
    """4 threads x 10-frame videos: launches must carry frames of several videos (fewer launches than frames)."""
    errors, (launches, frames, _) = _run(svc, 64, 4, 4, 200, [10], frame_bytes=65536)
    assert errors == 0 and frames == 8000
    assert launches < frames


@pytest.mark.parametrize("fail_after", [0, 2, 4])
def test_device_error_surfaces_and_never_hangs(svc, fail_after):
    got = svc.emu_service_failure(fail_after)
    assert got & 1, "wait_all must report the device error"
    assert got & 4, "the service must stay broken (fail loudly) after a device error"
