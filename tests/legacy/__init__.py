"""TEST-ONLY binding of tests/legacy/libvpdq_b200_legacy.so: the round-1 PDQ pipelines (v1 line kernels, the
one-frame and the frame-pair tiled fused kernels) kept as independent CUDA implementations to cross-check the
product's systolic kernel against.  Not imported by the product."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

_DIR = Path(__file__).resolve().parent
LIB_PATH = _DIR / "libvpdq_b200_legacy.so"
IMPLS = {"lines": 0, "fused": 1, "fused2": 2}
_lib = None


def build() -> Path:
    r = subprocess.run(["bash", str(_DIR / "build.sh")], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("building libvpdq_b200_legacy.so failed:\n" + r.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        L.legacy_scratch_bytes.restype = C.c_size_t
        L.legacy_scratch_bytes.argtypes = [C.c_longlong]
        L.legacy_debug_flags.argtypes = [C.POINTER(C.c_int)]
        L.legacy_last_error.restype = C.c_char_p
        L.legacy_pdq_stages_dev.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def hash_frames(frames, impl: str):
    """frames: [n, 512, 512, 3] or [n, 512, 512] uint8 CUDA tensor -> (hashes, quality, a64, b16) CUDA tensors"""
    import torch

    frames = frames.contiguous()
    ch = 3 if frames.dim() == 4 else 1
    n = frames.shape[0]
    dev = frames.device
    hashes = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    quality = torch.empty((n,), dtype=torch.int32, device=dev)
    a64 = torch.empty((n, 64, 64), dtype=torch.float32, device=dev)
    b16 = torch.empty((n, 16, 16), dtype=torch.float32, device=dev)
    L = lib()
    scratch = torch.empty(L.legacy_scratch_bytes(n), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.legacy_pdq_stages_dev(IMPLS[impl], frames.data_ptr(), ch, n, hashes.data_ptr(), quality.data_ptr(),
                                     a64.data_ptr(), b16.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
    if rc:
        raise RuntimeError(f"legacy pipeline {impl} failed ({rc}): {L.legacy_last_error().decode()}")
    return hashes, quality, a64, b16


def debug_flags() -> int:
    f = C.c_int(0)
    lib().legacy_debug_flags(C.byref(f))
    return f.value
