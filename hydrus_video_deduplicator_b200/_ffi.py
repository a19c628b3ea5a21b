"""ctypes binding of libvpdq_b200.so (C ABI: include/vpdq_b200.h).

There is deliberately no fallback: if the CUDA library has not been built (run
``python -c "import __graft_entry__ as g; g.build()"`` or ``csrc/build.sh``) loading it raises, and
every entry point fails with VPDQ_B200_ERR_CUDA when no device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# (VPDQ_B200_LIB: development only -- time an alternative build of the same sources, tools/variants.sh)
LIB_PATH = Path(os.environ["VPDQ_B200_LIB"]) if os.environ.get("VPDQ_B200_LIB") else _PKG / "libvpdq_b200.so"

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_OVERFLOW = 0, -1, -2, -3, -4, -5
FRAME_DIM = 512
HASH_BYTES = 32
QUALITY_KEEP = 31
DEFAULT_TOLERANCE = 31


class VpdqB200Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libvpdq_b200 error {code}: {message}")
        self.code = code


def build(verbose: bool = False) -> Path:
    """Compile the CUDA sources in csrc/ for sm_100a into libvpdq_b200.so (nvcc; works without a GPU)."""
    import subprocess

    r = subprocess.run(["bash", str(_PKG / "csrc" / "build.sh")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("building libvpdq_b200.so failed:\n" + r.stderr)
    return LIB_PATH


_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p

# name -> (restype, argtypes); one table so tests can check it against include/vpdq_b200.h
PROTOTYPES = {
    "vpdq_b200_last_error": (C.c_char_p, []),
    "vpdq_b200_abi_version": (C.c_int, []),
    "vpdq_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "vpdq_b200_kernel_launches": (C.c_int, [C.POINTER(C.c_uint64)]),
    "vpdq_b200_debug_flags": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "vpdq_b200_debug_force_timeout": (C.c_int, [C.c_int, C.c_int]),
    "vpdq_b200_dct_matrix": (C.c_int, [_f32p]),
    "vpdq_b200_pdq_scratch_bytes": (C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    "vpdq_b200_pdq_hash_frames_dev": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp,
                                                C.c_size_t, _vp]),
    "vpdq_b200_pdq_stages_dev": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                           C.c_size_t, _vp]),
    "vpdq_b200_pdq_jarosz_dev": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]),
    "vpdq_b200_point_resize_dev": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, _vp, _vp]),
    "vpdq_b200_pdq_hash_frames_host": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int, C.c_int, _vp, _vp, C.c_int]),
    "vpdq_b200_hasher_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "vpdq_b200_hasher_push": (C.c_int, [_vp, _vp, C.c_int64]),
    "vpdq_b200_hasher_push_nocopy": (C.c_int, [_vp, _vp, C.c_int64]),
    "vpdq_b200_hasher_consumed": (C.c_int, [_vp, _i64p]),
    "vpdq_b200_service_stats": (C.c_int, [C.c_int, C.c_int, _i64p]),
    "vpdq_b200_hasher_pushed": (C.c_int, [_vp, _i64p]),
    "vpdq_b200_hasher_finish": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _i64p, _vp, _vp]),
    "vpdq_b200_hasher_destroy": (C.c_int, [_vp]),
    "vpdq_b200_hamming_scan_dev": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "vpdq_b200_hamming_scan_multi_dev": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "vpdq_b200_video_match_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, _vp, _vp]),
    "vpdq_b200_hamming_pairs_dev": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, C.c_int64,
                                              _vp, _vp]),
    "vpdq_b200_match_hash_host": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "vpdq_b200_db_create": (C.c_int, [C.c_int, _vp, C.c_int64, _vp, C.c_int64, C.POINTER(_vp)]),
    "vpdq_b200_db_search": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp]),
    "vpdq_b200_db_search_radius": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, C.c_int64, _i64p]),
    "vpdq_b200_db_destroy": (C.c_int, [_vp]),
    "vpdq_b200_search_host": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, C.c_int, _vp, C.c_int]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
            )
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().vpdq_b200_last_error()
        text = msg.decode("utf-8", "replace") if msg else ""
        if rc in (ERR_INVALID, ERR_UNSUPPORTED):
            raise ValueError(f"libvpdq_b200 error {rc}: {text}")
        if rc == ERR_NOMEM:
            raise MemoryError(f"libvpdq_b200 error {rc}: {text}")
        raise VpdqB200Error(rc, text)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().vpdq_b200_device_count(C.byref(n))
    return n.value if rc == OK else 0


def kernel_launches() -> int:
    n = C.c_uint64(0)
    check(lib().vpdq_b200_kernel_launches(C.byref(n)))
    return int(n.value)


def debug_flags(device: int = 0) -> int:
    f = C.c_int(0)
    check(lib().vpdq_b200_debug_flags(int(device), C.byref(f)))
    return int(f.value)


def service_stats(device: int = 0, channels: int = 3) -> dict:
    out = (C.c_int64 * 8)()
    check(lib().vpdq_b200_service_stats(int(device), int(channels), out))
    keys = ("frames_pushed", "upload_calls", "launches", "frames_launched", "push_blocked_ns", "finish_wait_ns", "largest_launch")
    return dict(zip(keys, list(out)[:7]))


def default_device() -> int:
    """One process per GPU: LOCAL_RANK picks the device unless VPDQ_B200_DEVICE overrides it."""
    for key in ("VPDQ_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(key)
        if v is not None and v.strip().isdigit():
            return int(v)
    return 0
