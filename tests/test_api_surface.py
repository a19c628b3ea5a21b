"""Host-side checks that need no GPU: the C ABI library loads and exports exactly what
include/vpdq_b200.h declares, and the Python surface mirrors the reference's."""
from __future__ import annotations

import inspect
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
from hydrus_video_deduplicator_b200 import _ffi, hashing, search, vpdq
from hydrus_video_deduplicator_b200.vpdqpy import DOWNSCALE_DIMENSIONS, Vpdq, VpdqHash
from hydrus_video_deduplicator_b200.vpdqpy.vpdqpy import point_resize_indices
from tests.conftest import HAS_GPU

ROOT = Path(__file__).resolve().parents[1]


def _header_functions() -> set[str]:
    text = (ROOT / "include" / "vpdq_b200.h").read_text()
    return set(re.findall(r"VPDQ_B200_API\s+[\w\s\*]+?\b(vpdq_b200_\w+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    declared = _header_functions()
    assert len(declared) >= 20
    L = _ffi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/vpdq_b200.h but not exported"
    assert declared == set(_ffi.PROTOTYPES), declared ^ set(_ffi.PROTOTYPES)
    assert L.vpdq_b200_abi_version() == 1


def test_dct_table_is_bit_identical_to_the_oracle():
    out = np.zeros((16, 64), np.float32)
    import ctypes as C

    _ffi.check(_ffi.lib().vpdq_b200_dct_matrix(out.ctypes.data_as(C.POINTER(C.c_float))))
    assert out.tobytes() == oracle.dct_matrix().tobytes()


def test_product_does_not_import_the_oracle():
    pkg = ROOT / "hydrus_video_deduplicator_b200"
    for py in pkg.rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), py
    for cu in (pkg / "csrc").glob("*"):
        if not cu.is_file():
            continue
        assert "oracle/" not in cu.read_text().replace("the oracle", ""), cu


def test_vpdqhash_value_semantics(golden_dir):
    s = (golden_dir / "video_hashes" / "S02_Sintel_1080_10s_1MB.mp4.txt").read_text()
    h = VpdqHash.from_string(s)
    assert VpdqHash.bytesPerPdqHash == 32
    assert len(h) == 8 and len(h.bytes) == 256 and len(h.bytes) % VpdqHash.bytesPerPdqHash == 0
    assert str(h) == s and hashing.encode_phash_to_str(h) == s
    assert hashing.decode_phash_from_str(s) == h and not (hashing.decode_phash_from_str(s) != h)
    assert VpdqHash.from_string(s.upper()) == h
    assert len(VpdqHash()) == 0 and str(VpdqHash()) == ""
    assert h != VpdqHash.from_string(s[64:])
    with pytest.raises(ValueError):
        VpdqHash.from_string(s[:-1])
    with pytest.raises(ValueError):
        VpdqHash(b"\x00" * 31)


def test_reference_surface_names_and_defaults():
    # vpdqpy.py:28-131
    for name in ("get_video_bytes", "match_hash", "frame_extract_pyav", "computeHash", "is_similar"):
        assert isinstance(inspect.getattr_static(Vpdq, name), staticmethod), name
    assert inspect.signature(Vpdq.is_similar).parameters["threshold"].default == 75.0
    assert inspect.signature(Vpdq.match_hash).parameters["distance_tolerance"].default == 31.0
    assert inspect.signature(Vpdq.computeHash).parameters["num_threads"].default == 0
    assert DOWNSCALE_DIMENSIONS == 512
    # hashing.py:14-53
    for name in ("compute_phash", "encode_phash_to_str", "decode_phash_from_str", "get_phash_similarity"):
        assert callable(getattr(hashing, name))
    # hvdaccelerators.vpdq (Appendix B)
    assert list(inspect.signature(vpdq.VideoHasher.__init__).parameters)[1:5] == ["average_fps", "width", "height",
                                                                                 "num_threads"]
    for name in ("hash_frame", "finish"):
        assert callable(getattr(vpdq.VideoHasher, name))
    assert callable(vpdq.matchHash) and callable(vpdq.matchHashBytes)
    # vptree.py:22-31
    assert [search.fix_vpdq_similarity(s) for s in (100.0, 75.0, 50.0, 99.9, 0.0)] == [1, 26, 51, 2, 101]


def test_get_video_bytes_errors(tmp_path):
    with pytest.raises(ValueError):
        Vpdq.get_video_bytes(tmp_path / "missing.mp4")
    with pytest.raises(ValueError):
        Vpdq.get_video_bytes(123)  # type: ignore[arg-type]
    p = tmp_path / "x.bin"
    p.write_bytes(b"abc")
    assert Vpdq.get_video_bytes(p) == b"abc" and Vpdq.get_video_bytes(str(p)) == b"abc"
    assert Vpdq.get_video_bytes(b"xyz") == b"xyz"


def test_point_resize_is_centre_based_nearest():
    for src in (360, 640, 720, 1080, 1920, 512, 64):
        idx = point_resize_indices(src)
        expect = np.floor((np.arange(512) + 0.5) * src / 512).astype(np.int64)
        assert np.abs(idx - expect).max() <= (0 if src in (360, 640, 512, 1080, 1920, 64, 720) else 1), src
        assert idx.min() >= 0 and idx.max() <= src - 1


@pytest.mark.skipif(HAS_GPU, reason="only meaningful where no CUDA device exists")
def test_fails_loudly_without_a_gpu():
    """No CPU fallback: every compute entry point raises when there is no device."""
    with pytest.raises(_ffi.VpdqB200Error):
        vpdq.VideoHasher(1, 512, 512, 0)
    with pytest.raises(_ffi.VpdqB200Error):
        vpdq.matchHashBytes(b"\x00" * 32, b"\x00" * 32, 31)
    with pytest.raises(_ffi.VpdqB200Error):
        search.HashIndex([1], [b"\x00" * 32])
    assert vpdq.matchHashBytes(b"", b"\x00" * 32, 31) == 0.0  # decided on sizes alone, as in the reference


def test_argument_validation_happens_before_any_cuda_call():
    with pytest.raises(ValueError):
        vpdq.VideoHasher(1, 640, 480, 0)  # only 512x512 is supported (vpdqpy.py:23,113)
    with pytest.raises(ValueError):
        vpdq.matchHashBytes(b"\x00" * 31, b"\x00" * 32, 31)
