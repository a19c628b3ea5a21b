"""GPU parity of the PDQ frame-hash path against the CPU oracle (bit-exact: hashes, quality, and the
fp32 intermediates).  Every call goes through the C ABI (libvpdq_b200.so)."""
from __future__ import annotations

import numpy as np
import pytest

import oracle
from hydrus_video_deduplicator_b200 import vpdq
from hydrus_video_deduplicator_b200.vpdqpy.vpdqpy import point_resize_rgb
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_dev():
    import torch

    return torch, torch.device("cuda", 0)


def _gpu_hash(torch, dev, frames, stages=False):
    from hydrus_video_deduplicator_b200 import device

    out = device.hash_frames(torch.from_numpy(frames).to(dev), stages=stages)
    torch.cuda.synchronize()
    return tuple(o.cpu().numpy() for o in out)


def test_gif_known_answer_on_gpu(golden_dir, torch_dev):
    torch, dev = torch_dev
    native = np.load(golden_dir / "bbb_gif_frames.npz")["frames"]
    frames = np.stack([point_resize_rgb(f) for f in native])
    hashes, quality = _gpu_hash(torch, dev, frames)
    gold = bytes.fromhex((golden_dir / "video_hashes" / "S01_Big_Buck_Bunny_360_10s.gif.txt").read_text().strip())
    assert hashes.tobytes() == gold
    assert (quality == 100).all()


def test_device_point_resize_and_native_hash(golden_dir, torch_dev):
    """8f-2: swscale POINT resize on the device == the host index rule, for several native sizes; and the golden
    GIF clip hashed from its NATIVE 360x640 frames entirely on the device reproduces the reference's hashes."""
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    rng = np.random.default_rng(2)
    for h, w in ((360, 640), (1080, 1920), (64, 48), (512, 512), (719, 1281), (2160, 3840)):
        src = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        got = device.point_resize(torch.from_numpy(src).to(dev)).cpu().numpy()
        assert got.tobytes() == np.stack([point_resize_rgb(f) for f in src]).tobytes(), (h, w)
    native = np.load(golden_dir / "bbb_gif_frames.npz")["frames"]
    hashes, quality = device.hash_native_frames(torch.from_numpy(native).to(dev))
    gold = bytes.fromhex((golden_dir / "video_hashes" / "S01_Big_Buck_Bunny_360_10s.gif.txt").read_text().strip())
    assert hashes.cpu().numpy().tobytes() == gold and (quality.cpu().numpy() == 100).all()


def test_stages_bit_exact_vs_oracle(torch_dev):
    """The 64x64 decimated plane and the 16x16 DCT must match the oracle to the last bit: this is what
    pins the fp32 operation order (no FMA contraction, sequential running sums)."""
    torch, dev = torch_dev
    frames = synth.synth_frames(9, seed=21)
    hashes, quality, a64, b16 = _gpu_hash(torch, dev, frames, stages=True)
    for k in range(len(frames)):
        h, q, a, b = oracle.pdq_stages(frames[k])
        assert a64[k].tobytes() == a.tobytes(), f"decimated plane differs, frame {k}"
        assert b16[k].tobytes() == b.tobytes(), f"DCT differs, frame {k}"
        assert hashes[k].tobytes() == h.tobytes() and quality[k] == q


def test_jarosz_planes_entry_point(torch_dev):
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    frames = synth.synth_frames(5, seed=33)
    planes = device.jarosz_planes(torch.from_numpy(frames).to(dev)).cpu().numpy()
    for k in range(5):
        assert planes[k].tobytes() == oracle.pdq_stages(frames[k])[2].tobytes()


def test_two_hashers_interleaved():
    """Two VideoHashers fed alternately (each owns its stream, ring and scratch): no cross-talk."""
    fa, fb = synth.synth_frames(70, seed=50), synth.synth_frames(45, seed=51)
    ha, hb = vpdq.VideoHasher(1, 512, 512, 0), vpdq.VideoHasher(1, 512, 512, 4)
    for k in range(70):
        ha.hash_frame(fa[k].tobytes())
        if k < 45:
            hb.hash_frame(fb[k].tobytes())
    assert ha.finish().bytes == oracle.video_hash(fa, nthreads=8)
    assert hb.finish().bytes == oracle.video_hash(fb, nthreads=8)
    ha.close()
    hb.close()


@pytest.mark.parametrize("channels", [3, 1])
def test_synthetic_frames_bit_exact(torch_dev, channels):
    torch, dev = torch_dev
    n = 300 if channels == 3 else 90  # > one internal chunk (256) for RGB
    frames = synth.synth_frames(n, seed=7, channels=channels)
    hashes, quality = _gpu_hash(torch, dev, frames)
    ref_h, ref_q = oracle.pdq_hash_frames(frames, nthreads=8)
    bad = np.flatnonzero((hashes != ref_h).any(axis=1) | (quality != ref_q))
    assert bad.size == 0, f"{bad.size} of {n} frames differ, first {bad[:5]}"


def test_degenerate_frames(torch_dev):
    torch, dev = torch_dev
    frames = np.stack([np.zeros((512, 512, 3), np.uint8), np.full((512, 512, 3), 128, np.uint8),
                       np.full((512, 512, 3), 255, np.uint8)])
    frames[2, 100:200, 300:400] = 0
    hashes, quality = _gpu_hash(torch, dev, frames)
    ref_h, ref_q = oracle.pdq_hash_frames(frames)
    assert hashes.tobytes() == ref_h.tobytes() and (quality == ref_q).all()
    assert hashes[0].tobytes() == bytes(32) and quality[0] == 0


def test_video_hasher_streaming_matches_oracle():
    """VideoHasher.hash_frame / finish (vpdqpy.py:113-119): > 3 staged batches so the pinned ring wraps,
    quality filter >= 31 applied in finish()."""
    frames = synth.synth_frames(110, seed=9)
    frames[5] = 0  # a black frame: quality 0, must be dropped
    frames[64] = 77  # flat frame
    hasher = vpdq.VideoHasher(1, 512, 512, 0)
    for k in range(0, 40):
        hasher.hash_frame(frames[k].tobytes())
    hasher.hash_frames(frames[40:])  # batch push, numpy buffer
    phash, all_h, all_q = hasher.finish(return_all=True)
    ref_h, ref_q = oracle.pdq_hash_frames(frames, nthreads=8)
    assert all_h == ref_h.tobytes() and all_q == ref_q.tolist()
    assert phash.bytes == ref_h[ref_q >= 31].tobytes() == oracle.video_hash(frames, nthreads=8)
    assert len(phash) < 110
    # the hasher is reusable after finish()
    hasher.hash_frame(frames[0].tobytes())
    assert hasher.finish().bytes == ref_h[0].tobytes()
    assert len(hasher.finish()) == 0  # nothing pushed -> empty hash (legal, dedup.py:82)
    with pytest.raises(ValueError):
        hasher.hash_frame(b"\x00" * 100)
    hasher.close()


def test_host_batch_call_matches_oracle():
    import ctypes as C

    from hydrus_video_deduplicator_b200 import _ffi

    frames = synth.synth_frames(20, seed=13)
    hashes = np.zeros((20, 32), np.uint8)
    quality = np.zeros(20, np.int32)
    _ffi.check(_ffi.lib().vpdq_b200_pdq_hash_frames_host(frames.ctypes.data_as(C.c_void_p), 3, 20, 512, 512,
                                                         hashes.ctypes.data_as(C.c_void_p),
                                                         quality.ctypes.data_as(C.c_void_p), 0))
    ref_h, ref_q = oracle.pdq_hash_frames(frames, nthreads=8)
    assert hashes.tobytes() == ref_h.tobytes() and (quality == ref_q).all()


def test_large_batch_properties(torch_dev):
    """Size-independent properties at a size the oracle cannot cover quickly: determinism, batch-split
    invariance, and popcount 128 wherever the DCT values are distinct."""
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    g = torch.Generator(device=dev).manual_seed(3)
    frames = torch.randint(0, 256, (1024, 512, 512, 3), dtype=torch.uint8, device=dev, generator=g)
    h1, q1 = device.hash_frames(frames)
    h2, q2 = device.hash_frames(frames)
    ha, qa = device.hash_frames(frames[:333])
    hb, qb = device.hash_frames(frames[333:])
    torch.cuda.synchronize()
    assert torch.equal(h1, h2) and torch.equal(q1, q2)
    assert torch.equal(h1, torch.cat([ha, hb])) and torch.equal(q1, torch.cat([qa, qb]))
    pop = np.unpackbits(h1.cpu().numpy(), axis=1).sum(axis=1)
    assert (pop == 128).all()
    sample = frames[::97].cpu().numpy()
    ref_h, ref_q = oracle.pdq_hash_frames(sample, nthreads=8)
    assert h1[::97].cpu().numpy().tobytes() == ref_h.tobytes() and (q1[::97].cpu().numpy() == ref_q).all()


def test_config1_100k_frames(torch_dev):
    """BASELINE configs[1]: 100k synthetic frames through the device-resident path, in 8192-frame pieces.
    Full-size checks: split invariance (two different batchings give identical bytes), popcount 128 on every
    non-degenerate frame, and an oracle comparison on a strided sample."""
    torch, dev = torch_dev
    from bench import device_frames
    from hydrus_video_deduplicator_b200 import device

    total, piece = 100_000, 8192
    checked = 0
    for p0 in range(0, total, piece):
        n = min(piece, total - p0)
        frames = device_frames(torch, n, dev, seed=p0)
        h1, q1 = device.hash_frames(frames)
        cut = 1 + (p0 // piece) * 37 % (n - 1)
        ha, qa = device.hash_frames(frames[:cut])
        hb, qb = device.hash_frames(frames[cut:])
        assert torch.equal(h1, torch.cat([ha, hb])) and torch.equal(q1, torch.cat([qa, qb]))
        pop = torch.from_numpy(np.unpackbits(h1.cpu().numpy(), axis=1).sum(axis=1))
        assert int((pop != 128).sum()) == 0
        idx = torch.arange(p0 % 7, n, 1021, device=dev)
        ref_h, ref_q = oracle.pdq_hash_frames(frames[idx].cpu().numpy(), nthreads=16)
        assert h1[idx].cpu().numpy().tobytes() == ref_h.tobytes()
        assert (q1[idx].cpu().numpy() == ref_q).all()
        checked += len(idx)
        del frames
    assert checked >= 100


def test_fused_kernel_is_deterministic_under_load(torch_dev):
    """Regression test for a shared-memory WAR race (TMA refill vs in-flight LDS): many frames per CTA, same
    input hashed repeatedly -> identical decimated planes every time."""
    torch, dev = torch_dev
    from bench import device_frames
    from hydrus_video_deduplicator_b200 import device

    frames = device_frames(torch, 8192, dev, seed=123)
    h0, q0, a0, _ = device.hash_frames(frames, stages=True)
    for _ in range(4):
        h, q, a, _ = device.hash_frames(frames, stages=True)
        assert torch.equal(a, a0) and torch.equal(h, h0) and torch.equal(q, q0)
    from hydrus_video_deduplicator_b200 import _ffi

    assert _ffi.debug_flags(0) == 0  # no TMA wait ever timed out


@pytest.mark.parametrize("n_frames,channels", [(1, 3), (2, 3), (3, 3), (147, 3), (149, 3), (297, 3), (1000, 3),
                                               (1, 1), (150, 1), (601, 1)])
def test_three_cuda_pipelines_agree(torch_dev, n_frames, channels):
    """kx_fused_jarosz2 (frame pairs, default), kx_fused_jarosz and the v1 line kernels give identical decimated
    planes, hashes and quality -- for frame counts that leave CTAs with 0, 1, odd and even numbers of frames, RGB24
    and 8-bit gray input (gray: the one-frame fused kernel is RGB-only and falls back to the line kernels) -- and
    the default equals the oracle on a strided sample."""
    torch, dev = torch_dev
    from bench import device_frames
    from hydrus_video_deduplicator_b200 import _ffi, device

    frames = device_frames(torch, n_frames, dev, seed=900 + n_frames)
    if channels == 1:
        frames = frames[..., 1].contiguous()
    out = {}
    try:
        for impl in ("fused2", "fused", "lines"):
            _ffi.set_pdq_impl(impl)
            assert _ffi.get_pdq_impl() == impl
            out[impl] = device.hash_frames(frames, stages=True)
    finally:
        _ffi.set_pdq_impl("fused2")
    torch.cuda.synchronize()
    for impl in ("fused", "lines"):
        for got, want in zip(out[impl][:3], out["fused2"][:3]):
            assert torch.equal(got, want), impl
    idx = list(range(0, n_frames, max(1, n_frames // 12)))
    ref_h, ref_q = oracle.pdq_hash_frames(frames[idx].cpu().numpy(), nthreads=8)
    assert out["fused2"][0][idx].cpu().numpy().tobytes() == ref_h.tobytes()
    assert (out["fused2"][1][idx].cpu().numpy() == ref_q).all()
    assert _ffi.debug_flags(0) == 0
