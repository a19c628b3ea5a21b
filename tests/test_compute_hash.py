"""Vpdq.computeHash / hashing.compute_phash end to end (vpdqpy.py:104-119): container decode on the host (OpenCV
here, PyAV when installed), every round(fps)-th frame, POINT resize to 512x512, CUDA hash, quality filter.  The
clip is written by the test itself (lossless FFV1, so decode is deterministic) -- the reference's own clips cannot
travel to the GPU box."""
from __future__ import annotations

import numpy as np
import pytest

import oracle
from hydrus_video_deduplicator_b200 import hashing
from hydrus_video_deduplicator_b200.vpdqpy import Vpdq
from hydrus_video_deduplicator_b200.vpdqpy.vpdqpy import point_resize_rgb

cv2 = pytest.importorskip("cv2")


def write_clip(path, n_frames=50, size=(320, 240), fps=10.0):
    w = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"FFV1"), fps, size)
    assert w.isOpened()
    rng = np.random.default_rng(4)
    base = rng.integers(0, 256, (size[1] // 8, size[0] // 8, 3), dtype=np.uint8)
    for k in range(n_frames):
        f = np.roll(np.repeat(np.repeat(base, 8, 0), 8, 1), 3 * k, axis=1)
        if k >= 40:
            f = np.zeros_like(f)  # a black tail: sampled frames 40.. are dropped by the quality filter
        w.write(np.ascontiguousarray(f))
    w.release()


def decoded_reference(path):
    cap = cv2.VideoCapture(str(path))
    step = round(cap.get(cv2.CAP_PROP_FPS))
    frames, idx = [], 0
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        if idx % step == 0:
            frames.append(point_resize_rgb(bgr[:, :, ::-1]))
        idx += 1
    return np.stack(frames)


def test_frame_extraction_samples_every_round_fps_frame(tmp_path):
    clip = tmp_path / "clip.mkv"
    write_clip(clip)
    frames = list(Vpdq.frame_extract_cv2(clip.read_bytes()))
    assert len(frames) == 5 and all(len(f) == 512 * 512 * 3 for f in frames)  # frames 0, 10, 20, 30, 40
    assert b"".join(frames) == decoded_reference(clip).tobytes()
    with pytest.raises(ValueError):
        list(Vpdq.frame_extract_cv2(b"this is not a video"))


@pytest.mark.gpu
def test_compute_phash_matches_oracle(tmp_path):
    clip = tmp_path / "clip.mkv"
    write_clip(clip)
    ref = oracle.video_hash(decoded_reference(clip))
    for source in (clip, str(clip), clip.read_bytes()):
        phash = hashing.compute_phash(source)
        assert phash.bytes == ref
    assert len(phash) == 4  # 5 sampled frames, the black one dropped
    assert hashing.decode_phash_from_str(hashing.encode_phash_to_str(phash)) == phash
    similar, sim = Vpdq.is_similar(phash, phash)
    assert similar and sim == 100.0 and hashing.get_phash_similarity(phash, phash) == 100.0


@pytest.mark.gpu
def test_decode_feed_pool_matches_serial(tmp_path):
    """Vpdq.computeHashes: several clips decoded by a host thread pool, every hasher feeding the
    device's one submission service -> the same hashes as one computeHash after another, in input order."""
    clips = []
    for k in range(6):
        clip = tmp_path / f"clip{k}.mkv"
        write_clip(clip, n_frames=30 + 10 * k, size=(320 + 16 * k, 240))
        clips.append(clip)
    serial = [hashing.compute_phash(c) for c in clips]
    pooled = hashing.compute_phashes(clips, num_threads=4)
    assert [p.bytes for p in pooled] == [s.bytes for s in serial]
    assert [p.bytes for p in pooled] == [oracle.video_hash(decoded_reference(c)) for c in clips]
