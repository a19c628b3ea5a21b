"""Mirror of the reference's ``hydrusvideodeduplicator/vpdqpy/vpdqpy.py`` (class ``Vpdq``, lines 28-131)
with the native calls landing in libvpdq_b200.so instead of hvdaccelerators.

Same static methods, argument meaning and exceptions:
    Vpdq.get_video_bytes   vpdqpy.py:30-47     Vpdq.match_hash   vpdqpy.py:50-56
    Vpdq.frame_extract_*   vpdqpy.py:59-101    Vpdq.computeHash  vpdqpy.py:104-119
    Vpdq.is_similar        vpdqpy.py:122-131

Decoding is NOT part of the accelerated path: frames are decoded on the host with PyAV
when it is installed (exactly the reference's code path) and with OpenCV's bundled FFmpeg otherwise (with a
warning: that path is not bit-compatible), sampled every round(fps)-th frame and POINT-resized to 512x512
RGB24 -- then every frame goes through VideoHasher.hash_frame -> CUDA.  Vpdq.computeHashes is the decode feed
for many videos: a host decode pool whose hashers all feed the device's one submission service, which applies
the back-pressure.
"""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Iterator

import numpy as np

from .. import vpdq

log = logging.getLogger(__name__)
log.setLevel(logging.CRITICAL)

# The dimensions of the image after downscaling for pdq (vpdqpy.py:23)
DOWNSCALE_DIMENSIONS = 512

VpdqHash = vpdq.VpdqHash  # vpdqpy.py:25


def point_resize_indices(src: int, dst: int = DOWNSCALE_DIMENSIONS) -> np.ndarray:
    """Source index of every destination pixel for swscale's POINT scaler (vpdqpy.py:90-95):
    xInc = ((src << 16) + (dst >> 1)) // dst ;  idx(i) = ((xInc >> 1) + i * xInc) >> 16  -- the
    centre-based nearest neighbour that reproduces the reference's golden hashes."""
    x_inc = ((src << 16) + (dst >> 1)) // dst
    i = np.arange(dst, dtype=np.int64)
    return np.minimum((((x_inc >> 1) + i * x_inc) >> 16), src - 1)


def point_resize_rgb(frame: np.ndarray, dst: int = DOWNSCALE_DIMENSIONS) -> np.ndarray:
    """[H, W, 3] u8 -> [dst, dst, 3] u8, nearest ("POINT") sampling."""
    h, w = frame.shape[:2]
    return np.ascontiguousarray(frame[point_resize_indices(h, dst)][:, point_resize_indices(w, dst)])


class Vpdq:
    @staticmethod
    def get_video_bytes(video_file: Path | str | bytes) -> bytes:
        """Get the bytes of a video (vpdqpy.py:30-47)."""
        if isinstance(video_file, (Path, str)):
            if not Path(video_file).is_file():
                raise ValueError("Failed to get video file bytes. Video does not exist")
            try:
                with open(str(video_file), "rb") as file:
                    return file.read()
            except OSError as exc:
                raise ValueError("Failed to get video file bytes. Invalid object type.") from exc
        elif isinstance(video_file, bytes):
            return video_file
        raise ValueError("Failed to get video file bytes. Invalid object type.")

    @staticmethod
    def match_hash(query_features: VpdqHash, target_features: VpdqHash, distance_tolerance: float = 31.0):
        """Get the similarity of two videos by comparing their list of features (vpdqpy.py:50-56)."""
        return vpdq.matchHash(query_features, target_features, int(distance_tolerance))

    # ---- decode (host side; not the accelerated path) -------------------------------------------------
    @staticmethod
    def frame_extract_pyav(video_bytes: bytes) -> Iterator[bytes]:
        """The reference's extractor (vpdqpy.py:59-101); yields 512x512 RGB24 frame bytes."""
        import io

        import av  # noqa: PLC0415  (optional dependency, same as the reference)

        with av.open(io.BytesIO(video_bytes), metadata_encoding="utf-8", metadata_errors="ignore") as container:
            video_streams = container.streams.video
            if video_streams is None or len(video_streams) < 1:
                raise ValueError("Video stream not found.")
            video = container.streams.video[0]
            video.thread_type = "AUTO"
            raw_average_fps = video.average_rate
            average_fps = 1
            if raw_average_fps is None or raw_average_fps < 1:
                log.warning("Average FPS is None or less than 1. Every frame will be hashed.")
            else:
                average_fps = round(raw_average_fps)
            frame_generator = container.decode(video)
            frame_index = 0
            while True:
                try:
                    frame = next(frame_generator)
                    if frame_index % average_fps == 0:
                        out = frame.reformat(width=DOWNSCALE_DIMENSIONS, height=DOWNSCALE_DIMENSIONS, format="rgb24",
                                             interpolation=av.video.reformatter.Interpolation.POINT)
                        yield bytes(out.planes[0])
                    frame_index += 1
                except StopIteration:
                    break
                except av.error.InvalidDataError as exc:
                    log.error(f"Skipping bad frame at index {frame_index}: {exc}")
                    frame_index += 1

    @staticmethod
    def frame_extract_cv2(video_bytes: bytes) -> Iterator[bytes]:
        """Same sampling rule with OpenCV's FFmpeg (the image this was built in has no PyAV)."""
        import os
        import tempfile

        import cv2  # noqa: PLC0415

        with tempfile.NamedTemporaryFile(suffix=".video", delete=False) as tmp:
            tmp.write(video_bytes)
            path = tmp.name
        cap = None
        try:
            cap = cv2.VideoCapture(path)
            if not cap.isOpened():
                raise ValueError("Video stream not found.")
            fps = cap.get(cv2.CAP_PROP_FPS)
            average_fps = 1 if (not fps or fps != fps or fps < 1) else round(fps)
            frame_index = 0
            while True:
                ok, bgr = cap.read()
                if not ok:
                    break
                if frame_index % average_fps == 0:
                    yield point_resize_rgb(bgr[:, :, ::-1]).tobytes()
                frame_index += 1
        finally:  # also when the generator is abandoned early
            if cap is not None:
                cap.release()
            os.unlink(path)

    _warned_cv2 = False

    @staticmethod
    def frame_extract(video_bytes: bytes) -> Iterator[bytes]:
        try:
            import av  # noqa: F401, PLC0415
        except ImportError:
            if not Vpdq._warned_cv2:
                Vpdq._warned_cv2 = True
                import warnings

                warnings.warn(
                    "PyAV is not installed: decoding with OpenCV instead.  That path is NOT bit-compatible with the "
                    "reference's (sampling on CAP_PROP_FPS instead of average_rate, rotation metadata applied, OpenCV's "
                    "YUV->RGB conversion): hashes may differ from the reference's by a few bits per frame.",
                    RuntimeWarning, stacklevel=2)
            return Vpdq.frame_extract_cv2(video_bytes)
        return Vpdq.frame_extract_pyav(video_bytes)

    @staticmethod
    def computeHash(video_file: Path | str | bytes, num_threads: int = 0) -> VpdqHash:
        """Perceptually hash video from a file path or the bytes (vpdqpy.py:104-119)."""
        video = Vpdq.get_video_bytes(video_file)
        if video is None:
            raise ValueError
        average_fps = 1  # discarded by the hasher, as in the reference (vpdqpy.py:110-112)
        hasher = vpdq.VideoHasher(average_fps, DOWNSCALE_DIMENSIONS, DOWNSCALE_DIMENSIONS, num_threads)
        try:
            for frame in Vpdq.frame_extract(video):
                hasher.hash_frame(frame)  # blocks while the device ring is full
            return hasher.finish()
        finally:
            hasher.close()

    @staticmethod
    def computeHashes(video_files, num_threads: int = 0) -> list:
        """Decode feed for MANY videos: a pool of host decode threads (FFmpeg releases the GIL), one
        VideoHasher per video as in computeHash.  All hashers feed the device's one submission service, so frames of
        different videos share uploads and kernel launches, and a full ring blocks the decoders -- the reference's
        hash_frame back-pressure (vpdqpy.py:115-117) -- instead of buffering frames without bound.
        num_threads <= 0: min(8, cpu count).  Returns the VpdqHash of every video, in input order; a video that
        fails to decode raises from here exactly as computeHash would."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        videos = list(video_files)
        n = num_threads if num_threads and num_threads > 0 else min(8, os.cpu_count() or 1)
        if len(videos) <= 1 or n == 1:
            return [Vpdq.computeHash(v) for v in videos]
        with ThreadPoolExecutor(max_workers=n) as pool:
            return list(pool.map(Vpdq.computeHash, videos))

    @staticmethod
    def is_similar(vpdq_features1: VpdqHash, vpdq_features2: VpdqHash, threshold: float = 75.0) -> tuple[bool, float]:
        """Threshold is minimum similarity to be considered similar (vpdqpy.py:122-131)."""
        similarity = Vpdq.match_hash(query_features=vpdq_features1, target_features=vpdq_features2)
        return similarity >= threshold, similarity
