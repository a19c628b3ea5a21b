"""Generates the committed golden fixtures from the reference's own test database.

Run ONCE in the authoring container (needs /root/reference and cv2); the GPU box has neither, so
tests only read the files written here:

  tests/golden/video_hashes/*.txt     verbatim copies of the reference's known-answer files
                                      (/root/reference/tests/testdb/video hashes/*.txt, read by
                                      tests/unit_tests/test_vpdqpy.py:105-128) -- data, CC-BY 3.0
  tests/golden/bbb_gif_frames.npz     the 10 frames the reference samples from
                                      S01_Big_Buck_Bunny_360_10s.gif (every round(fps)-th frame,
                                      vpdqpy.py:71-77,89), at native 360x640 RGB, before the POINT resize
  tests/golden/bbb_clip_frames.npz    for each of the 5 other Big Buck Bunny clips that OpenCV can decode here
                                      (h264 / vp9; the Sintel clips are AV1): two of the 10 sampled frames AFTER the
                                      512x512 POINT resize, PNG-compressed (lossless), with their indices.  These
                                      are a SOFT pin: OpenCV's YUV->RGB conversion differs from PyAV/swscale's by an
                                      LSB here and there, which moves 0..4 hash bits per frame (measured); the
                                      reference's own test only demands >= 99 % similarity
                                      (tests/unit_tests/test_vpdqpy.py:116-128)

The clips are (c) Blender Foundation | www.bigbuckbunny.org / durian.blender.org, CC-BY 3.0
(tests/unit_tests/test_vpdqpy.py:3-8).
"""
from __future__ import annotations

import shutil
from pathlib import Path

import cv2
import numpy as np

REF = Path("/root/reference/tests/testdb")
HERE = Path(__file__).resolve().parent


def sampled_frames(path: Path) -> np.ndarray:
    cap = cv2.VideoCapture(str(path))
    fps = cap.get(cv2.CAP_PROP_FPS)
    step = 1 if (not fps or fps < 1) else round(fps)  # vpdqpy.py:71-77
    out, idx = [], 0
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        if idx % step == 0:  # vpdqpy.py:89
            out.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
        idx += 1
    return np.stack(out)


BBB_CLIPS = ["S01_Big_Buck_Bunny_1080_10s_5MB-vp9.webm", "S01_Big_Buck_Bunny_1080_10s_5MB_H264.mp4",
             "S01_Big_Buck_Bunny_360_10s_5MB_H264.mp4", "S01_Big_Buck_Bunny_720_10s_1MB.mkv",
             "S01_Big_Buck_Bunny_720_10s_5MB_H264.mp4"]


def clip_fixtures() -> None:
    import sys

    sys.path.insert(0, str(HERE.parents[1]))
    import oracle
    from hydrus_video_deduplicator_b200.vpdqpy.vpdqpy import point_resize_rgb

    out = {}
    for c in BBB_CLIPS:
        frames = sampled_frames(REF / "videos" / "big_buck_bunny" / c)
        small = np.stack([point_resize_rgb(f) for f in frames])
        gold = np.frombuffer(bytes.fromhex((REF / "video hashes" / (c + ".txt")).read_text().strip()), np.uint8).reshape(-1, 32)
        h, _ = oracle.pdq_hash_frames(small, nthreads=8)
        dist = [int(np.unpackbits(h[k] ^ gold[k]).sum()) for k in range(len(gold))]
        exact = [k for k, d in enumerate(dist) if d == 0]
        other = [k for k, d in enumerate(dist) if d != 0]
        pick = sorted([exact[len(exact) // 2], other[0] if other else exact[0]])
        key = c.replace(".", "_").replace("-", "_")
        out[key + "__idx"] = np.array(pick, np.int32)
        for j, k in enumerate(pick):
            ok, buf = cv2.imencode(".png", cv2.cvtColor(small[k], cv2.COLOR_RGB2BGR), [cv2.IMWRITE_PNG_COMPRESSION, 9])
            assert ok
            out[f"{key}__png{j}"] = np.frombuffer(buf.tobytes(), np.uint8)
        print(c, "frames", pick, "bit distances to the golden hashes", [dist[k] for k in pick], "all:", dist)
    np.savez(HERE / "bbb_clip_frames.npz", **out)
    print("->", (HERE / "bbb_clip_frames.npz").stat().st_size, "bytes")


def main() -> None:
    dst = HERE / "video_hashes"
    dst.mkdir(exist_ok=True)
    for f in sorted((REF / "video hashes").glob("*.txt")):
        shutil.copyfile(f, dst / f.name)
    gif = REF / "videos" / "big_buck_bunny" / "S01_Big_Buck_Bunny_360_10s.gif"
    frames = sampled_frames(gif)
    assert frames.shape == (10, 360, 640, 3), frames.shape
    np.savez_compressed(HERE / "bbb_gif_frames.npz", frames=frames)
    print("frames", frames.shape, "->", (HERE / "bbb_gif_frames.npz").stat().st_size, "bytes")
    clip_fixtures()


if __name__ == "__main__":
    main()
