"""Generates the committed golden fixtures from the reference's own test database.

Run ONCE in the authoring container (needs /root/reference and cv2); the GPU box has neither, so
tests only read the files written here:

  tests/golden/video_hashes/*.txt     verbatim copies of the reference's known-answer files
                                      (/root/reference/tests/testdb/video hashes/*.txt, read by
                                      tests/unit_tests/test_vpdqpy.py:105-128) -- data, CC-BY 3.0
  tests/golden/bbb_gif_frames.npz     the 10 frames the reference samples from
                                      S01_Big_Buck_Bunny_360_10s.gif (every round(fps)-th frame,
                                      vpdqpy.py:71-77,89), at native 360x640 RGB, before the POINT resize

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


if __name__ == "__main__":
    main()
