import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a box without a GPU is a skip, never a silent pass through some fallback
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir() -> Path:
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def bbb_clip_frames(golden_dir):
    """{clip file name: (frame indices [k], frames [k, 512, 512, 3] uint8, golden hashes [10, 32] uint8)} for the five
    OpenCV-decodable Big Buck Bunny clips (tests/golden/make_golden.py): post-POINT-resize frames, PNG-compressed."""
    import cv2
    import numpy as np

    z = np.load(golden_dir / "bbb_clip_frames.npz")
    out = {}
    for p in sorted((golden_dir / "video_hashes").glob("S01_Big_Buck_Bunny*.txt")):
        name = p.name[:-4]
        key = name.replace(".", "_").replace("-", "_")
        if key + "__idx" not in z:
            continue
        idx = z[key + "__idx"]
        frames = np.stack([cv2.cvtColor(cv2.imdecode(z[f"{key}__png{j}"], cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
                           for j in range(len(idx))])
        gold = np.frombuffer(bytes.fromhex(p.read_text().strip()), np.uint8).reshape(-1, 32)
        out[name] = (idx, np.ascontiguousarray(frames), gold)
    assert len(out) == 5
    return out
