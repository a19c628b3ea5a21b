"""Whole-table de-duplication (BASELINE configs[4] in miniature): synthetic clips -> CUDA hashes -> all-pairs
frame matches -> video-level similarity, against the oracle's brute-force search_file on the same hashes."""
from __future__ import annotations

import numpy as np
import pytest

import oracle
from tests import synth

pytestmark = pytest.mark.gpu


def test_hash_and_dedupe_clips():
    import torch

    from hydrus_video_deduplicator_b200 import dedupe, device

    dev = torch.device("cuda", 0)
    n_clips, fpc = 24, 6
    base = synth.synth_frames(n_clips // 2 * fpc, seed=77)
    clips = [base[k * fpc:(k + 1) * fpc] for k in range(n_clips // 2)]
    clips += [synth.noisy_copy(c, seed=500 + k) for k, c in enumerate(clips)]  # clip k + 12 duplicates clip k
    clips[3] = clips[3].copy()
    clips[3][2:] = 0  # a clip whose tail is black: those frames are dropped by the quality filter
    frames = torch.from_numpy(np.concatenate(clips)).to(dev)
    hashes, quality = device.hash_frames(frames)
    keep = (quality >= 31).cpu().numpy()
    h = hashes.cpu().numpy()
    vids = [h[k * fpc:(k + 1) * fpc][keep[k * fpc:(k + 1) * fpc]].tobytes() for k in range(n_clips)]
    assert vids == [oracle.video_hash(c) for c in clips]
    offsets = np.concatenate([[0], np.cumsum([len(v) // 32 for v in vids])]).astype(np.int64)
    table = torch.from_numpy(np.frombuffer(b"".join(vids), np.uint8).reshape(-1, 32).copy()).to(dev)
    a, b, d = dedupe.find_duplicate_videos(table, torch.from_numpy(offsets).to(dev), threshold=50.0)
    got = sorted(zip(a.tolist(), b.tolist(), d.tolist()))
    ref = sorted((q, v, dist) for q in range(n_clips) for v, dist in oracle.search_file(vids, q, 51) if v != q)
    assert got == ref
    found = {(x, y) for x, y, _ in got}
    assert found >= {(k, k + 12) for k in range(12) if k != 3} | {(k + 12, k) for k in range(12) if k != 3}
    # the score is directional (percent of QUERY frames matched): the 2 surviving frames of clip 3 all match
    # clip 15, but only 2 of clip 15's 6 frames match clip 3 (33 % < 50 %)
    assert (3, 15) in found and (15, 3) not in found
