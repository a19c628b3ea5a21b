"""Batch de-duplication over a whole hash table on the GPU(s): the all-files form of what the reference does one file
at a time in ``find_potential_duplicates`` (``dedup.py:445-502``, ``db/vptree.py:865-902``).

For every video q of the table (in blocks, so that the per-chunk masks stay small):

    scan      k_hamming_scan, many query chunks per launch: bit i of qmask[c][v] = query frame i of chunk c has a
              frame of target video v within Hamming distance 31
 -> reduce    k_video_reduce: matched(q, v) = popcount over q's chunks       == numerator of vpdq.matchHash(q, v, 31)
              distance = (100 - 100 * matched // n_q) + 1                    == vptree.calculate_distance (vptree.py:22-31)
 -> rows      (q, v, matched, distance) with distance <= fix_vpdq_similarity(threshold), q != v

All of it hand-written CUDA behind the C ABI (vpdq_b200_hamming_scan_multi_dev, vpdq_b200_video_match_dev); torch
only holds the buffers.  With torch.distributed initialised the TARGET side is this rank's shard (video-aligned,
dist.shard_videos), queries are replicated, and the (small) row lists of all ranks are all-gathered.
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev_api
from . import dist as hdist


def find_duplicate_videos(hashes: torch.Tensor, offsets: torch.Tensor, threshold: float = 50.0, tolerance: int = 31,
                          query_frames_per_block: int = 1 << 16):
    """hashes [N, 32] uint8 CUDA (all videos, replicated on every rank), offsets [V+1] int64 CUDA.
    -> (a, b, distance) int64 tensors: every directed pair a != b with calculate_distance(a, b) <= radius.
    Sharded over the ranks of the default process group when one is initialised."""
    from .search import fix_vpdq_similarity

    world, rank = hdist.world_size(), hdist.rank()
    off_h = offsets.cpu().numpy().astype(np.int64)
    n_videos = len(off_h) - 1
    radius = fix_vpdq_similarity(threshold)
    bounds = hdist.shard_videos(off_h, world)
    v0, v1 = int(bounds[rank]), int(bounds[rank + 1])
    f0, f1 = int(off_h[v0]), int(off_h[v1])
    targets = hashes[f0:f1]
    t_off = (offsets[v0:v1 + 1] - f0).contiguous()
    found = []
    q0 = 0
    while q0 < n_videos:  # a block of query videos: about query_frames_per_block frames
        q1 = int(np.searchsorted(off_h, off_h[q0] + query_frames_per_block, side="left"))
        q1 = min(max(q1, q0 + 1), n_videos)
        q_off = off_h[q0:q1 + 1] - off_h[q0]
        if v1 > v0 and f1 > f0 and off_h[q1] > off_h[q0]:
            rows = dev_api.video_matches(targets, t_off, hashes[int(off_h[q0]):int(off_h[q1])], q_off, tolerance,
                                         max_distance=radius)
            if rows.numel():
                rows = rows.to(torch.int64)
                rows[:, 0] += q0
                rows[:, 1] += v0
                found.append(rows)
        q0 = q1
    mine = torch.cat(found) if found else torch.zeros((0, 4), dtype=torch.int64, device=hashes.device)
    rows = torch.cat(hdist.all_gather_varlen(mine))
    keep = rows[:, 0] != rows[:, 1]
    rows = rows[keep]
    return rows[:, 0], rows[:, 1], rows[:, 3]
