"""Batch de-duplication over a whole hash table on the GPU(s): the all-pairs form of what the reference does
one file at a time in ``find_potential_duplicates`` (``dedup.py:445-502``).

    frame pairs  (CUDA: k_hamming_pairs, popcount(q ^ t) <= 31)
 -> per (query video, target video): number of DISTINCT query frames with a match      (torch, tiny)
 -> similarity = 100 * matched / n_frames(query video)      == vpdq.matchHash(query, target, 31)
 -> distance   = (100 - int(similarity)) + 1                == vptree.calculate_distance (vptree.py:22-31)
 -> duplicates = directed pairs with distance <= fix_vpdq_similarity(threshold), a != b

With torch.distributed initialised the TARGET side is this rank's shard (video-aligned, dist.shard_videos) and
the frame pairs of all ranks are all-gathered before the (replicated, cheap) video-level reduction.
"""
from __future__ import annotations

import torch

from . import device as dev_api
from . import dist as hdist


def video_of(frame_idx: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """video index of every frame index (offsets: CSR int64 [V+1] on the same device)"""
    return torch.searchsorted(offsets, frame_idx.contiguous(), right=True) - 1


def video_similarities(pairs: torch.Tensor, q_offsets: torch.Tensor, t_offsets: torch.Tensor):
    """pairs [n, 2] int64 (query frame, target frame) -> (vq, vt, matched, similarity) per video pair that has
    at least one matching frame; similarity is float64 = 100 * matched / frames(vq)."""
    n_t = t_offsets.numel() - 1
    if pairs.numel() == 0:
        z = torch.zeros(0, dtype=torch.int64, device=q_offsets.device)
        return z, z, z, torch.zeros(0, dtype=torch.float64, device=q_offsets.device)
    vt = video_of(pairs[:, 1], t_offsets)
    qf_vt = torch.unique(pairs[:, 0] * n_t + vt)            # distinct (query frame, target video)
    qf, vt = qf_vt // n_t, qf_vt % n_t
    vq = video_of(qf, q_offsets)
    key, matched = torch.unique(vq * n_t + vt, return_counts=True)
    vq, vt = key // n_t, key % n_t
    n_q = (q_offsets[1:] - q_offsets[:-1])[vq]
    sim = (100.0 * matched.to(torch.float64)) / n_q.to(torch.float64)
    return vq, vt, matched, sim


def find_duplicate_videos(hashes: torch.Tensor, offsets: torch.Tensor, threshold: float = 50.0, tolerance: int = 31,
                          capacity: int = 1 << 22):
    """hashes [N, 32] uint8 CUDA (all videos, replicated on every rank), offsets [V+1] int64 CUDA.
    -> (a, b, distance) int64 tensors: every directed pair a != b with calculate_distance(a, b) <= radius.
    Sharded over the ranks of the default process group when one is initialised."""
    from .search import fix_vpdq_similarity

    world, rank = hdist.world_size(), hdist.rank()
    off_h = offsets.cpu().numpy()
    bounds = hdist.shard_videos(off_h, world)
    f0, f1 = int(off_h[bounds[rank]]), int(off_h[bounds[rank + 1]])
    targets = hashes[f0:f1]
    while True:
        n, pairs, _ = dev_api.hamming_pairs(hashes, targets, tolerance, capacity=capacity, want_bitmap=False)
        if n <= capacity:
            break
        capacity = int(n)  # the list overflowed: rerun this shard with room for everything
    pairs = hdist.merge_pairs(pairs, f0)
    vq, vt, _, sim = video_similarities(pairs, offsets, offsets)
    dist_ = (100 - sim.to(torch.int64)) + 1  # int() truncation, as fix_vpdq_similarity
    keep = (vq != vt) & (dist_ <= fix_vpdq_similarity(threshold))
    return vq[keep], vt[keep], dist_[keep]
