"""world_size-2 gloo tests (CPU) for the multi-GPU host logic in dist.py: video-boundary sharding, the
padded all_gather of per-video masks / pair lists, and the all_gather + OR of candidate bitmaps.  The
compute between the collectives is stood in for by the oracle (this is a test of the exchange logic; the
CUDA scan/pairs kernels are covered by the -m gpu parity tests)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from hydrus_video_deduplicator_b200 import dist as hdist
from tests import synth


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_videos_cuts_only_at_video_boundaries():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        lens = rng.integers(0, 40, size=101)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        b = hdist.shard_videos(offsets, world)
        assert b[0] == 0 and b[-1] == 101 and (np.diff(b) >= 0).all()
        frames = np.diff(offsets[b])
        assert frames.sum() == offsets[-1]
        assert frames.max() - frames.min() <= 2 * lens.max()  # balanced to within a video
        db = (np.arange(offsets[-1] * 32, dtype=np.int64) % 251).astype(np.uint8).reshape(-1, 32)
        cat = np.concatenate([hdist.local_shard(db, offsets, world, r)[0] for r in range(world)])
        assert (cat == db).all()
    # degenerate: fewer videos than ranks, empty DB
    b = hdist.shard_videos(np.array([0, 5]), 4)
    assert b[0] == 0 and b[-1] == 1 and (np.diff(b) >= 0).all()
    assert hdist.shard_videos(np.array([0]), 2).tolist() == [0, 0, 0]
    assert hdist.round_robin(10, 4, 1).tolist() == [1, 5, 9]


def _worker(rank: int, world: int, port: int, tmp: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vids, offsets = synth.synth_video_db(60, 0, seed=12, dup_frac=0.3)
        db = np.frombuffer(b"".join(vids), np.uint8).reshape(-1, 32)
        shard, off, v0 = hdist.local_shard(db, offsets, world, rank)
        f0 = int(offsets[v0])
        query = np.frombuffer(vids[7] + vids[21], np.uint8).reshape(-1, 32)

        # (1) per-video matched counts: local (oracle stands in for the scan kernel) -> merged
        if len(off) > 1:
            local = torch.from_numpy(oracle.video_matched(query, np.ascontiguousarray(shard), off, 31))
        else:
            local = torch.zeros(0, dtype=torch.int32)
        merged = hdist.merge_video_masks(local).numpy()
        ref = oracle.video_matched(query, db, offsets, 31)
        assert (merged == ref).all()

        # (2) pair lists with target rebasing + (3) candidate bitmaps all_gather + OR
        lp = oracle.hamming_pairs(query, np.ascontiguousarray(shard), 31)
        pairs = hdist.merge_pairs(torch.from_numpy(lp), f0).numpy()
        refp = oracle.hamming_pairs(query, db, 31)
        assert {tuple(p) for p in pairs.tolist()} == {tuple(p) for p in refp.tolist()}
        bits = np.zeros((len(query) + 31) // 32 * 32, np.uint8)
        bits[np.unique(lp[:, 0])] = 1
        words = torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy())
        gathered = hdist.all_gather_bitmaps(words)
        assert gathered.shape == (world, words.numel())
        allbits = np.unpackbits(hdist.or_reduce(gathered).numpy().view(np.uint8), bitorder="little")
        assert (np.flatnonzero(allbits) == np.unique(refp[:, 0])).all()

        # (4) variable-length gather incl. an empty contribution
        t = torch.arange(rank * 3, dtype=torch.int64).reshape(-1, 1)
        parts = hdist.all_gather_varlen(t)
        assert [p.shape[0] for p in parts] == [r * 3 for r in range(world)]
        with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_cpulist_parsing():
    assert hdist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hdist._parse_cpulist("") == set()
    assert hdist.bind_to_gpu_numa_node(0) is None or isinstance(hdist.bind_to_gpu_numa_node(0), int)


def test_query_chunking():
    """device.chunk_rows: query videos cut into scan chunks of <= 64 frames -- ragged, empty and exactly-64 videos."""
    from hydrus_video_deduplicator_b200 import device

    lens = [0, 1, 64, 65, 128, 0, 300, 63, 0]
    off = np.concatenate([[0], np.cumsum(lens)])
    rows, qv = device.chunk_rows(off)
    assert qv.tolist() == [0, 0, 1, 2, 4, 6, 6, 11, 12, 12]
    assert rows[0] == 0 and rows[-1] == off[-1] and (np.diff(rows) <= 64).all() and (np.diff(rows) > 0).all()
    for v, n in enumerate(lens):  # every video's chunks tile exactly its rows
        c0, c1 = qv[v], qv[v + 1]
        assert c1 - c0 == (n + 63) // 64
        if n:
            assert rows[c0] == off[v] and rows[c1] == off[v + 1]
    rows, qv = device.chunk_rows([0])
    assert rows.tolist() == [0] and qv.tolist() == [0]
