"""2-GPU NCCL parity test of the sharded similarity search (VERDICT r01 missing 1): two ranks, the
target database sharded at video boundaries, the real CUDA scan / pairs / reduce kernels on each shard, NCCL
all_gather of the per-video counts, pair lists and candidate bitmaps -- compared with the oracle on the whole DB.
Skipped on a box with fewer than two GPUs (the gloo world-2 test covers the exchange logic on CPU)."""
from __future__ import annotations

import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["VPDQ_ROOT"])
import numpy as np
import torch
import oracle
from hydrus_video_deduplicator_b200 import dedupe, dist as hdist
from tests import synth

rank, world, local = hdist.init("nccl")
dev = torch.device("cuda", local)
vids, offsets = synth.synth_video_db(240, 0, seed=12, dup_frac=0.35)
long_vid = np.concatenate([np.frombuffer(vids[k], np.uint8).reshape(-1, 32) for k in (3, 50, 99, 180)] * 6)
vids[17] = long_vid.tobytes()                       # > 64 frames: several scan chunks
offsets = np.concatenate([[0], np.cumsum([len(v) // 32 for v in vids])]).astype(np.int64)
db = np.frombuffer(b"".join(vids), np.uint8).reshape(-1, 32)
index = hdist.ShardedIndex(db, offsets, dev)

# (1) per-video matched-frame counts of a long and a short query, every rank gets the full vector
for qv in (17, 7, 200):
    q = np.frombuffer(vids[qv], np.uint8).reshape(-1, 32)
    if len(q) == 0:
        continue
    got = index.matched_frames(q).cpu().numpy()
    assert (got == oracle.video_matched(q, db, offsets, 31)).all(), qv

# (2) all pairs of a query block against the sharded DB + the OR of the all-gathered candidate bitmaps
qb = np.ascontiguousarray(db[::3])
pairs, bitmap = index.candidate_pairs(torch.from_numpy(qb.copy()).to(dev), 31)
ref = oracle.hamming_pairs(qb, db, 31)
assert {tuple(p) for p in pairs.cpu().numpy().tolist()} == {tuple(p) for p in ref.tolist()}
bits = np.unpackbits(bitmap.cpu().numpy().view(np.uint8), bitorder="little")[: len(qb)]
assert (np.flatnonzero(bits) == np.unique(ref[:, 0])).all()

# (3) whole-table dedupe, targets sharded, rows all-gathered
a, b, d = dedupe.find_duplicate_videos(torch.from_numpy(db.copy()).to(dev), torch.from_numpy(offsets).to(dev), threshold=50.0)
got = sorted(zip(a.tolist(), b.tolist(), d.tolist()))
want = sorted((q, v, dist) for q in range(len(vids)) for v, dist in oracle.search_file(vids, q, 51) if v != q)
assert got == want, (len(got), len(want))
hdist.barrier()
torch.distributed.destroy_process_group()
print(f"rank {rank} ok: {len(got)} duplicate rows, {len(ref)} frame pairs", flush=True)
'''


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_sharded_search_on_two_gpus_matches_the_oracle(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the exchange logic is covered by tests/test_dist_gloo.py")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, VPDQ_ROOT=str(ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                       capture_output=True, text=True, env=env, timeout=580)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
