"""BASELINE configs[4]: pre-decoded synthetic clips -> hash -> dedupe on N GPUs (torchrun) next to the CPU path.

  python tools/e2e_dedupe.py [--clips 1000] [--frames 300]            (1 GPU)
  python -m torch.distributed.run --nproc-per-node N tools/e2e_dedupe.py ...

Clip k + clips/2 is a noisy copy (+-2 LSB) of clip k, so clips/2 true duplicate pairs exist.  Frames are
generated on the device (there is no decoder in the loop: "pre-decoded"), each rank hashes its round-robin
share of the clips, the 32-byte hashes are all-gathered, the all-pairs search is sharded over the target side.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np
import torch

from bench import device_frames
from hydrus_video_deduplicator_b200 import dedupe, device, dist as hdist


def clip_frames(k: int, n_clips: int, fpc: int, dev):
    src = k % (n_clips // 2)
    f = device_frames(torch, fpc, dev, seed=10_000 + src)
    if k >= n_clips // 2:
        g = torch.Generator(device=dev).manual_seed(90_000 + k)
        noise = torch.randint(-2, 3, f.shape, dtype=torch.int16, device=dev, generator=g)
        f = (f.to(torch.int16) + noise).clamp_(0, 255).to(torch.uint8)
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1000)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--cpu-clips", type=int, default=8, help="clips hashed by the CPU oracle for the baseline rate")
    ap.add_argument("--cpu-query-clips", type=int, default=64, help="query clips searched by the CPU oracle (dedupe baseline)")
    ap.add_argument("--out", type=str, default="", help="also write the JSON line to this file")
    args = ap.parse_args()
    rank, world, local = hdist.init()
    dev = torch.device("cuda", local)
    mine = hdist.round_robin(args.clips, world, rank)
    device.hash_frames(clip_frames(0, args.clips, min(args.frames, 32), dev))  # warm-up: module load, attributes
    torch.cuda.synchronize()
    t_hash = 0.0
    local_hashes = torch.zeros((len(mine), args.frames, 32), dtype=torch.uint8, device=dev)
    local_quality = torch.zeros((len(mine), args.frames), dtype=torch.int32, device=dev)
    group = max(1, 8192 // args.frames)  # clips hashed per launch pair (the frames of one group live in HBM together)
    for n0 in range(0, len(mine), group):
        ks = mine[n0:n0 + group]
        frames = torch.cat([clip_frames(int(k), args.clips, args.frames, dev) for k in ks])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h, q = device.hash_frames(frames)
        torch.cuda.synchronize()
        t_hash += time.perf_counter() - t0
        local_hashes[n0:n0 + len(ks)] = h.view(len(ks), args.frames, 32)
        local_quality[n0:n0 + len(ks)] = q.view(len(ks), args.frames)
    # exchange: every rank gets the whole table, in clip order
    parts_h = hdist.all_gather_varlen(local_hashes.reshape(len(mine), -1))
    parts_q = hdist.all_gather_varlen(local_quality)
    table = torch.zeros((args.clips, args.frames, 32), dtype=torch.uint8, device=dev)
    qual = torch.zeros((args.clips, args.frames), dtype=torch.int32, device=dev)
    for r in range(world):
        idx = torch.from_numpy(hdist.round_robin(args.clips, world, r)).to(dev)
        table[idx] = parts_h[r].reshape(-1, args.frames, 32)
        qual[idx] = parts_q[r]
    keep = qual >= 31  # finish(): drop low-quality frames (DedupeDB.py:550-553)
    counts = keep.sum(dim=1)
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(counts, 0)])
    flat = table[keep]
    torch.cuda.synchronize()
    hdist.barrier()
    t0 = time.perf_counter()
    a, b, d = dedupe.find_duplicate_videos(flat, offsets, threshold=50.0)
    torch.cuda.synchronize()
    t_dedupe = time.perf_counter() - t0
    found = {(int(x), int(y)) for x, y in zip(a.tolist(), b.tolist())}
    half = args.clips // 2
    planted = {(k, k + half) for k in range(half)} | {(k + half, k) for k in range(half)}
    t = torch.tensor([t_hash, t_dedupe], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        out = {"clips": args.clips, "frames_per_clip": args.frames, "n_gpus": world,
               "hash_s": float(t[0]), "frames_per_s": args.clips * args.frames / float(t[0]),
               "dedupe_s": float(t[1]), "frame_pairs_per_s": float(flat.shape[0]) ** 2 / float(t[1]),
               "kept_frames": int(flat.shape[0]), "duplicate_pairs_found": len(found) // 2,
               "planted_pairs": half, "planted_recall": len(found & planted) / max(1, len(planted)),
               "false_pairs": len(found - planted) // 2}
        if args.cpu_clips:
            import oracle

            cores = os.cpu_count() or 1
            fr = torch.cat([clip_frames(k, args.clips, args.frames, dev) for k in range(args.cpu_clips)]).cpu().numpy()
            t0 = time.perf_counter()
            ref_h, ref_q = oracle.pdq_hash_frames(fr, nthreads=cores)
            dt = time.perf_counter() - t0
            out["cpu_frames_per_s"] = len(fr) / dt
            out["cpu_cores"] = cores
            out["cpu_sample"] = f"{args.cpu_clips} clips x {args.frames} frames, oracle port"
            got = table[: args.cpu_clips].reshape(-1, 32).cpu().numpy()
            out["parity_vs_oracle"] = bool((got == ref_h).all() and (qual[: args.cpu_clips].reshape(-1).cpu().numpy() == ref_q).all())
            # the CPU side of the dedupe: brute-force search of a SAMPLE of query clips against the whole table on all
            # cores (what the reference's search_file loop computes, dedup.py:445-502), extrapolated to all clips
            flat_h = flat.cpu().numpy()
            off_h = offsets.cpu().numpy()
            nq = min(args.cpu_query_clips, args.clips)
            from concurrent.futures import ThreadPoolExecutor

            def cpu_search(qv):  # (the C oracle releases the GIL: one query clip per host thread)
                q = flat_h[off_h[qv]:off_h[qv + 1]]
                if len(q) == 0:
                    return set()
                m = oracle.video_matched(q, flat_h, off_h, 31)
                dist = (100 - (100 * m.astype(np.int64)) // len(q)) + 1
                return {(qv, int(v)) for v in np.flatnonzero((m > 0) & (dist <= 51)) if int(v) != qv}

            t0 = time.perf_counter()
            cpu_rows = set()
            with ThreadPoolExecutor(cores) as ex:
                for rows in ex.map(cpu_search, range(nq)):
                    cpu_rows |= rows
            dt = time.perf_counter() - t0
            out["cpu_dedupe_s_sample"] = dt
            out["cpu_dedupe_sample"] = f"{nq} of {args.clips} query clips vs the whole table, oracle port"
            out["cpu_dedupe_s_extrapolated"] = dt * args.clips / max(1, nq)
            out["cpu_total_s_extrapolated"] = args.clips * args.frames / out["cpu_frames_per_s"] + out["cpu_dedupe_s_extrapolated"]
            out["gpu_total_s"] = out["hash_s"] + out["dedupe_s"]
            out["speedup_vs_cpu_extrapolated"] = out["cpu_total_s_extrapolated"] / out["gpu_total_s"]
            out["dedupe_rows_match_cpu_sample"] = bool({(x, y) for x, y in found if x < nq} == cpu_rows)
        line = json.dumps(out)
        print(line, flush=True)
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(line + "\n")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
