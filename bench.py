#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the B200 VPDQ hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload hash|hamming]
  (N > 1: launched by torchrun, one rank per GPU; rank 0 prints ONE JSON line)

workload "hash" (default; BASELINE.json configs[1]: "hash 100k synthetic frames"):
    a step = one pass of the PDQ frame-hash path over one batch of synthetic 512x512 RGB24 frames.
    value = frames/s with the batch already resident in HBM (CUDA events on the launching stream, max over
            ranks); e2e = the same through the host-pointer C ABI call (pinned host memory -> H2D -> kernels
            -> D2H inside the timed region), next to the measured copy-only roof of the same buffers and to the rates
            of the reference-shaped API (vpdq.VideoHasher.hash_frame(bytes), one hasher per video, 300- and 10-frame
            videos).  Each step's input (6.4 GB) is far larger than L2 (126 MB).
    Every N also carries, under "hamming", BASELINE configs[3]: a FIXED 10 M-hash database (300-frame videos,
    SURVEY config 3's generator) sharded at video boundaries over the N ranks, a replicated query block, one NCCL
    all_gather of the candidate bitmaps per step inside the timed region -> pair-comparisons/s, plus the collective's
    own time.  At N = 1 additionally: streaming scan GB/s vs the HBM peak by resident query count, 1M x 1M all pairs,
    and the same kernels on a video-like database (dense near-duplicates).
workload "hamming": the sharded all-pairs line above as the headline value (same config).

--impl reference times the reference's CPU path (the golden-pinned oracle port; the real arithmetic lives
in the absent hvdaccelerators wheel) on the host cores with all threads, same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAME_BYTES = 512 * 512 * 3
ALGO_BYTES_PER_FRAME = FRAME_BYTES + 32 + 4  # RGB24 in, hash + quality out
HASH_BYTES = 32
_REAL_STDOUT = None
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def jarosz_kernel_name() -> str:
    return "kx_systolic_jarosz"


def finalize_kernel_name() -> str:
    return "k5_finalize"


def measured_traffic_per_frame() -> float | None:
    """DRAM bytes per frame of the dominant PDQ kernel from the committed ncu capture (profiles/r02_traffic.json)."""
    p = ROOT / "profiles" / "r02_traffic.json"
    try:
        d = json.loads(p.read_text())
        k = d[jarosz_kernel_name()]
        return (float(k["dram_read_bytes"]) + float(k["dram_write_bytes"])) / float(k["frames"])
    except Exception:
        return None


def hbm_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


# ----------------------------------------------------------------------------------------------------
# clocks sampling (recipe: /opt/skills/guides/B200_PROFILING.md)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower() == "active":
                    reasons.add(name)
        self.tmp.close()
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------------
def device_frames(torch, n: int, device, seed: int):
    """[n, 512, 512, 3] u8 on the device: frame k is noise / smooth / blocks for k % 3 = 0 / 1 / 2
    (the generator of tests/synth.py, restated with torch ops so that 100k frames take seconds)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, 512, 512, 3), dtype=torch.uint8, device=device)
    step = 256
    for f0 in range(0, n, step):
        m = min(step, n - f0)
        blk = out[f0:f0 + m]
        blk.copy_(torch.randint(0, 256, (m, 512, 512, 3), dtype=torch.uint8, device=device, generator=g))
        ks = torch.arange(f0, f0 + m, device=device) % 3
        i1 = torch.nonzero(ks == 1).flatten()
        if i1.numel():
            grid = torch.rand((i1.numel(), 3, 8, 8), device=device, generator=g) * 255.0
            up = torch.nn.functional.interpolate(grid, size=(512, 512), mode="bilinear", align_corners=False)
            blk[i1] = up.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
        i2 = torch.nonzero(ks == 2).flatten()
        if i2.numel():
            small = torch.randint(0, 256, (i2.numel(), 32, 32, 3), dtype=torch.uint8, device=device, generator=g)
            blk[i2] = small.repeat_interleave(16, dim=1).repeat_interleave(16, dim=2)
    return out


def _pack_bits(torch, bits):
    """[n, 256] {0,1} uint8 -> [n, 32] uint8, bit k of a hash in byte k >> 3, bit k & 7 (native PDQ order)"""
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=bits.device)
    return (bits.view(-1, 32, 8).to(torch.int32) * w).sum(dim=2).to(torch.uint8)


def _flip_bits(torch, h, n_flip, g):
    """flip n_flip[i] distinct random bits of hash i ([n, 32] uint8; n_flip [n] int64, <= 64)"""
    n = h.shape[0]
    order = torch.rand((n, 256), device=h.device, generator=g).argsort(dim=1)[:, :64]       # 64 distinct positions
    take = torch.arange(64, device=h.device).unsqueeze(0) < n_flip.unsqueeze(1)
    mask = torch.zeros((n, 256), dtype=torch.uint8, device=h.device)
    mask.scatter_(1, order, take.to(torch.uint8))
    return h ^ _pack_bits(torch, mask)


def device_hashes(torch, n: int, device, seed: int, planted_frac: float = 0.01):
    """SURVEY.md 8d config 3 on the device: [n, 32] u8 random words of popcount exactly 128 (like genuine PDQ hashes,
    F5) with a planted_frac share of near-duplicates at Hamming distances 0, 2, .., 40 -- both sides of the tolerance
    31 (the generator of tests/synth.py, restated with torch ops so that 10 M hashes take seconds)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, 32), dtype=torch.uint8, device=device)
    step = 1 << 19
    for i0 in range(0, n, step):
        m = min(step, n - i0)
        rank = torch.rand((m, 256), device=device, generator=g).argsort(dim=1).argsort(dim=1)
        out[i0:i0 + m] = _pack_bits(torch, (rank < 128).to(torch.uint8))
    k = int(n * planted_frac)
    if k and n > 2 * k:
        src = torch.randint(0, n // 2, (k,), device=device, generator=g)
        dst = n // 2 + torch.randperm(n - n // 2, device=device, generator=g)[:k]
        dist = 2 * torch.randint(0, 21, (k,), device=device, generator=g)
        for i0 in range(0, k, step):
            sl = slice(i0, min(k, i0 + step))
            out[dst[sl]] = _flip_bits(torch, out[src[sl]], dist[sl], g)
    return out


def device_video_hashes(torch, n_videos: int, frames_per_video: int, device, seed: int, dup_frac: float = 0.05):
    """A video-like database: every video is a random walk (consecutive frames differ in 0..16 bits, so a frame has
    many neighbours of its own video within the tolerance) and dup_frac of the videos are noisy copies (0..8 bits per
    frame) of another one -- the match density of a real library, unlike uniformly random hashes.
    -> ([n_videos * frames_per_video, 32] u8, offsets [n_videos + 1] i64)"""
    g = torch.Generator(device=device).manual_seed(seed)
    n = n_videos * frames_per_video
    out = torch.empty((n_videos, frames_per_video, 32), dtype=torch.uint8, device=device)
    cur = device_hashes(torch, n_videos, device, seed + 1, planted_frac=0.0)
    for k in range(frames_per_video):
        out[:, k] = cur
        cur = _flip_bits(torch, cur, torch.randint(0, 17, (n_videos,), device=device, generator=g), g)
    n_dup = int(n_videos * dup_frac)
    if n_dup:
        src = torch.randint(0, n_videos // 2, (n_dup,), device=device, generator=g)
        dst = n_videos // 2 + torch.randperm(n_videos - n_videos // 2, device=device, generator=g)[:n_dup]
        for i0 in range(0, n_dup, 256):
            sl = slice(i0, min(n_dup, i0 + 256))
            flat = out[src[sl]].reshape(-1, 32)
            flat = _flip_bits(torch, flat, torch.randint(0, 9, (flat.shape[0],), device=device, generator=g), g)
            out[dst[sl]] = flat.view(-1, frames_per_video, 32)
    offsets = torch.arange(0, n + 1, frames_per_video, dtype=torch.int64, device=device)
    return out.view(n, 32), offsets


# ----------------------------------------------------------------------------------------------------
# CPU reference (oracle port) -- the one place outside tests/ and smoke() that may run oracle/
# ----------------------------------------------------------------------------------------------------
def cpu_frames(n_frames: int):
    import numpy as np

    from tests import synth

    base = synth.synth_frames(min(n_frames, 48), seed=0)
    reps = (n_frames + len(base) - 1) // len(base)
    return np.ascontiguousarray(np.concatenate([base] * reps)[:n_frames])


def cpu_hash_rate(n_frames: int, threads: int, repeats: int = 1) -> tuple[float, float]:
    """-> (frames/s, seconds) hashing n_frames synthetic frames `repeats` times on `threads` host threads"""
    import oracle

    frames = cpu_frames(n_frames)
    oracle.pdq_hash_frames(frames[: min(len(frames), threads)], nthreads=threads)  # warm the tables
    t0 = time.perf_counter()
    for _ in range(repeats):
        oracle.pdq_hash_frames(frames, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_frames * repeats / dt, dt


def cpu_pairs_rate(n: int, threads: int) -> tuple[float, float]:
    import oracle
    from tests import synth

    h = synth.synth_hashes(n, seed=1)
    t0 = time.perf_counter()
    oracle.hamming_count_mt(h, h, 31, threads)
    dt = time.perf_counter() - t0
    return float(n) * n / dt, dt


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    cores = os.cpu_count() or 1
    if args.workload == "hash":
        import oracle

        per_step = 16 * cores  # bounded sample: ~0.1-0.2 s of all-core work per step
        frames = cpu_frames(per_step)
        for _ in range(args.warmup):
            oracle.pdq_hash_frames(frames, nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle.pdq_hash_frames(frames, nthreads=cores)
        secs = time.perf_counter() - t0
        rate = per_step * args.steps / secs
        value, unit, metric = rate, "frames/s", "frame_hashes_per_sec"
        ms = secs / max(1, args.steps) * 1e3
        sample = f"{args.steps} steps x {per_step} synthetic 512x512 RGB24 frames, {cores} threads (oracle port)"
        # same config object as the B200 arm (the metric is a rate; the CPU arm times a bounded SAMPLE of that
        # workload, described in cpu_baseline.sample)
        config = hash_config(args, args.batch)
        config["parallelism"] = f"{args.gpus} x independent shards (no data-path collective)"
    else:
        n = 16384
        for _ in range(max(1, args.warmup) - 1):
            cpu_pairs_rate(4096, cores)
        rates = [cpu_pairs_rate(n, cores) for _ in range(max(1, min(args.steps, 5)))]
        value = sum(r for r, _ in rates) / len(rates)
        ms = sum(s for _, s in rates) / len(rates) * 1e3
        unit, metric = "pair-comparisons/s", "pair_comparisons_per_sec"
        sample = f"{len(rates)} x ({n} x {n}) 256-bit hashes, tolerance 31, {cores} threads (oracle port)"
        config = hamming_config(args, args.gpus)
        config["parallelism"] = f"target DB sharded over {args.gpus} ranks, queries replicated, all_gather of bitmaps"
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.workload == "hash" else "u64", "data": "synthetic", "config": config,
        "impl": "reference",
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def hash_config(args, frames_per_step: int) -> dict:
    return {"workload": "PDQ/VPDQ frame hash of synthetic 512x512 RGB24 frames (BASELINE configs[1]: 100k frames)",
            "frames_per_step": frames_per_step, "frame_bytes": FRAME_BYTES,
            "l2_policy": "each step's input batch is larger than L2 (126 MB); no flush needed",
            "quality_filter": "none in the timed region (finish() filters on the host)"}


def hamming_config(args, world: int) -> dict:
    return {"workload": "all-pairs Hamming search, 256-bit hashes, tolerance 31 (BASELINE configs[3]): fixed 10 M-hash DB "
                        "sharded over the GPUs",
            "db_hashes": args.shard_db_hashes, "queries_per_step": args.query_block,
            "l2_policy": "L2 flushed between timed iterations by a 256 MB write"}


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import numpy as np
    import torch

    from hydrus_video_deduplicator_b200 import _ffi, device as dev_api, dist as hdist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    rank, world, local = hdist.init()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _ffi.lib()
    peak, peak_src = hbm_peak()

    def sync_all():
        torch.cuda.synchronize()
        hdist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush_buf.fill_(1)

    extra = {}
    if args.workload == "hash":
        B = args.batch
        pool_n = max(1, min(args.steps, args.pool_batches))
        pool = [device_frames(torch, B, dev, seed=1000 * rank + p) for p in range(pool_n)]
        for w in range(args.warmup):
            dev_api.hash_frames(pool[w % pool_n])
        sync_all()
        sampler = ClockSampler(local)
        launches0 = _ffi.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            hashes, quality = dev_api.hash_frames(pool[k % pool_n])
        e1.record()
        torch.cuda.synchronize()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        hdist.barrier()
        launches = _ffi.kernel_launches() - launches0
        value = world * B * args.steps / (ms_total / 1e3)
        ms_per_step = ms_total / args.steps
        # the dominant kernel alone (kx_fused_jarosz2 = vpdq_b200_pdq_jarosz_dev), CUDA events per launch
        a64 = torch.empty((B, 64, 64), dtype=torch.float32, device=dev)
        kx_ms = []
        stream = torch.cuda.current_stream().cuda_stream
        for k in range(3 + min(args.steps, 10)):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            _ffi.check(_ffi.lib().vpdq_b200_pdq_jarosz_dev(pool[k % pool_n].data_ptr(), B, 512, 512, a64.data_ptr(),
                                                          stream))
            eb.record()
            torch.cuda.synchronize()
            if k >= 3:
                kx_ms.append(ea.elapsed_time(eb))
        kx_ms = statistics.mean(kx_ms)
        del a64
        # roofline: algorithmic bytes per launch / that kernel's launch duration; the whole pipeline beside it
        achieved = B * ALGO_BYTES_PER_FRAME / (kx_ms / 1e3) / 1e9
        pipeline_gbs = B * ALGO_BYTES_PER_FRAME / (ms_per_step / 1e3) / 1e9
        tpf = measured_traffic_per_frame()
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (tpf * B) if tpf else None, "peak_source": peak_src,
                    "kernel": jarosz_kernel_name(), "kernel_ms_per_launch": kx_ms,
                    "kernel_share_of_step": kx_ms / ms_per_step,
                    "pipeline": {"kernels": jarosz_kernel_name() + " + " + finalize_kernel_name(), "achieved": pipeline_gbs,
                                 "frac": pipeline_gbs / peak},
                    "algorithmic_bytes_per_launch": B * ALGO_BYTES_PER_FRAME,
                    "algorithmic_bytes_per_frame": ALGO_BYTES_PER_FRAME,
                    "traffic_source": "profiles/r02_traffic.json (ncu --set full dram__bytes_read+write, per frame x "
                                      "frames per launch)",
                    "note": "one warp per frame, row-pass chain state handed lane to lane: no shared-memory transposes; two "
                            "branch-free loop bodies (plain rows / frame boundaries); bound by instruction issue (half-rate "
                            "packed fp32 and byte-unpack operations of the exact running sums), see DESIGN.md 4.1"}

        # ---- end to end through the host-pointer C ABI (pinned host memory) ----
        import ctypes as C

        e2e_batch = min(B, args.e2e_batch)
        h_frames = torch.empty((e2e_batch, 512, 512, 3), dtype=torch.uint8, pin_memory=True)
        h_frames.copy_(pool[0][:e2e_batch])
        h_hash = torch.empty((e2e_batch, 32), dtype=torch.uint8, pin_memory=True)
        h_q = torch.empty((e2e_batch,), dtype=torch.int32, pin_memory=True)
        torch.cuda.synchronize()

        def e2e_step():
            _ffi.check(_ffi.lib().vpdq_b200_pdq_hash_frames_host(
                C.c_void_p(h_frames.data_ptr()), 3, e2e_batch, 512, 512, C.c_void_p(h_hash.data_ptr()),
                C.c_void_p(h_q.data_ptr()), local))

        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        clocks = sampler.stop()  # sampled across both timed regions (device-resident steps, then e2e steps)
        e2e = {"value": world * e2e_batch * e2e_steps / e2e_s, "unit": "frames/s",
               "h2d_bytes_per_step": e2e_batch * FRAME_BYTES, "d2h_bytes_per_step": e2e_batch * 36,
               "steps": e2e_steps, "frames_per_step": e2e_batch,
               "api": "vpdq_b200_pdq_hash_frames_host (C ABI, pinned host buffers, 4 stages of 64 frames in flight)"}
        # the roof of that number: the same pinned buffer copied host -> device and nothing else, all ranks at once
        roof = copy_roof(torch, h_frames, dev, sync_all, max_over_ranks)
        e2e["copy_roof"] = {"GB/s_all_ranks": world * roof, "frames/s_all_ranks": world * roof * 1e9 / FRAME_BYTES,
                            "e2e_over_roof": e2e["value"] / (world * roof * 1e9 / FRAME_BYTES),
                            "what": "cudaMemcpyAsync of the e2e step's pinned input alone, all ranks concurrently"}
        # the e2e results must equal the device-resident ones
        chk_h, _ = dev_api.hash_frames(pool[0][:e2e_batch])
        assert torch.equal(chk_h.cpu(), h_hash), "e2e hashes differ from the device-resident path"
        metric, unit = "frame_hashes_per_sec", "frames/s"
        config = hash_config(args, B)
        config["parallelism"] = f"{world} x independent shards (no data-path collective)"
        dtype = "f32"

        del pool[1:]
        # the reference-shaped API: vpdq.VideoHasher.hash_frame(bytes) per frame, one hasher per video
        if not args.no_hasher_api:
            api = hasher_api_section(torch, dev_api, pool[0], local, sync_all, max_over_ranks, world, args)
            e2e["video_hasher_api"] = api
        if not args.no_hamming:  # every rank: the sharded search is a collective
            extra["hamming"] = {"sharded_all_pairs": sharded_pairs_section(torch, hdist, dev, rank, world, flush_l2,
                                                                          max_over_ranks, sync_all, args)}
        if rank == 0 and world == 1 and not args.no_hamming:
            extra["hamming"].update(hamming_section(torch, dev_api, dev, peak, peak_src, flush_l2, args))
        if not args.no_luma:
            luma = luma_section(torch, dev_api, dev, local, args, sync_all, max_over_ranks, world)
            if rank == 0:
                extra["luma_frames"] = luma
    else:
        # ---- BASELINE configs[3] as the headline: fixed 10 M-hash DB sharded over the ranks ----
        sampler = ClockSampler(local)
        launches0 = _ffi.kernel_launches()
        sp = sharded_pairs_section(torch, hdist, dev, rank, world, flush_l2, max_over_ranks, sync_all, args)
        clocks = sampler.stop()
        launches = _ffi.kernel_launches() - launches0
        ms_per_step = sp["ms_per_step"]
        value = sp["pair_comparisons_per_s"]
        algo = (sp["n_db"] // world + sp["n_query"]) * HASH_BYTES
        achieved = algo / (ms_per_step / 1e3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src, "kernel": "k_hamming_pairs",
                    "note": "all-pairs is POPC-issue bound, not HBM bound (compulsory bytes are tiny); see "
                            "DESIGN.md -- the HBM-bound regime is the streaming scan reported by --workload hash"}
        e2e = sp["e2e"]
        extra["hamming"] = {"sharded_all_pairs": sp}
        metric, unit = "pair_comparisons_per_sec", "pair-comparisons/s"
        config = hamming_config(args, world)
        config["parallelism"] = f"target DB sharded over {world} ranks, queries replicated, all_gather of bitmaps"
        dtype = "u64"

    if world > 1:
        hdist.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, pool[0] if args.workload == "hash" else None, torch)
    line.update(extra)
    emit(line)


def cpu_baseline(args, pool0, torch) -> dict:
    """The oracle port timed on this box's host cores (bounded sample), plus a parity spot-check."""
    cores = os.cpu_count() or 1
    if args.workload == "hash":
        n = 24 * cores
        passes = 16  # ~1 s on all cores = 15-20 s of CPU work: a bounded sample, long enough for a stable rate
        rate_all, secs = cpu_hash_rate(n, cores, repeats=passes)
        rate_1, _ = cpu_hash_rate(24, 1)
        import oracle
        from hydrus_video_deduplicator_b200 import device as dev_api

        sample = pool0[:9]
        gh, gq = dev_api.hash_frames(sample)
        rh, rq = oracle.pdq_hash_frames(sample.cpu().numpy(), nthreads=cores)
        ok = bool((gh.cpu().numpy() == rh).all() and (gq.cpu().numpy() == rq).all())
        return {"value": rate_all, "unit": "frames/s", "cores": cores, "kind": "port",
                "sample": f"{passes} passes over {n} synthetic 512x512 RGB24 frames on {cores} threads ({secs:.1f} s); "
                          f"1 thread: {rate_1:.1f} frames/s", "single_thread_value": rate_1,
                "parity_spot_check": "9/9 frames bit-exact vs oracle" if ok else "MISMATCH vs oracle"}
    rate, secs = cpu_pairs_rate(16384, cores)
    return {"value": rate, "unit": "pair-comparisons/s", "cores": cores, "kind": "port",
            "sample": f"16384 x 16384 hashes on {cores} threads ({secs:.1f} s)"}


def copy_roof(torch, h_pinned, dev, sync_all, max_over_ranks) -> float:
    """GB/s of this rank when every rank copies its pinned e2e input to the device and does nothing else."""
    d = torch.empty(h_pinned.shape, dtype=h_pinned.dtype, device=dev)
    for _ in range(2):
        d.copy_(h_pinned, non_blocking=True)
    sync_all()
    reps = 6
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h_pinned, non_blocking=True)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    return reps * h_pinned.numel() / dt / 1e9


def hasher_api_section(torch, dev_api, pool0, local, sync_all, max_over_ranks, world, args) -> dict:
    """Frames/s through the drop-in surface the reference binds to: vpdq.VideoHasher(1, 512, 512, n).hash_frame(bytes)
    per sampled frame, one hasher per video, finish() per video (vpdqpy/vpdqpy.py:113-119, dedup.py:346-352).  The
    frames are Python bytes objects in pageable memory, exactly what bytes(frame.planes[0]) hands over; copies into
    the pinned ring, uploads, kernels and result read-back are all inside the timed region.  Results are compared
    with the device-resident path."""
    import threading

    from hydrus_video_deduplicator_b200 import vpdq

    n_distinct = 96
    src = pool0[:n_distinct].cpu()
    want_h, want_q = dev_api.hash_frames(pool0[:n_distinct])
    want_h, want_q = want_h.cpu().numpy(), want_q.cpu().numpy()
    frames = [src[k].numpy().tobytes() for k in range(n_distinct)]

    def run_videos(n_videos: int, n_frames: int, first: int, check: bool):
        ok = True
        for v in range(n_videos):
            h = vpdq.VideoHasher(1, 512, 512, 0, device=local)
            base = (first + v * n_frames) % n_distinct
            for k in range(n_frames):
                h.hash_frame(frames[(base + k) % n_distinct])
            if check and v == 0:
                ph, all_h, all_q = h.finish(return_all=True)
                idx = [(base + k) % n_distinct for k in range(n_frames)]
                ok = ok and all_h == want_h[idx].tobytes() and all_q == want_q[idx].tolist()
            else:
                h.finish()
            h.close()
        return ok

    out = {"api": "vpdq.VideoHasher.hash_frame(bytes) x n, finish(); one hasher per video; frames are pageable Python "
                  "bytes", "n_gpus": world}
    from hydrus_video_deduplicator_b200 import _ffi

    run_videos(2, 40, 0, False)  # warm the service (arena allocation, threads)
    for name, n_frames, n_videos, threads in (("300_frame_videos", 300, 10, 1), ("10_frame_videos", 10, 300, 1),
                                              ("10_frame_videos_4_threads", 10, 300, 4)):
        results = [True] * threads
        sync_all()
        st0 = _ffi.service_stats(local)
        t_push = [0.0] * threads
        t0 = time.perf_counter()
        if threads == 1:
            results[0] = run_videos(n_videos, n_frames, 0, True)
        else:
            def work(t):
                results[t] = run_videos(n_videos // threads, n_frames, 7 * t, True)
            ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        dt = max_over_ranks(time.perf_counter() - t0)
        total = (n_videos // threads) * threads * n_frames
        st1 = _ffi.service_stats(local)
        ds = {k: st1[k] - st0[k] for k in st1}
        out[name] = {"frames/s": world * total / dt, "videos": (n_videos // threads) * threads, "frames_per_video": n_frames,
                     "caller_threads": threads, "bit_identical_to_device_path": bool(all(results)),
                     "service": {"frames_per_launch": ds["frames_launched"] / max(1, ds["launches"]),
                                 "frames_per_upload": ds["frames_launched"] / max(1, ds["upload_calls"]),
                                 "push_blocked_share": ds["push_blocked_ns"] / 1e9 / dt / threads,
                                 "finish_wait_share": ds["finish_wait_ns"] / 1e9 / dt / threads}}
    return out


def sharded_pairs_section(torch, hdist, dev, rank, world, flush_l2, max_over_ranks, sync_all, args) -> dict:
    """BASELINE configs[3]: ONE fixed 10 M-hash database (33 334 videos of 300 frames, SURVEY config 3's generator,
    same seed on every rank) sharded at video boundaries over the ranks; a replicated block of query hashes (the first
    rows of the DB, so every query has at least itself as a match); per step one all-pairs launch over this rank's
    shard and ONE all_gather of the per-query candidate bitmaps (NCCL when N > 1) inside the timed region.
    The N = 1 line is the same job on one GPU.  This is what the reference does file by file over the whole DB
    (db/vptree.py:865-902)."""
    from hydrus_video_deduplicator_b200 import _ffi

    n_db, fpv, nq = args.shard_db_hashes, 300, args.query_block
    db = device_hashes(torch, n_db, dev, seed=4242)
    import numpy as np

    offsets = np.arange(0, n_db + fpv, fpv, dtype=np.int64)
    offsets[-1] = n_db
    offsets = np.unique(offsets)
    bounds = hdist.shard_videos(offsets, world)
    f0, f1 = int(offsets[bounds[rank]]), int(offsets[bounds[rank + 1]])
    shard = db[f0:f1].clone()
    queries = db[:nq].clone()
    del db
    nt = f1 - f0
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    bitmap = torch.zeros(((nq + 31) // 32,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def search():
        count.zero_()
        bitmap.zero_()
        _ffi.check(_ffi.lib().vpdq_b200_hamming_pairs_dev(queries.data_ptr(), nq, shard.data_ptr(), nt, 31, 0,
                                                          bitmap.data_ptr(), None, 0, count.data_ptr(), stream))

    def step():
        search()
        return hdist.all_gather_bitmaps(bitmap)

    steps = max(1, min(args.steps, args.shard_steps))
    for _ in range(2):
        step()
    sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush_l2()
        ev[k][0].record()
        gathered = step()
        ev[k][1].record()
    torch.cuda.synchronize()
    ms_per_step = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / steps
    # the collective on its own
    cev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    sync_all()
    for a, b in cev:
        a.record()
        gathered = hdist.all_gather_bitmaps(bitmap)
        b.record()
    torch.cuda.synchronize()
    gather_us = max_over_ranks(statistics.median(a.elapsed_time(b) for a, b in cev) * 1e3)
    # matches: every rank counted its shard's; sum over ranks, and every query must be flagged (it matches itself)
    total = count.to(torch.float64).clone()
    if world > 1:
        torch.distributed.all_reduce(total)
    merged = hdist.or_reduce(gathered)
    flagged = int(sum(bin(int(v) & 0xFFFFFFFF).count("1") for v in merged[:1024].cpu().tolist()))
    # e2e: pinned host query block -> H2D -> search -> all_gather -> OR -> D2H of the merged bitmap
    hq = queries.cpu().pin_memory()
    hb = torch.empty_like(bitmap, device="cpu").pin_memory()
    sync_all()
    e2e_steps = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        queries.copy_(hq, non_blocking=True)
        g = step()
        hb.copy_(hdist.or_reduce(g), non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    return {"n_db": n_db, "frames_per_video": fpv, "n_query": nq, "n_gpus": world, "shard_hashes_this_rank": nt,
            "ms_per_step": ms_per_step, "pair_comparisons_per_s": float(n_db) * nq / (ms_per_step / 1e3),
            "all_gather_us": gather_us, "all_gather_bytes_per_rank": int(bitmap.numel() * 4),
            "all_gather_share_of_step": gather_us / 1e3 / ms_per_step,
            "matches_all_ranks": int(total.item()), "first_32768_queries_flagged": flagged,
            "collective": "nccl all_gather_into_tensor" if world > 1 else "none (single rank)",
            "generator": "SURVEY 8d config 3: popcount-128 words + 1 % planted near-duplicates at distances 0..40",
            "e2e": {"value": float(n_db) * nq * e2e_steps / e2e_s, "unit": "pair-comparisons/s",
                    "h2d_bytes_per_step": nq * HASH_BYTES, "d2h_bytes_per_step": int(hb.numel() * 4), "steps": e2e_steps}}


def luma_section(torch, dev_api, dev, local, args, sync_all, max_over_ranks, world) -> dict:
    """BASELINE's "512x512 luma frames" variant: 8-bit gray input, DEFINED as the RGB frame R=G=B=L (SURVEY.md note
    a-1; one byte per pixel crosses PCIe instead of three).  Same kernels as the RGB24 path (kx_systolic_jarosz<1>).
    Runs on every rank (whole-job rates, max over ranks)."""
    import ctypes as C

    from hydrus_video_deduplicator_b200 import _ffi

    n = min(args.batch, 2048)
    g = torch.Generator(device=dev).manual_seed(4242 + local)
    gray = torch.randint(0, 256, (n, 512, 512), dtype=torch.uint8, device=dev, generator=g)
    for _ in range(2):
        dev_api.hash_frames(gray)
    sync_all()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        h, q = dev_api.hash_frames(gray)
    b.record()
    torch.cuda.synchronize()
    dev_rate = world * 5 * n / (max_over_ranks(a.elapsed_time(b)) / 1e3)
    hg = torch.empty((n, 512, 512), dtype=torch.uint8, pin_memory=True)
    hg.copy_(gray)
    hh = torch.empty((n, 32), dtype=torch.uint8, pin_memory=True)
    hq = torch.empty((n,), dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize()

    def step():
        _ffi.check(_ffi.lib().vpdq_b200_pdq_hash_frames_host(C.c_void_p(hg.data_ptr()), 1, n, 512, 512,
                                                             C.c_void_p(hh.data_ptr()), C.c_void_p(hq.data_ptr()), local))

    step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    e2e_rate = world * 5 * n / max_over_ranks(time.perf_counter() - t0)
    rgb = gray[:4].unsqueeze(-1).expand(-1, -1, -1, 3).contiguous()
    same = bool(torch.equal(dev_api.hash_frames(rgb)[0], h[:4]) and torch.equal(hh[:4], h[:4].cpu()))
    return {"frames_per_s_device_resident": dev_rate, "frames_per_s_e2e": e2e_rate, "frames_per_rank": n, "n_gpus": world,
            "bytes_per_frame": 512 * 512, "equals_rgb_expansion": same, "kernels": "kx_systolic_jarosz<1> + k5_finalize"}


def hamming_section(torch, dev_api, dev, peak, peak_src, flush_l2, args) -> dict:
    """One GPU: streaming-scan GB/s (the HBM-bound regime) by resident query count, 1M x 1M all pairs, and both
    kernels again on a video-like database whose match density is that of a real library."""
    from hydrus_video_deduplicator_b200 import _ffi

    out = {}
    stream = torch.cuda.current_stream().cuda_stream

    def time_scan(db, offsets, q):
        n_db, nq, n_videos = db.shape[0], q.shape[0], offsets.numel() - 1
        qmask = torch.zeros((n_videos,), dtype=torch.int64, device=dev)

        def run():
            _ffi.check(_ffi.lib().vpdq_b200_hamming_scan_dev(db.data_ptr(), n_db, offsets.data_ptr(), n_videos,
                                                             q.data_ptr(), nq, 31, qmask.data_ptr(), None, stream))

        for _ in range(3):
            run()
        times = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        ms = statistics.mean(times)
        gbs = n_db * HASH_BYTES / (ms / 1e3) / 1e9
        return {"ms": ms, "GB/s": gbs, "frac_of_hbm_peak": gbs / peak, "pair_comparisons_per_s": n_db * nq / (ms / 1e3),
                "videos_hit": int((qmask != 0).sum())}

    n_db, fpv = args.scan_hashes, 300
    db = device_hashes(torch, n_db, dev, seed=9)
    offsets = torch.arange(0, n_db + 1, fpv, dtype=torch.int64, device=dev)
    if int(offsets[-1]) != n_db:
        offsets = torch.cat([offsets, torch.tensor([n_db], dtype=torch.int64, device=dev)])
    scan = {str(nq): time_scan(db, offsets, db[:nq].clone()) for nq in (1, 2, 4, 8, 16, 64)}
    out["scan"] = {"n_db": n_db, "db_bytes": n_db * HASH_BYTES, "frames_per_video": fpv, "by_n_query": scan,
                   "peak_GB/s": peak, "peak_source": peak_src, "algorithmic_bytes_per_hash": HASH_BYTES,
                   "generator": "SURVEY 8d config 3",
                   "note": "DB (320 MB) exceeds L2 (126 MB): every launch streams it from HBM"}
    del db

    def time_pairs(h, reps=2):
        n = h.shape[0]
        dev_api.hamming_pairs(h[: n // 8], h, 31, skip_diagonal=True, capacity=1, want_bitmap=True)
        times, found = [], 0
        for _ in range(reps):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            found, _, _ = dev_api.hamming_pairs(h, h, 31, skip_diagonal=True, capacity=1, want_bitmap=True)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        ms = statistics.mean(times)
        return {"n_q": n, "n_t": n, "ms": ms, "pair_comparisons_per_s": float(n) * n / (ms / 1e3), "matches": int(found)}

    n = args.pairs_hashes
    out["all_pairs"] = time_pairs(device_hashes(torch, n, dev, seed=10))
    out["all_pairs"]["bound"] = "POPC issue (DB is L2 resident; compulsory HBM bytes = 64 MB)"
    out["all_pairs"]["generator"] = "SURVEY 8d config 3"
    # the same kernels where matches are dense: 300-frame random-walk videos, 5 % duplicated videos
    vdb, voff = device_video_hashes(torch, n_db // fpv, fpv, dev, seed=11)
    vscan = {str(nq): time_scan(vdb, voff, vdb[1000 * fpv:1000 * fpv + nq].clone()) for nq in (1, 8, 64)}
    vpairs = time_pairs(vdb[: (n // fpv) * fpv].contiguous(), reps=1)
    out["video_like_db"] = {"what": "random-walk videos (adjacent frames <= 16 bits apart), 5 % duplicate videos",
                            "scan_n_db": int(vdb.shape[0]), "scan_by_n_query": vscan, "all_pairs": vpairs,
                            "matches_per_query_frame": vpairs["matches"] / max(1, vpairs["n_q"])}
    return out


def main() -> None:
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter) is sent to stderr, and the line is written to the real stdout at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=["hash", "hamming"], default="hash")
    ap.add_argument("--batch", type=int, default=8192, help="frames per step per GPU (hash workload)")
    ap.add_argument("--pool-batches", type=int, default=3, help="distinct device batches cycled through")
    ap.add_argument("--e2e-batch", type=int, default=2048)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--scan-hashes", type=int, default=10_000_000)
    ap.add_argument("--pairs-hashes", type=int, default=1 << 20)
    ap.add_argument("--shard-db-hashes", type=int, default=10_000_000, help="the FIXED database sharded over the ranks")
    ap.add_argument("--shard-steps", type=int, default=3)
    ap.add_argument("--query-block", type=int, default=262_144)
    ap.add_argument("--no-hasher-api", action="store_true")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-luma", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
