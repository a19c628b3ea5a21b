"""Multi-GPU layer: one process per GPU (torchrun), torch.distributed for the plumbing (NCCL over NVLink on
the box, gloo in the CPU tests).

What shards:
  hashing     independent frames/videos -> round-robin over ranks, NO collective on the data path; one
              all_gather of the 32-byte hashes afterwards if a caller wants the whole table everywhere.
  similarity  the TARGET database is sharded row-wise AT VIDEO BOUNDARIES (so every (query frame, target
              video) predicate is decided on one GPU); queries are replicated.  The only exchange step is
              an all_gather of the fixed-size per-query candidate bitmaps and of the (small) per-video
              match masks / pair lists.

The host-side logic here (partitioning, padding, merging) is device-agnostic and is what the world_size-2
gloo tests cover; the compute in between is the CUDA scan / pairs kernels.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank: int) -> int | None:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers
    (first-touch placement) and the copy-issuing threads are local to the GPU's PCIe root.  With 8 ranks on a
    two-socket host this is what keeps every H2D stream off the inter-socket link.  Best effort: returns the
    node, or None when the topology cannot be read (containers without /sys, single-node hosts, ...)."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def init(backend: str | None = None, *, numa_bind: bool = True) -> tuple[int, int, int]:
    """-> (rank, world_size, local_rank).  Safe to call when not launched by torchrun (world 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        if numa_bind and world > 1 and os.environ.get("VPDQ_B200_NUMA_BIND", "1") != "0":
            bind_to_gpu_numa_node(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


# ---------------------------------------------------------------------------------------------------
# partitioning
# ---------------------------------------------------------------------------------------------------
def round_robin(n_items: int, world: int, r: int) -> np.ndarray:
    """indices of the items (videos to hash) owned by rank r"""
    return np.arange(r, n_items, world, dtype=np.int64)


def shard_videos(offsets: np.ndarray, world: int) -> np.ndarray:
    """Split a CSR database into `world` contiguous video ranges of near-equal FRAME counts, cutting only at
    video boundaries.  -> bounds int64[world+1] (video indices): rank r owns videos bounds[r]..bounds[r+1]-1."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n_videos = len(offsets) - 1
    total = int(offsets[-1])
    bounds = np.zeros(world + 1, dtype=np.int64)
    for r in range(1, world):
        target = (total * r) // world
        # first video whose start is >= target, never moving backwards
        v = int(np.searchsorted(offsets[:-1], target, side="left"))
        bounds[r] = min(max(v, bounds[r - 1]), n_videos)
    bounds[world] = n_videos
    return bounds


def local_shard(db: np.ndarray, offsets: np.ndarray, world: int, r: int):
    """-> (db_shard [n_local, 32], offsets_shard rebased to 0, first_video) for rank r"""
    bounds = shard_videos(offsets, world)
    v0, v1 = int(bounds[r]), int(bounds[r + 1])
    f0, f1 = int(offsets[v0]), int(offsets[v1])
    return db[f0:f1], (np.asarray(offsets[v0:v1 + 1]) - f0).astype(np.int64), v0


# ---------------------------------------------------------------------------------------------------
# collectives
# ---------------------------------------------------------------------------------------------------
def all_gather_varlen(t: torch.Tensor) -> list[torch.Tensor]:
    """all_gather of 1-D/2-D tensors whose first dimension differs per rank (padded exchange)."""
    if not dist.is_initialized():
        return [t]
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[:s] for o, s in zip(out, sizes)]


def all_gather_bitmaps(bitmap: torch.Tensor) -> torch.Tensor:
    """[words] int32 per rank -> [world, words]: the "all-gather of candidate bitmaps"; OR over dim 0 gives
    the global any-match bitmap."""
    if not dist.is_initialized():
        return bitmap.unsqueeze(0)
    world = dist.get_world_size()
    flat = torch.empty((world * bitmap.numel(),), dtype=bitmap.dtype, device=bitmap.device)
    dist.all_gather_into_tensor(flat, bitmap.contiguous().view(-1))
    return flat.view((world,) + tuple(bitmap.shape))


def or_reduce(gathered: torch.Tensor) -> torch.Tensor:
    out = gathered[0].clone()
    for k in range(1, gathered.shape[0]):
        out |= gathered[k]
    return out


def merge_video_masks(local_masks: torch.Tensor) -> torch.Tensor:
    """Per-rank per-video match masks (shards are contiguous video ranges in rank order) -> the full
    [n_videos] vector on every rank."""
    return torch.cat(all_gather_varlen(local_masks))


def merge_pairs(local_pairs: torch.Tensor, first_target: int) -> torch.Tensor:
    """Per-rank (query, local target) pairs -> global pair list on every rank (targets rebased by the
    shard's first frame)."""
    if local_pairs.numel():
        local_pairs = local_pairs.clone()
        local_pairs[:, 1] += first_target
    return torch.cat(all_gather_varlen(local_pairs))


# ---------------------------------------------------------------------------------------------------
# sharded similarity search (CUDA compute + the collectives above)
# ---------------------------------------------------------------------------------------------------
class ShardedIndex:
    """Each rank holds its contiguous slice of the hash DB in HBM; a search scans every shard in parallel
    and all-gathers the per-video masks."""

    def __init__(self, db: np.ndarray, offsets: np.ndarray, device: torch.device | None = None):
        self.world, self.rank = world_size(), rank()
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        shard, off, self.first_video = local_shard(np.asarray(db).reshape(-1, 32), offsets, self.world, self.rank)
        self.first_frame = int(offsets[self.first_video])
        self.n_videos = len(offsets) - 1
        self.frames_per_video = torch.from_numpy(np.diff(np.asarray(offsets, dtype=np.int64)))
        self.d_db = torch.from_numpy(np.ascontiguousarray(shard)).to(self.device)
        self.d_off = torch.from_numpy(off).to(self.device)

    def matched_frames(self, query: np.ndarray | torch.Tensor, tolerance: int = 31) -> torch.Tensor:
        """[n_videos] int32 on every rank: # query frames with a match in each video (any number of query frames:
        one scan launch over the 64-frame chunks + the popcounts on the device, then the all_gather)."""
        from . import device as dev_api

        q = torch.as_tensor(query).reshape(-1, 32).to(self.device)
        n_local = self.d_off.numel() - 1
        if n_local and self.d_db.shape[0] and q.shape[0]:
            total = dev_api.video_matches(self.d_db, self.d_off, q, [0, q.shape[0]], tolerance, dense=True)[0]
        else:
            total = torch.zeros((n_local,), dtype=torch.int32, device=self.device)
        return merge_video_masks(total)

    def candidate_pairs(self, queries: torch.Tensor, tolerance: int = 31, capacity: int = 1 << 20):
        """All (query, global target frame) matches + the OR-ed candidate bitmap, on every rank."""
        from . import device as dev_api

        q = queries.reshape(-1, 32).to(self.device)
        n, pairs, bitmap = dev_api.hamming_pairs(q, self.d_db, tolerance, capacity=capacity)
        if n > capacity:
            raise OverflowError(f"{n} pairs on rank {self.rank} exceed capacity {capacity}")
        return merge_pairs(pairs, self.first_frame), or_reduce(all_gather_bitmaps(bitmap))
