"""write the reference's vp-tree tables from the GPU hash index, so that a database whose
perceptual hashes were produced (or searched) here stays usable by the stock CLI.

The reference keeps its similarity index in (``db/DedupeDB.py:169-176``)

    shape_vptree ( phash_id INTEGER PRIMARY KEY, parent_id INTEGER, radius INTEGER, inner_id INTEGER,
                   inner_population INTEGER, outer_id INTEGER, outer_population INTEGER )
    shape_maintenance_branch_regen ( phash_id INTEGER PRIMARY KEY )

and builds it with ``VpTreeManager.regenerate_tree`` -> ``generate_branch`` (``db/vptree.py:285-420``): a
breadth-first split of the children of every node at the median of ``calculate_distance(node, child)``
(node = query, child = target; ``vptree.py:344-350``), ties going to the smaller side (``:373-383``), each side's
root chosen by ``pop_best_root_node`` (``:422-495``, a randomised balance heuristic).  Here the same construction
runs with ONE streaming CUDA pass per node for its distances to all stored hashes (``HashIndex.distances``)
instead of one Python->native ``matchHashBytes`` hop per (node, child) pair.

What is identical to the reference: the table layout, the split rule (median index ``len // 2``, strict ``<`` /
``>`` partition, radius decremented when the tie group goes outside), populations, parent links, and the
clearing of ``shape_maintenance_branch_regen``.  What is not: ``pop_best_root_node``'s choice -- any member of a
side is a valid root, the choice only affects balance -- uses the reference's score (ratio of the split, then
standard deviation of the views) on at most ``max_viewpoints`` candidates (reference: 256) because every
viewpoint costs a pass here; ``rng`` makes it reproducible.  The reference's own randomness (``random.sample``)
means no two of its trees are alike either.
"""
from __future__ import annotations

import collections
import random
import sqlite3
from pathlib import Path
from typing import Callable, Sequence

import numpy as np

from .dbio import connect

# distances from one perceptual hash (the query) to each of a list of stored rows (the targets):
# distance_fn(query_phash, target_rows) -> int array, values in [1, 101] (calculate_distance, vptree.py:29-31)
DistanceFn = Callable[[bytes, np.ndarray], np.ndarray]

MAX_SAMPLE = 64  # vptree.py:428


def gpu_distance_fn(phashes: Sequence[bytes], *, device: int | None = None) -> tuple[DistanceFn, Callable[[], None]]:
    """One resident HashIndex over all rows; each call is one streaming pass over the stored hashes."""
    from .search import HashIndex

    index = HashIndex(list(range(len(phashes))), phashes, device=device)

    def fn(query: bytes, rows: np.ndarray) -> np.ndarray:
        return index.distances(query)[rows]

    return fn, index.close


def _score_viewpoint(views: np.ndarray) -> tuple[int, float]:
    """vptree.py:456-489: how evenly the median splits the sample, then the spread of the views."""
    views = np.sort(views)
    radius = views[len(views) // 2]
    num_left = int((views < radius).sum())
    num_radius = int((views == radius).sum())
    num_right = int((views > radius).sum())
    if num_left <= num_right:
        num_left += num_radius
    else:
        num_right += num_radius
    smaller, larger = min(num_left, num_right), max(num_left, num_right)
    ratio_score = int((smaller / larger) * MAX_SAMPLE / 2)
    return ratio_score, float(views.std())


def _pop_best_root(rows: list[int], phashes: Sequence[bytes], distance_fn: DistanceFn, rng: random.Random,
                   max_viewpoints: int) -> int:
    """Remove and return the member of ``rows`` that will root this side (vptree.py:422-495)."""
    if len(rows) == 1:
        return rows.pop()
    viewpoints = rng.sample(rows, max_viewpoints) if len(rows) > max_viewpoints else list(rows)
    sample = rng.sample(rows, MAX_SAMPLE) if len(rows) > MAX_SAMPLE else list(rows)
    scores = []
    for v in viewpoints:
        others = np.asarray([s for s in sample if s != v], dtype=np.int64)
        if len(others) == 0:
            scores.append((0, 0.0, v))
            continue
        ratio_score, sd = _score_viewpoint(distance_fn(phashes[v], others))
        scores.append((ratio_score, sd, v))
    scores.sort()
    root = scores[-1][2]
    rows.remove(root)
    return root


def build_tree(phash_ids: Sequence[int], phashes: Sequence[bytes], distance_fn: DistanceFn, *,
               rng: random.Random | None = None, max_viewpoints: int = 8
               ) -> list[tuple[int, int | None, int | None, int | None, int, int | None, int]]:
    """-> rows (phash_id, parent_id, radius, inner_id, inner_population, outer_id, outer_population) of
    ``shape_vptree`` for all given hashes (generate_branch, vptree.py:315-420)."""
    n = len(phash_ids)
    if n == 0:
        return []
    rng = rng or random.Random(0)
    all_rows = list(range(n))
    root = _pop_best_root(all_rows, phashes, distance_fn, rng, max_viewpoints)
    out = []
    queue = collections.deque([(None, root, all_rows)])
    while queue:
        parent, node, children = queue.popleft()
        if not children:
            out.append((phash_ids[node], parent, None, None, 0, None, 0))
            continue
        rows = np.asarray(children, dtype=np.int64)
        dist = np.asarray(distance_fn(phashes[node], rows), dtype=np.int64)
        order = np.lexsort((np.asarray([phash_ids[r] for r in children]), dist))  # sorted((distance, child_id, ...))
        median_radius = int(dist[order[len(children) // 2]])
        inner = [int(rows[i]) for i in order if dist[i] < median_radius]
        tie = [int(rows[i]) for i in order if dist[i] == median_radius]
        outer = [int(rows[i]) for i in order if dist[i] > median_radius]
        if len(inner) <= len(outer):
            radius = median_radius
            inner.extend(tie)
        else:
            radius = median_radius - 1
            outer.extend(tie)
        inner_population, outer_population = len(inner), len(outer)
        inner_root = _pop_best_root(inner, phashes, distance_fn, rng, max_viewpoints)
        outer_root = _pop_best_root(outer, phashes, distance_fn, rng, max_viewpoints) if outer else None
        out.append((phash_ids[node], parent, radius, phash_ids[inner_root], inner_population,
                    None if outer_root is None else phash_ids[outer_root], outer_population))
        queue.append((phash_ids[node], inner_root, inner))
        if outer_root is not None:
            queue.append((phash_ids[node], outer_root, outer))
    return out


def regenerate_tree(db: str | Path | sqlite3.Connection, *, distance_fn: DistanceFn | None = None,
                    device: int | None = None, rng: random.Random | None = None, max_viewpoints: int = 8) -> int:
    """``VpTreeManager.regenerate_tree`` (vptree.py:285-313) on the GPU: purge orphans, rebuild ``shape_vptree``
    from every row of ``shape_perceptual_hashes``, clear ``shape_maintenance_branch_regen``.  Returns the number
    of nodes written.  ``distance_fn`` defaults to the resident GPU index."""
    con = connect(db)
    con.execute("DELETE FROM shape_perceptual_hash_map WHERE hash_id NOT IN ( SELECT hash_id FROM files )")
    con.execute("CREATE TABLE IF NOT EXISTS shape_vptree ( phash_id INTEGER PRIMARY KEY, parent_id INTEGER, "
                "radius INTEGER, inner_id INTEGER, inner_population INTEGER, outer_id INTEGER, "
                "outer_population INTEGER )")
    con.execute("CREATE TABLE IF NOT EXISTS shape_maintenance_branch_regen ( phash_id INTEGER PRIMARY KEY )")
    con.execute("DELETE FROM shape_vptree;")
    nodes = con.execute("SELECT phash_id, phash FROM shape_perceptual_hashes ORDER BY phash_id;").fetchall()
    ids, phashes = [int(r[0]) for r in nodes], [bytes(r[1]) for r in nodes]
    close = None
    if distance_fn is None and ids:
        distance_fn, close = gpu_distance_fn(phashes, device=device)
    try:
        rows = build_tree(ids, phashes, distance_fn, rng=rng, max_viewpoints=max_viewpoints)
    finally:
        if close is not None:
            close()
    con.executemany(
        "INSERT OR REPLACE INTO shape_vptree ( phash_id, parent_id, radius, inner_id, inner_population, outer_id, "
        "outer_population ) VALUES ( ?, ?, ?, ?, ?, ?, ? );", rows)
    con.execute("DELETE FROM shape_maintenance_branch_regen;")
    con.commit()
    return len(rows)


def search_tree(db: str | Path | sqlite3.Connection, search_phash: bytes, search_radius: int,
                distance: Callable[[bytes, bytes], int]) -> dict[int, int]:
    """The reference's traversal of the tables (``search_perceptual_hashes``, vptree.py:707-777), restated so that a
    written tree can be exercised without the reference's native module: {phash_id: distance} of every visited
    node within the radius.  ``distance(search_phash, node_phash)`` = ``calculate_distance``."""
    con = connect(db)
    nodes = {int(r[0]): (bytes(r[1]), r[2], r[3], r[4]) for r in con.execute(
        "SELECT phash_id, phash, radius, inner_id, outer_id FROM shape_perceptual_hashes "
        "NATURAL JOIN shape_vptree;")}
    root = con.execute("SELECT phash_id FROM shape_vptree WHERE parent_id IS NULL;").fetchone()
    found: dict[int, int] = {}
    if root is None:
        return found
    next_potentials = [int(root[0])]
    while next_potentials:
        current, next_potentials = next_potentials, []
        for node_id in current:
            if node_id not in nodes:
                continue
            node_phash, node_radius, inner_id, outer_id = nodes[node_id]
            d = distance(search_phash, node_phash)
            if d <= search_radius:
                found[node_id] = min(d, found.get(node_id, d))
            if node_radius is not None:
                if inner_id is not None and not d > (node_radius + search_radius):
                    next_potentials.append(int(inner_id))
                if outer_id is not None and not (d + search_radius) <= node_radius:
                    next_potentials.append(int(outer_id))
    return found
