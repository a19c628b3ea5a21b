"""the reference's sqlite wire format <-> the GPU hash index.

Reads the tables the reference keeps its perceptual hashes in (``db/DedupeDB.py:159-180``):

    files                     ( hash_id INTEGER PRIMARY KEY, file_hash BLOB_BYTES UNIQUE )
    shape_perceptual_hashes   ( phash_id INTEGER PRIMARY KEY, phash BLOB_BYTES UNIQUE )      n x 32 bytes
    shape_perceptual_hash_map ( phash_id INTEGER, hash_id INTEGER, PRIMARY KEY (phash_id, hash_id) )
    shape_search_cache        ( hash_id INTEGER PRIMARY KEY, searched_distance INTEGER )

and drives the same search loop as ``HydrusVideoDeduplicator.find_potential_duplicates`` (``dedup.py:445-502``)
with ``HashIndex.search_file`` in place of ``VpTreeManager.search_file`` -- including the incremental
``searched_distance`` bookkeeping, so a database produced by the stock CLI can be searched here and vice versa.
Only sqlite3 from the standard library is used; nothing here touches the vp-tree tables.
"""
from __future__ import annotations

import sqlite3
from pathlib import Path
from typing import Callable

from .search import HashIndex, fix_vpdq_similarity


def connect(db: str | Path | sqlite3.Connection) -> sqlite3.Connection:
    return db if isinstance(db, sqlite3.Connection) else sqlite3.connect(str(db))


def load_phashes(db: str | Path | sqlite3.Connection) -> tuple[list[int], list[bytes]]:
    """-> (hash_ids, phash blobs) of every file that has a perceptual hash, ordered by hash_id."""
    con = connect(db)
    rows = con.execute(
        "SELECT m.hash_id, p.phash FROM shape_perceptual_hash_map AS m "
        "JOIN shape_perceptual_hashes AS p USING ( phash_id ) ORDER BY m.hash_id;"
    ).fetchall()
    return [int(r[0]) for r in rows], [bytes(r[1]) for r in rows]


def load_hash_index(db: str | Path | sqlite3.Connection, *, device: int | None = None) -> HashIndex:
    """The whole phash table, resident on one GPU (one H2D copy of 32 bytes per frame)."""
    ids, phashes = load_phashes(db)
    return HashIndex(ids, phashes, device=device)


def pending_searches(con: sqlite3.Connection, search_threshold: int) -> list[int]:
    """Files never searched, or searched at a smaller radius (dedup.py:458-461)."""
    rows = con.execute(
        "SELECT hash_id FROM shape_search_cache WHERE searched_distance is NULL or searched_distance < :threshold",
        {"threshold": search_threshold},
    ).fetchall()
    return [int(r[0]) for r in rows]


def find_potential_duplicates(db: str | Path | sqlite3.Connection, threshold: float = 50.0, *,
                              mark: Callable[[bytes, bytes], None] | None = None, index: HashIndex | None = None,
                              device: int | None = None, commit_every: int = 64) -> int:
    """dedup.py:445-502 with the GPU index: returns the number of similar file pairs found (directed hits // 2).
    ``mark(file_hash_a, file_hash_b)`` stands where the reference POSTs the relationship to Hydrus
    (``mark_videos_as_duplicates``, dedup.py:477-482)."""
    con = connect(db)
    search_threshold = fix_vpdq_similarity(threshold)
    assert search_threshold > 0 and isinstance(search_threshold, int)
    own_index = index is None
    if index is None:
        index = load_hash_index(con, device=device)
    file_hash = dict(con.execute("SELECT hash_id, file_hash FROM files;").fetchall())
    num_similar_pairs = 0
    try:
        for n, hash_id in enumerate(pending_searches(con, search_threshold), 1):
            if hash_id in index._row:  # a file without a phash row cannot be searched (nor found)
                for similar_hash_id, _distance in index.search_file(hash_id, search_threshold):
                    if hash_id != similar_hash_id:
                        if mark is not None:
                            mark(file_hash.get(hash_id), file_hash.get(similar_hash_id))
                        num_similar_pairs += 1
            con.execute("UPDATE shape_search_cache SET searched_distance = ? WHERE hash_id = ?;",
                        (search_threshold, hash_id))
            if n % commit_every == 0:
                con.commit()
        con.commit()
    finally:
        if own_index:
            index.close()
    return num_similar_pairs // 2
