"""reading the reference's sqlite tables and driving its search loop with the GPU index."""
from __future__ import annotations

import sqlite3

import pytest

from hydrus_video_deduplicator_b200 import dbio

SCHEMA = [  # verbatim shapes from the reference's DedupeDB.create_tables (DedupeDB.py:153-189)
    "CREATE TABLE IF NOT EXISTS files ( hash_id INTEGER PRIMARY KEY, file_hash BLOB_BYTES UNIQUE )",
    "CREATE TABLE IF NOT EXISTS shape_perceptual_hashes ( phash_id INTEGER PRIMARY KEY, phash BLOB_BYTES UNIQUE )",
    "CREATE TABLE IF NOT EXISTS shape_perceptual_hash_map ( phash_id INTEGER, hash_id INTEGER, PRIMARY KEY ( phash_id, hash_id ) )",
    "CREATE TABLE IF NOT EXISTS shape_search_cache ( hash_id INTEGER PRIMARY KEY, searched_distance INTEGER )",
]


def make_db(path, golden_dir):
    con = sqlite3.connect(str(path))
    for stmt in SCHEMA:
        con.execute(stmt)
    names = sorted(p.name[:-4] for p in (golden_dir / "video_hashes").glob("*.txt"))
    for k, name in enumerate(names, 1):
        phash = bytes.fromhex((golden_dir / "video_hashes" / f"{name}.txt").read_text().strip())
        con.execute("INSERT INTO files VALUES (?, ?)", (k, name.encode()))
        con.execute("INSERT OR IGNORE INTO shape_perceptual_hashes ( phash ) VALUES (?)", (phash,))
        (pid,) = con.execute("SELECT phash_id FROM shape_perceptual_hashes WHERE phash = ?", (phash,)).fetchone()
        con.execute("INSERT INTO shape_perceptual_hash_map VALUES (?, ?)", (pid, k))
        con.execute("INSERT INTO shape_search_cache VALUES (?, NULL)", (k,))
    con.commit()
    return con, names


def test_load_phashes_reads_the_reference_schema(tmp_path, golden_dir):
    con, names = make_db(tmp_path / "dedupe.sqlite", golden_dir)
    ids, phashes = dbio.load_phashes(con)
    assert ids == list(range(1, len(names) + 1))
    assert all(len(p) % 32 == 0 and len(p) > 0 for p in phashes)
    assert dbio.pending_searches(con, 51) == ids
    con.execute("UPDATE shape_search_cache SET searched_distance = 51 WHERE hash_id <= 4")
    assert dbio.pending_searches(con, 51) == ids[4:] and dbio.pending_searches(con, 52) == ids


@pytest.mark.gpu
def test_find_potential_duplicates_matches_the_acceptance_run(tmp_path, golden_dir):
    """6 Big Buck Bunny + 4 Sintel files at the CLI default threshold 50: C(6,2) + C(4,2) = 21 pairs (the
    reference's VCR run holds the 6 BBB files only: 15, tests/acceptance_tests/test_main_vcr.py:64-66)."""
    con, names = make_db(tmp_path / "dedupe.sqlite", golden_dir)
    marked = []
    n = dbio.find_potential_duplicates(con, 50.0, mark=lambda a, b: marked.append((a, b)))
    assert n == 15 + 6
    assert len(marked) == 2 * n
    assert all(a.split(b"_")[0] == b.split(b"_")[0] for a, b in marked)  # only same-group files pair up
    # incremental semantics: everything is now searched at radius 51 -> a second run does no work
    assert dbio.pending_searches(con, 51) == []
    assert dbio.find_potential_duplicates(con, 50.0) == 0
    # a stricter threshold (larger similarity) means a smaller radius: still nothing pending
    assert dbio.find_potential_duplicates(con, 75.0) == 0
