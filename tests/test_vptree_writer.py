"""the vp-tree tables written from the GPU index obey the reference's invariants
(generate_branch, vptree.py:315-420) and can be walked by its traversal (vptree.py:707-777)."""
from __future__ import annotations

import random
import sqlite3

import numpy as np
import pytest

import oracle
from hydrus_video_deduplicator_b200 import vptree_writer
from tests import synth
from tests.test_dbio import SCHEMA

VPTREE_SCHEMA = [  # DedupeDB.py:169-176
    "CREATE TABLE IF NOT EXISTS shape_vptree ( phash_id INTEGER PRIMARY KEY, parent_id INTEGER, radius INTEGER, inner_id INTEGER, inner_population INTEGER, outer_id INTEGER, outer_population INTEGER )",
    "CREATE TABLE IF NOT EXISTS shape_maintenance_branch_regen ( phash_id INTEGER PRIMARY KEY )",
]


def make_db(n_videos=60, frames_per_video=6, seed=5):
    con = sqlite3.connect(":memory:")
    for stmt in SCHEMA + VPTREE_SCHEMA:
        con.execute(stmt)
    vids, _offsets = synth.synth_video_db(n_videos, frames_per_video, seed=seed, dup_frac=0.3)
    for k, phash in enumerate(vids, 1):
        con.execute("INSERT INTO files VALUES (?, ?)", (k, b"file%d" % k))
        con.execute("INSERT OR IGNORE INTO shape_perceptual_hashes ( phash ) VALUES (?)", (phash,))
        (pid,) = con.execute("SELECT phash_id FROM shape_perceptual_hashes WHERE phash = ?", (phash,)).fetchone()
        con.execute("INSERT OR IGNORE INTO shape_perceptual_hash_map VALUES (?, ?)", (pid, k))
    con.execute("INSERT INTO shape_maintenance_branch_regen VALUES (3)")
    con.commit()
    return con


def oracle_distance_fn(phashes):
    def fn(query, rows):
        return np.asarray([oracle.calculate_distance(query, phashes[int(r)]) for r in rows], dtype=np.int64)
    return fn


def check_invariants(con):
    phash = {int(r[0]): bytes(r[1]) for r in con.execute("SELECT phash_id, phash FROM shape_perceptual_hashes")}
    tree = {int(r[0]): r[1:] for r in con.execute(
        "SELECT phash_id, parent_id, radius, inner_id, inner_population, outer_id, outer_population FROM shape_vptree")}
    assert set(tree) == set(phash), "every perceptual hash is a node exactly once"
    roots = [n for n, row in tree.items() if row[0] is None]
    assert len(roots) == 1
    assert con.execute("SELECT COUNT(*) FROM shape_maintenance_branch_regen").fetchone()[0] == 0

    def subtree(n):
        out, stack = [], [n]
        while stack:
            x = stack.pop()
            out.append(x)
            _, _, inner, _, outer, _ = tree[x]
            stack.extend(c for c in (inner, outer) if c is not None)
        return out

    seen = subtree(roots[0])
    assert sorted(seen) == sorted(tree), "the tree reaches every node once"
    for n, (parent, radius, inner, inner_pop, outer, outer_pop) in tree.items():
        for child in (inner, outer):
            if child is not None:
                assert tree[child][0] == n, "parent links"
        ins = subtree(inner) if inner is not None else []
        outs = subtree(outer) if outer is not None else []
        assert (len(ins), len(outs)) == (inner_pop, outer_pop)
        if radius is None:
            assert not ins and not outs
            continue
        assert all(oracle.calculate_distance(phash[n], phash[x]) <= radius for x in ins)
        assert all(oracle.calculate_distance(phash[n], phash[x]) > radius for x in outs)
    return phash, tree


def test_tree_built_with_injected_distances_obeys_the_reference_invariants():
    con = make_db()
    ids = [int(r[0]) for r in con.execute("SELECT phash_id FROM shape_perceptual_hashes ORDER BY phash_id")]
    phashes = [bytes(r[0]) for r in con.execute("SELECT phash FROM shape_perceptual_hashes ORDER BY phash_id")]
    n = vptree_writer.regenerate_tree(con, distance_fn=oracle_distance_fn(phashes), rng=random.Random(7))
    assert n == len(ids)
    phash, _ = check_invariants(con)
    # the reference's traversal over the written tables: everything it reports is truly within the radius, at
    # the true distance (the walk itself is lossy: the distance is not a metric)
    for q in list(phash)[:12]:
        found = vptree_writer.search_tree(con, phash[q], 51, oracle.calculate_distance)
        brute = {x: oracle.calculate_distance(phash[q], p) for x, p in phash.items()}
        assert all(brute[x] == d and d <= 51 for x, d in found.items())
        assert q in found and found[q] == 1  # a stored hash finds itself (100 % match -> distance 1)


@pytest.mark.parametrize("n_videos,frames_per_video,seed", [(2, 3, 1), (3, 0, 2), (41, 0, 3), (97, 2, 4)])
def test_ragged_and_empty_hashes(n_videos, frames_per_video, seed):
    """Videos of 0..11 frames (frames_per_video = 0 draws ragged lengths, some empty: an empty hash is similar to
    nothing, distance 101, DedupeDB.py:555-557), tiny tables, duplicates collapsing onto one phash row."""
    con = make_db(n_videos=n_videos, frames_per_video=frames_per_video, seed=seed)
    phashes = [bytes(r[0]) for r in con.execute("SELECT phash FROM shape_perceptual_hashes ORDER BY phash_id")]
    n = vptree_writer.regenerate_tree(con, distance_fn=oracle_distance_fn(phashes), rng=random.Random(seed))
    assert n == len(phashes)
    check_invariants(con)


def test_degenerate_tables():
    con = sqlite3.connect(":memory:")
    for stmt in SCHEMA + VPTREE_SCHEMA:
        con.execute(stmt)
    assert vptree_writer.regenerate_tree(con, distance_fn=lambda q, rows: np.zeros(len(rows), np.int64)) == 0
    con.execute("INSERT INTO files VALUES (1, x'00')")
    con.execute("INSERT INTO shape_perceptual_hashes ( phash ) VALUES (?)", (bytes(64),))
    con.execute("INSERT INTO shape_perceptual_hash_map VALUES (1, 1)")
    assert vptree_writer.regenerate_tree(con, distance_fn=lambda q, rows: np.zeros(len(rows), np.int64)) == 1
    assert con.execute("SELECT * FROM shape_vptree").fetchall() == [(1, None, None, None, 0, None, 0)]


@pytest.mark.gpu
def test_gpu_distances_build_the_same_tree_as_the_oracle():
    con_a, con_b = make_db(n_videos=150, seed=9), make_db(n_videos=150, seed=9)
    phashes = [bytes(r[0]) for r in con_a.execute("SELECT phash FROM shape_perceptual_hashes ORDER BY phash_id")]
    vptree_writer.regenerate_tree(con_a, rng=random.Random(3))  # GPU index (default)
    vptree_writer.regenerate_tree(con_b, distance_fn=oracle_distance_fn(phashes), rng=random.Random(3))
    q = "SELECT * FROM shape_vptree ORDER BY phash_id"
    assert con_a.execute(q).fetchall() == con_b.execute(q).fetchall()
    check_invariants(con_a)
