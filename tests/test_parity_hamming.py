"""GPU parity of the Hamming similarity path against the CPU oracle: frame-pair sets, per-video matched
counts, matchHash / is_similar values and search_file result lists -- all bit/integer exact."""
from __future__ import annotations

import numpy as np
import pytest

import oracle
from hydrus_video_deduplicator_b200 import hashing, search, vpdq
from hydrus_video_deduplicator_b200.vpdqpy import Vpdq, VpdqHash
from tests import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_dev():
    import torch

    return torch, torch.device("cuda", 0)


def _golden(golden_dir):
    return {p.name[:-4]: VpdqHash.from_string(p.read_text()) for p in sorted((golden_dir / "video_hashes").glob("*.txt"))}


def test_is_similar_on_golden_hashes(golden_dir):
    """The reference's own similarity test (test_vpdqpy.py:131-145) and benchmark loop
    (test_benchmark_vpdqpy.py:49-73), values checked against the oracle."""
    gold = _golden(golden_dir)
    for a, ha in gold.items():
        for b, hb in gold.items():
            similar, sim = Vpdq.is_similar(ha, hb)
            assert sim == oracle.match_hash(ha.bytes, hb.bytes), (a, b)
            assert hashing.get_phash_similarity(ha, hb) == sim
            if a != b:
                assert similar == (a.split("_")[0] == b.split("_")[0]), (a, b, sim)
    assert vpdq.matchHash(VpdqHash(), gold["S02_Sintel_720_10s_1MB.mp4"], 31) == 0.0
    assert search.calculate_distance(b"", b"") == 101


def test_match_hash_boundary_and_long_queries():
    rng = np.random.default_rng(17)
    base = synth.random_balanced_hashes(150, rng)  # > 64 query frames: several scan chunks
    tgt = np.stack([synth.flip_bits(h, int(d), rng) for h, d in zip(base, rng.integers(24, 40, size=150))])
    rng.shuffle(tgt)
    for tol in (0, 30, 31, 32):
        assert vpdq.matchHashBytes(base.tobytes(), tgt.tobytes(), tol) == oracle.match_hash(base, tgt, tol)
    assert vpdq.matchHashBytes(base.tobytes(), base.tobytes(), 0) == 100.0


def test_scan_matched_counts_vs_oracle(torch_dev):
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    vids, offsets = synth.synth_video_db(400, 0, seed=4)  # ragged: 0..11 frames per video, empty ones included
    db = np.frombuffer(b"".join(vids), np.uint8).reshape(-1, 32)
    d_db = torch.from_numpy(db.copy()).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    for qv in (3, 57, 200, 399):
        q = np.frombuffer(vids[qv], np.uint8).reshape(-1, 32)
        if len(q) == 0:
            continue
        qmask, tcount = device.hamming_scan(d_db, torch.from_numpy(q.copy()).to(dev), d_off, 31, reverse_counts=True)
        got = np.array([bin(int(m) & (2**64 - 1)).count("1") for m in qmask.cpu().numpy()], np.int32)
        ref = oracle.video_matched(q, db, offsets, 31)
        assert (got == ref).all()
        rev = np.array([oracle.lib().oracle_matched_frames(
            oracle._u8(db[offsets[v]:offsets[v + 1]].copy()), int(offsets[v + 1] - offsets[v]), oracle._u8(q), len(q), 31)
            if offsets[v + 1] > offsets[v] else 0 for v in range(400)], np.int32)
        assert (tcount.cpu().numpy() == rev).all()


def test_scan_frame_level_identity_offsets(torch_dev):
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    h = synth.synth_hashes(5000, seed=8, planted_frac=0.05)
    q = h[:64]
    qmask = device.hamming_scan(torch.from_numpy(h).to(dev), torch.from_numpy(q.copy()).to(dev), None, 31)
    got = {(i, j) for j, m in enumerate(qmask.cpu().numpy()) for i in range(64) if (int(m) >> i) & 1}
    ref = {(int(i), int(j)) for i, j in oracle.hamming_pairs(q, h, 31)}
    assert got == ref and len(ref) >= 64


@pytest.mark.parametrize("nq,nt", [(1, 1), (37, 5000), (4099, 3001), (2048, 1024)])
def test_pairs_set_equals_oracle(torch_dev, nq, nt):
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    pool = synth.synth_hashes(max(nq, nt), seed=nq + nt, planted_frac=0.05)
    q, t = pool[:nq], pool[:nt]
    n, pairs, bitmap = device.hamming_pairs(torch.from_numpy(q.copy()).to(dev), torch.from_numpy(t.copy()).to(dev), 31)
    ref = oracle.hamming_pairs(q, t, 31)
    assert n == len(ref)
    assert {tuple(p) for p in pairs.cpu().numpy().tolist()} == {tuple(p) for p in ref.tolist()}
    bits = np.unpackbits(bitmap.cpu().numpy().view(np.uint8), bitorder="little")[:nq]
    assert (np.flatnonzero(bits) == np.unique(ref[:, 0])).all()
    # self-join without the diagonal
    n2, pairs2, _ = device.hamming_pairs(torch.from_numpy(q.copy()).to(dev), torch.from_numpy(q.copy()).to(dev), 31,
                                         skip_diagonal=True)
    ref2 = oracle.hamming_pairs(q, q, 31)
    ref2 = ref2[ref2[:, 0] != ref2[:, 1]]
    assert n2 == len(ref2) and {tuple(p) for p in pairs2.cpu().numpy().tolist()} == {tuple(p) for p in ref2.tolist()}


def test_pairs_capacity_overflow_is_reported(torch_dev):
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    h = np.repeat(synth.synth_hashes(4, seed=1), 50, axis=0)  # 200 hashes, 4 groups of identical ones
    n, pairs, _ = device.hamming_pairs(torch.from_numpy(h).to(dev), torch.from_numpy(h.copy()).to(dev), 31, capacity=100)
    assert n == 4 * 50 * 50 and len(pairs) == 100


def test_search_file_equals_oracle_brute_force():
    vids, _ = synth.synth_video_db(120, 10, seed=6, dup_frac=0.3)
    ids = [1000 + 7 * k for k in range(len(vids))]
    index = search.HashIndex(ids, vids)
    for radius in (26, 51):  # thresholds 75 and 50 (dedup.py:455)
        for k in (0, 5, 17, 60, 119):
            got = index.search_file(ids[k], radius)
            ref = oracle.search_file(vids, k, radius)
            assert got[0] == (ids[k], 0)
            ref_ids = [(ids[v], d) for v, d in ref]
            self_hit = [(ids[k], oracle.calculate_distance(vids[k], vids[k]))] if vids[k] else []
            assert sorted(got) == sorted(set(ref_ids + self_hit))
    # acceptance-shaped run: number of undirected pairs at the CLI default threshold 50
    directed = index.find_potential_duplicates(50.0)
    ref_directed = sum(len(oracle.search_file(vids, k, 51)) - 1 for k in range(len(vids)))
    assert len(directed) == ref_directed
    index.close()


def test_acceptance_pair_count_on_gpu(golden_dir):
    gold = _golden(golden_dir)
    bbb = [h.bytes for n, h in gold.items() if n.startswith("S01_")]
    index = search.HashIndex(list(range(6)), bbb)
    assert len(index.find_potential_duplicates(50.0)) // 2 == 15  # test_main_vcr.py:64-66
    index.close()


def test_large_scale_properties(torch_dev):
    """1M-hash sizes: symmetry and planted-duplicate recall (properties, not the oracle)."""
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    n = 1 << 20
    g = torch.Generator(device=dev).manual_seed(5)
    db = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    # plant: hash 3k+1 := hash 3k with 20 flipped bits, for the first 1000 k
    src = db[0:3000:3].clone()
    flip = torch.zeros_like(src)
    flip[:, :5] = 0x0F  # 20 bits
    db[1:3001:3] = src ^ flip
    q = db[:4096]
    n1, p1, _ = device.hamming_pairs(q, db, 31, skip_diagonal=True)
    n2, p2, _ = device.hamming_pairs(db, q, 31, skip_diagonal=True)
    a = {tuple(x) for x in p1.cpu().numpy().tolist()}
    b = {(j, i) for i, j in p2.cpu().numpy().tolist()}
    assert n1 == n2 and a == b
    for k in range(0, 1000, 37):
        assert (3 * k, 3 * k + 1) in a and (3 * k + 1, 3 * k) in a


def test_baseline_sizes_vs_oracle_slabs(torch_dev):
    """BASELINE configs[2]/[3] sizes: a 10M-hash streaming scan and a 1M-target all-pairs slab, each checked
    against the oracle on as many rows as the CPU finishes in seconds."""
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    g = torch.Generator(device=dev).manual_seed(11)
    n = 10_000_000
    db = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    q = db[torch.randint(0, n, (64,), device=dev, generator=g)].clone()
    flips = torch.zeros((64, 32), dtype=torch.uint8, device=dev)
    flips[::2, :4] = 0xFF      # every other query sits at distance 32 from its source: just outside
    flips[1::2, :3] = 0xFF     # the rest at 24 ... 
    flips[1::2, 3] = 0x7F      # ... + 7 = 31: just inside
    q ^= flips
    off = torch.arange(0, n + 1, 300, dtype=torch.int64, device=dev)
    if int(off[-1]) != n:
        off = torch.cat([off, torch.tensor([n], dtype=torch.int64, device=dev)])
    qmask = device.hamming_scan(db, q, off, 31).cpu().numpy()
    db_h, q_h, off_h = db.cpu().numpy(), q.cpu().numpy(), off.cpu().numpy()
    ref = oracle.hamming_pairs(q_h, db_h, 31)
    want = np.zeros(len(off_h) - 1, np.uint64)
    for i, j in ref:
        want[np.searchsorted(off_h, j, side="right") - 1] |= np.uint64(1) << np.uint64(i)
    assert (qmask.view(np.uint64) == want).all()
    assert len(ref) == 32  # exactly the distance-31 half matches (random 256-bit words never come close)

    m = 1 << 20
    slab = db[:m]
    nq = 2048
    cnt, pairs, _ = device.hamming_pairs(q, slab, 31)
    assert cnt == len(oracle.hamming_pairs(q_h, db_h[:m], 31))
    planted = slab[:nq].clone()
    planted[:, 31] ^= 0x55  # distance 4 from rows 0..nq-1
    cnt, pairs, bitmap = device.hamming_pairs(planted, slab, 31)
    refp = oracle.hamming_pairs(planted.cpu().numpy(), db_h[:m], 31)
    assert cnt == len(refp) == nq
    assert {tuple(p) for p in pairs.cpu().numpy().tolist()} == {tuple(p) for p in refp.tolist()}


def test_video_matches_many_queries_vs_oracle(torch_dev):
    """vpdq_b200_hamming_scan_multi_dev + vpdq_b200_video_match_dev: MANY query videos (ragged, empty, > 64 frames)
    against a ragged database with empty videos, in one scan launch and one reduce launch: the dense matched counts
    equal the oracle's for every (query video, target video), the compact rows equal the oracle's search_file lists
    (matchHash numerator and calculate_distance, db/vptree.py:22-31)."""
    torch, dev = torch_dev
    from hydrus_video_deduplicator_b200 import device

    vids, offsets = synth.synth_video_db(300, 0, seed=8, dup_frac=0.4)
    rng = np.random.default_rng(3)
    long_vid = np.concatenate([np.frombuffer(vids[k], np.uint8).reshape(-1, 32) for k in (5, 9, 40, 77, 120, 201, 250)] * 4)
    vids[33] = long_vid.tobytes()  # a long video: several scan chunks
    vids[34] = long_vid[:64].tobytes()  # exactly one full chunk
    offsets = np.concatenate([[0], np.cumsum([len(v) // 32 for v in vids])]).astype(np.int64)
    assert len(vids[33]) // 32 > 64 and (np.diff(offsets) == 0).any()
    db = np.frombuffer(b"".join(vids), np.uint8).reshape(-1, 32)
    d_db = torch.from_numpy(db.copy()).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    q_ids = [33, 34, 0, 1, 2, 3, 150, 299] + [int(v) for v in np.flatnonzero(np.diff(offsets) == 0)[:2]]
    q = np.concatenate([np.frombuffer(vids[v], np.uint8).reshape(-1, 32) for v in q_ids])
    q_off = np.concatenate([[0], np.cumsum([len(vids[v]) // 32 for v in q_ids])])
    d_q = torch.from_numpy(q.copy()).to(dev)
    dense = device.video_matches(d_db, d_off, d_q, q_off, 31, dense=True).cpu().numpy()
    for k, v in enumerate(q_ids):
        qv = np.frombuffer(vids[v], np.uint8).reshape(-1, 32)
        want = oracle.video_matched(qv, db, offsets, 31) if len(qv) else np.zeros(len(vids), np.int32)
        assert (dense[k] == want).all(), v
    for radius in (26, 51, 0):
        rows = device.video_matches(d_db, d_off, d_q, q_off, 31, max_distance=radius).cpu().numpy()
        got = {}
        for qi, tv, m, d in rows.tolist():
            got.setdefault(qi, set()).add((tv, d))
            assert m == dense[qi, tv]
        for k, v in enumerate(q_ids):
            # the oracle's search_file lists the query itself as (v, 0) (vptree.py:866); as a stored video it scores
            # like any other: 100 % of its frames match themselves -> distance 1
            want = {(tv, 1 if tv == v else d) for tv, d in oracle.search_file(vids, v, radius if radius else 101)}
            assert got.get(k, set()) == {(tv, d) for tv, d in want if dense[k, tv] > 0}, (v, radius)


def test_db_search_long_queries_and_radius():
    """vpdq_b200_db_search / _search_radius: any number of query frames in one launch sequence (the popcounts on the
    device, one synchronisation) -- checked for 1, 64, 65 and 300+ query frames against the oracle."""
    import ctypes as C

    from hydrus_video_deduplicator_b200 import _ffi

    vids, offsets = synth.synth_video_db(500, 0, seed=21, dup_frac=0.3)
    db = np.frombuffer(b"".join(vids), np.uint8).reshape(-1, 32)
    index = search.HashIndex(list(range(100, 600)), vids)
    rng = np.random.default_rng(5)
    for n_q in (1, 64, 65, 333):
        rows = rng.integers(0, len(db), size=n_q)
        q = np.ascontiguousarray(db[rows])
        q[::3] = np.stack([synth.flip_bits(h, 30, rng) for h in q[::3]])
        got = index.matched_frames(q.tobytes())
        want = oracle.video_matched(q, db, offsets, 31)
        assert (got == want).all(), n_q
        for radius in (26, 51):
            out = np.zeros((len(vids), 4), np.int32)
            n = C.c_int64(0)
            _ffi.check(_ffi.lib().vpdq_b200_db_search_radius(index._db, q.tobytes(), n_q, 31, radius,
                                                             out.ctypes.data_as(C.c_void_p), len(vids), C.byref(n)))
            dist = (100 - (100 * want.astype(np.int64)) // n_q) + 1
            ref = {(int(v), int(want[v]), int(dist[v])) for v in np.flatnonzero((want > 0) & (dist <= radius))}
            assert {(r[1], r[2], r[3]) for r in out[: n.value].tolist()} == ref, (n_q, radius)
    index.close()
