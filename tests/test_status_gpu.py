"""GPU parity tests for the status sweep diff (K2) through the C-ABI against the oracle."""
import json
import os

import numpy as np
import pytest

import np_restatement as npr
import oracle
import rpk

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_engine():
    return rpk.Engine(1)


def test_device_xxh64_matches_golden_vectors():
    """Every committed XXH64 KAT (lengths 0..255) through the kernel's hash column."""
    g = json.load(open(os.path.join(GOLD, "xxh64_kat.json")))
    vecs = [v for v in g["vectors"] if v["seed"] == 0 and len(v["hex"]) // 2 <= 255]
    for stride in (256,):
        recs = np.zeros((len(vecs), stride), np.uint8)
        for i, v in enumerate(vecs):
            d = bytes.fromhex(v["hex"])
            recs[i, 0] = len(d)
            recs[i, 1 : 1 + len(d)] = np.frombuffer(d, np.uint8)
        eng = make_engine()
        idx, hashes = eng.status_diff(recs, want_hashes=True)
        eng.close()
        for i, v in enumerate(vecs):
            assert format(int(hashes[i]), "016x") == v["xxh64"], (i, v["hex"][:32])
        assert len(idx) == len(vecs)


@pytest.mark.parametrize("stride", [16, 32, 48, 64, 128, 256])
def test_hash_column_all_lengths(stride):
    rng = np.random.default_rng(stride)
    N = 3000
    recs = rng.integers(0, 256, (N, stride), dtype=np.uint8)
    recs[:, 0] = rng.integers(0, stride, N)  # len in [0, stride-1]
    recs[:stride, 0] = np.arange(stride)
    eng = make_engine()
    _, hashes = eng.status_diff(recs, want_hashes=True)
    eng.close()
    assert np.array_equal(hashes, oracle.record_hashes(recs))


@pytest.mark.parametrize("N", [1, 255, 1024, 1025, 100_000])
def test_sweeps_match_reference_predicate(N):
    """BASELINE config 3's reconcile: N tracked pods, sweeps with f = 0, 1 %, 10 %, 100 % mutation; the
    changed list must equal the reference's string/bool predicate (oracle) exactly, in ascending order."""
    eng = make_engine()
    tab = oracle.StatusTable(N)
    for sweep, frac in enumerate([0.0, 0.0, 0.01, 0.10, 1.0, 0.0]):
        recs = rpk.synth.make_status_records(N, sweep=sweep, mutate_frac=frac)
        got, hashes = eng.status_diff(recs, want_hashes=True)
        want = tab.diff(recs)
        assert np.array_equal(got, want), (N, sweep, len(got), len(want))
        assert np.array_equal(hashes, oracle.record_hashes(recs))
        if sweep == 0:
            assert len(got) == N
        if sweep == 1:
            assert len(got) == 0
    eng.close()


def test_hash_column_large_tables():
    """Enough slots that every persistent CTA walks many tiles through both bulk-copy stages: heavy repetition
    of a small pool, fully random rows, and near-duplicates that differ only in one late byte."""
    N = 1_200_000
    rng = np.random.default_rng(21)
    pool = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    pool[:, 0] = rng.integers(0, 32, 50)
    recs = pool[rng.integers(0, 50, N)]
    rnd = rng.random(N) < 0.2
    recs[rnd] = rng.integers(0, 256, (int(rnd.sum()), 32), dtype=np.uint8)
    near = rng.random(N) < 0.1
    recs[near, 31] ^= rng.integers(1, 256, int(near.sum())).astype(np.uint8)   # same head, different tail byte
    recs[:, 0] = np.minimum(recs[:, 0], 31)
    recs = np.ascontiguousarray(recs)
    eng = make_engine()
    idx, hashes = eng.status_diff(recs, want_hashes=True)
    assert np.array_equal(hashes, oracle.record_hashes(recs))
    assert len(idx) == N
    recs2 = recs.copy()
    rows = rng.choice(N, 5000, replace=False)
    recs2[rows, 5] ^= 0x40
    idx2, hashes2 = eng.status_diff(recs2, want_hashes=True)
    assert np.array_equal(hashes2, oracle.record_hashes(recs2))
    lens = recs[rows, 0]
    want = np.sort(rows[lens > 4]).astype(np.uint32)   # byte 5 of the slot is data byte 4: inside the hashed prefix iff len > 4
    assert np.array_equal(idx2, want)
    eng.close()


def test_seed_reset_and_resize():
    N = 5000
    eng = make_engine()
    base = rpk.synth.make_status_records(N, 0)
    eng.status_seed(base)  # CreatePod / LoadRunning: fill state, report nothing
    got, _ = eng.status_diff(base)
    assert len(got) == 0
    nxt = rpk.synth.make_status_records(N, 3, 0.2)
    got, _ = eng.status_diff(nxt)
    want = np.nonzero((base != nxt).any(axis=1))[0].astype(np.uint32)
    assert np.array_equal(got, want)
    with pytest.raises(rpk.RpkError) as ei:
        eng.status_diff(nxt[:100].copy())
    assert ei.value.code == rpk._ffi.RPK_ESTATE
    eng.status_reset(100)
    got, _ = eng.status_diff(nxt[:100].copy())
    assert np.array_equal(got, np.arange(100, dtype=np.uint32))
    with pytest.raises(rpk.RpkError):
        eng.status_diff(np.zeros((4, 20), np.uint8))  # stride not a multiple of 16
    eng.close()


def test_large_sweep_ordering_and_count():
    """4M slots: the single-pass look-back scan must emit every changed slot exactly once, ascending."""
    N = 1 << 22
    eng = make_engine()
    base = rpk.synth.make_status_records(N, 0)
    eng.status_seed(base)
    nxt = rpk.synth.make_status_records(N, 2, 0.37)
    got, _ = eng.status_diff(nxt)
    want = np.nonzero((base != nxt).any(axis=1))[0].astype(np.uint32)
    assert np.array_equal(got, want)
    eng.close()
