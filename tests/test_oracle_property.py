"""Property tests (hypothesis, CPU only): the C oracle and the independently written numpy restatement must
agree on arbitrary small tables -- duplicated prices (tie order), zero / negative / NaN / infinite prices,
extreme int32 requests, unknown cloud bytes, missing extension columns."""
import numpy as np
from hypothesis import given, settings, strategies as st

import np_restatement as npr
import oracle

price = st.one_of(st.sampled_from([0.0, -0.0, 0.25, 0.5, 0.49999999999999994, 0.5000000000000001, 1.0, float("nan"), float("inf"), -1.0, 5e-324]),
                  st.floats(min_value=0.0, max_value=4.0, allow_nan=False))
i32 = st.one_of(st.sampled_from([0, 1, 16, 24, 48, 80, -1, -2**31, 2**31 - 2]), st.integers(-100, 200))
req = st.one_of(st.sampled_from([0, 16, 24, 2**31 - 1, -2**31]), st.integers(-100, 250))


@st.composite
def tables(draw):
    G = draw(st.integers(0, 40))
    P = draw(st.integers(1, 12))
    ext = draw(st.booleans())
    offers = {
        "mem_gb": np.array(draw(st.lists(i32, min_size=G, max_size=G)), np.int32),
        "secure_price": np.array(draw(st.lists(price, min_size=G, max_size=G)), np.float64),
        "community_price": np.array(draw(st.lists(price, min_size=G, max_size=G)), np.float64),
        "flags": np.array(draw(st.lists(st.integers(0, 3), min_size=G, max_size=G)), np.uint8),
        "vcpu": np.array(draw(st.lists(i32, min_size=G, max_size=G)), np.int32) if ext else None,
        "ram_gb": np.array(draw(st.lists(i32, min_size=G, max_size=G)), np.int32) if ext else None,
    }
    pods = {
        "req_mem_gb": np.array(draw(st.lists(req, min_size=P, max_size=P)), np.int32),
        "req_vcpu": np.array(draw(st.lists(req, min_size=P, max_size=P)), np.int32),
        "req_ram_gb": np.array(draw(st.lists(req, min_size=P, max_size=P)), np.int32),
        "max_price": np.array(draw(st.lists(price, min_size=P, max_size=P)), np.float64),
        "cloud": np.array(draw(st.lists(st.sampled_from([0, 0, 0, 1, 1, 2, 255]), min_size=P, max_size=P)), np.uint8),
    }
    return offers, pods


@settings(max_examples=300, deadline=None)
@given(tables())
def test_oracle_equals_numpy_restatement(t):
    offers, pods = t
    b0, t0 = oracle.select(offers, pods)
    b1, t1 = npr.select(offers, pods)
    assert np.array_equal(b0, b1) and np.array_equal(t0, t1)
    # structural properties of any GetGPUTypes result (runpod_client.go:465-509)
    for p in range(len(b0)):
        ids = [i for i in t0[p] if i >= 0]
        assert list(t0[p][: len(ids)]) == ids and all(x == -1 for x in t0[p][len(ids):])  # -1 padding only at the end
        assert len(set(ids)) == len(ids) and (b0[p] == (ids[0] if ids else -1))
        c = int(pods["cloud"][p])
        if c > 1:
            assert not ids
            continue
        pr = offers["secure_price"] if c == 0 else offers["community_price"]
        for i in ids:  # every listed offer satisfies the predicate of :478
            assert (offers["flags"][i] >> c) & 1 and pr[i] > 0 and pr[i] < pods["max_price"][p]
            assert offers["mem_gb"][i] >= pods["req_mem_gb"][p]
        assert all(pr[a] <= pr[b] for a, b in zip(ids, ids[1:]))  # ascending price
        assert all(a < b for a, b in zip(ids, ids[1:]) if pr[a] == pr[b])  # ties: lower index first


@settings(max_examples=200, deadline=None)
@given(st.binary(min_size=0, max_size=300), st.integers(0, 2**64 - 1))
def test_xxh64_matches_python_xxhash(data, seed):
    try:
        import xxhash
    except ImportError:
        return
    assert oracle.xxh64(data, seed) == xxhash.xxh64(data, seed=seed).intdigest()
