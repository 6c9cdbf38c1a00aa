"""The device entry points enqueue launches only (no allocation, no synchronisation once their scratch is sized),
so a caller can capture them in a CUDA graph and replay it -- bench.py does exactly that per step.  A replayed
graph must give the oracle's answer on whatever the input buffers hold at replay time, any number of times: the
select kernels' counters clean themselves (no per-call memset), which is what this pins."""
import numpy as np
import pytest

import oracle
import rpk

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("force", [None, "packed_pos", "generic"])
def test_select_and_status_replay_from_a_cuda_graph(engine, force):
    import torch

    from test_select_gpu import upload_forced

    dev = torch.device("cuda", 0)
    offers = rpk.synth.make_offers(40_000)       # three 16k-offer segments per row tile
    upload_forced(engine, offers, force)
    P, N = 40_000, 50_000                        # above the fused-kernel limit: k_pod_prep + grid kernel
    pods = [rpk.synth.make_pods(P, row0=r) for r in (0, 1_000_000, 2_000_000)]
    recs = [rpk.synth.make_status_records(N, s, f) for s, f in ((0, 0.0), (1, 0.05), (2, 0.5))]
    d_pods = {k: torch.from_numpy(v.copy()).to(dev) for k, v in pods[0].items()}
    d_recs = torch.from_numpy(recs[0].reshape(-1).copy()).to(dev)
    best = torch.full((P,), -9, dtype=torch.int32, device=dev)
    top5 = torch.full((P * 5,), -9, dtype=torch.int32, device=dev)
    hash_prev = torch.zeros(N, dtype=torch.int64, device=dev)
    changed = torch.empty(N, dtype=torch.int32, device=dev)
    n_changed = torch.zeros(1, dtype=torch.int32, device=dev)
    tab = oracle.StatusTable(N)

    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):                   # one eager call sizes the scratch buffers
        engine.select_device(d_pods, best, d_top5=top5)
        engine.status_diff_device(d_recs, 32, hash_prev, changed, n_changed)
    s.synchronize()
    want_idx = tab.diff(recs[0])
    assert int(n_changed.item()) == len(want_idx)

    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s, capture_error_mode="relaxed"):
        engine.select_device(d_pods, best, d_top5=top5)
        engine.status_diff_device(d_recs, 32, hash_prev, changed, n_changed)

    for it in (1, 2, 0, 1):                      # replays over changing inputs, one table repeated
        for k, v in pods[it].items():
            d_pods[k].copy_(torch.from_numpy(v))
        d_recs.copy_(torch.from_numpy(recs[it].reshape(-1)))
        best.fill_(-9)
        top5.fill_(-9)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        ob, ot = oracle.select(offers, pods[it], want_top5=True, n_threads=8)
        assert np.array_equal(best.cpu().numpy(), ob), f"replay on table {it}: best differs"
        assert np.array_equal(top5.cpu().numpy().reshape(P, 5), ot), f"replay on table {it}: top5 differs"
        want_idx = tab.diff(recs[it])
        n = int(n_changed.item())
        assert n == len(want_idx)
        assert np.array_equal(changed[:n].cpu().numpy().astype(np.uint32), want_idx)

    # eager calls after the replays still work on the same scratch (counters were left clean)
    best.fill_(-9)
    engine.select_device(d_pods, best)
    torch.cuda.synchronize()
    ob, _ = oracle.select(offers, pods[1], want_top5=False, n_threads=8)
    assert np.array_equal(best.cpu().numpy(), ob)
    upload_forced(engine, offers, None)


def test_counters_stay_clean_across_sizes_and_kernels(engine):
    """Back-to-back device selects with changing row counts (different tile counts, ragged last tiles, rows of
    an unknown cloud that belong to no tile) on one scratch: no call may see a previous call's tickets."""
    import torch

    dev = torch.device("cuda", 0)
    offers = rpk.synth.make_offers(33_000)
    engine.upload_offers(offers)
    for P in (70_000, 16_385, 250_001, 20_000, 131_072):
        pods = rpk.synth.make_pods(P, row0=P)
        pods["cloud"][::7] = 5               # neither SECURE nor COMMUNITY: nothing feasible, not in any row group
        d_pods = {k: torch.from_numpy(v).to(dev) for k, v in pods.items()}
        best = torch.full((P,), -9, dtype=torch.int32, device=dev)
        for _ in range(2):
            engine.select_device(d_pods, best)
        torch.cuda.synchronize()
        ob, _ = oracle.select(offers, pods, want_top5=False, n_threads=8)
        assert np.array_equal(best.cpu().numpy(), ob), P


@pytest.mark.parametrize("row0,P", [(0, 70_000), (3, 70_003), (1, 20_001), (2, 5), (5, 3), (7, 1)])
def test_gather_push_with_local_vectors(engine, row0, P):
    """rpk_select_device_gather with several output vectors on ONE GPU: the select kernels write the first vector,
    k_gather_push copies the slice [row0, row0 + P) into the others (head / 16-byte body / tail), nothing outside
    the slice is touched.  (With real peers the same kernel stores over NVLink: tests/test_multi_gpu.py.)"""
    import torch

    dev = torch.device("cuda", 0)
    offers = rpk.synth.make_offers(5_000)
    engine.upload_offers(offers)
    pods = rpk.synth.make_pods(P, row0=row0)
    d_pods = {k: torch.from_numpy(v).to(dev) for k, v in pods.items()}
    vecs = [torch.full((row0 + P + 9,), -9, dtype=torch.int32, device=dev) for _ in range(3)]
    ob, _ = oracle.select(offers, pods, want_top5=False, n_threads=4)
    for _ in range(2):
        engine.select_device_gather(d_pods, [v.data_ptr() for v in vecs], row0)
        torch.cuda.synchronize()
        for k, v in enumerate(vecs):
            h = v.cpu().numpy()
            assert np.array_equal(h[row0:row0 + P], ob), f"vector {k}"
            assert (h[:row0] == -9).all() and (h[row0 + P:] == -9).all(), f"vector {k}: written outside the slice"
