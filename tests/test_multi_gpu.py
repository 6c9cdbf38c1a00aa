"""Multi-GPU tests (need >= 2 B200s: `gpurun --gpus 2 -- pytest tests -m gpu`): one ctx over several GPUs, pod
rows / status slots sharded, every GPU ends with the whole assignment vector (NVLink peer stores)."""
import importlib

import numpy as np
import pytest

import oracle
import rpk

pytestmark = pytest.mark.gpu


def n_devices():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("n,P", [(2, 50_001), (2, 50_003), (2, 300_007), (2, 5), (2, 3), (4, 50_001), (4, 50_003), (8, 50_001), (8, 50_003)])
def test_sharded_select_and_gather(n, P):
    """P not divisible by n: ragged shards; 50_003 / 300_007 put shard and sub-batch boundaries off the 16-byte
    grid (head / tail of k_gather_push); 5 and 3 rows: slices shorter than one 16-byte unit."""
    import torch

    if n_devices() < n:
        pytest.skip(f"needs {n} GPUs")
    peer = importlib.import_module("k8s-runpod-kubelet_b200.peer")
    offers = rpk.synth.make_offers(20_000, correlated=True)
    pods = rpk.synth.make_pods(P)
    with rpk.Engine(n) as eng:
        eng.upload_offers(offers)
        best, top5 = eng.select(pods, want_top5=True)
        ob, ot = oracle.select(offers, pods, n_threads=8)
        assert np.array_equal(best, ob) and np.array_equal(top5, ot)
        for s in range(n):  # the all-gather: every GPU holds all P assignments
            with torch.cuda.device(s):
                t = peer.as_int32_tensor(eng.best_device_ptr(s), len(ob), torch.device("cuda", s))
                assert np.array_equal(t.cpu().numpy(), ob), f"GPU {s} does not hold the full vector"
        # status sweep sharded over the same GPUs
        N = 30_001
        tab = oracle.StatusTable(N)
        for sweep, frac in enumerate([0.0, 0.05, 1.0]):
            recs = rpk.synth.make_status_records(N, sweep, frac)
            got, codes, hashes = eng.status_diff(recs, want_hashes=True, want_codes=True)
            want = tab.diff(recs)
            assert np.array_equal(got, want)
            assert np.array_equal(codes, oracle.record_codes(recs)[want])
            assert np.array_equal(hashes, oracle.record_hashes(recs))


def test_two_contexts_are_independent():
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs")
    offers = rpk.synth.make_offers(3000, correlated=True)
    pods = rpk.synth.make_pods(4000)
    with rpk.Engine(1, device_ids=[1]) as e1, rpk.Engine(1, device_ids=[0]) as e0:
        e0.upload_offers(offers)
        e1.upload_offers(offers)
        b0, _ = e0.select(pods)
        b1, _ = e1.select(pods)
        assert np.array_equal(b0, b1)


@pytest.mark.parametrize("self_counting", [False, True])
def test_peer_fence(self_counting):
    """rpk_peer_fence across two GPUs of one ctx, explicit epochs and the self-counting (graph-replayable) mode:
    each GPU's fence completes only once the other one has signalled."""
    import torch

    if n_devices() < 2:
        pytest.skip("needs 2 GPUs")
    with rpk.Engine(2) as eng:
        flags = [eng.ipc_alloc(64 * 4, shard=s)[0] for s in range(2)]   # same process: peers reach them directly
        streams = [torch.cuda.Stream(device=torch.device("cuda", s)) for s in range(2)]
        for rnd in range(1, 6):
            for s in (0, 1) if rnd & 1 else (1, 0):
                eng.peer_fence(flags, s, 0 if self_counting else rnd, shard=s, stream=streams[s].cuda_stream)
            for st in streams:
                st.synchronize()
        for s in range(2):
            with torch.cuda.device(s):
                peer = importlib.import_module("k8s-runpod-kubelet_b200.peer")
                t = peer.as_int32_tensor(flags[s], 64, torch.device("cuda", s)).cpu().numpy()
                assert t[0] == 5 and t[1] == 5, t[:4]
                assert t[32] == (5 if self_counting else 0)


@pytest.mark.parametrize("n,inline", [(2, False), (2, True), (4, True), (8, False), (8, True)])
def test_device_gather_bound_signal_and_wait(n, inline):
    """The one-process-per-GPU shape (bench.py under torchrun) inside one process: every shard selects its rows with
    rpk_select_device_gather on its own stream -- blocks are pushed to all peers from inside the kernel, the last
    pusher signals -- sweeps its slots with rpk_status_diff_device_gather into every shard's exchange buffer, and
    rpk_peer_wait(3) closes the step -- or, with rpk_peer_inline_wait, nothing does: the signalling warps wait
    themselves.  Every GPU must end with the oracle's whole assignment vector and the whole
    changed list + codes; two steps on two buffer sets, ragged shards, slices off the 16-byte grid."""
    import torch

    if n_devices() < n:
        pytest.skip(f"needs {n} GPUs")
    peer = importlib.import_module("k8s-runpod-kubelet_b200.peer")
    offers = rpk.synth.make_offers(20_000, correlated=True)
    P, NS = 150_003, 70_001
    pods = rpk.synth.make_pods(P)
    ob, _ = oracle.select(offers, pods, want_top5=False, n_threads=8)
    cap = -(-NS // n) + 1
    with rpk.Engine(n) as eng:
        eng.upload_offers(offers)
        xb = eng.xchg_bytes(n, cap)
        vec = [[eng.ipc_alloc(P * 4, shard=s)[0] for s in range(n)] for _ in range(2)]   # [set][shard]
        xch = [[eng.ipc_alloc(xb, shard=s)[0] for s in range(n)] for _ in range(2)]
        flags = [eng.ipc_alloc(64 * 4, shard=s)[0] for s in range(n)]
        streams, sides, d_pods, d_recs, d_hash, d_n = [], [], [], [], [], []
        tabs = [rpk.synth.make_status_records(NS, 0), rpk.synth.make_status_records(NS, 1, 0.05)]
        for s in range(n):
            dev = torch.device("cuda", s)
            lo, hi = P * s // n, P * (s + 1) // n
            slo, shi = NS * s // n, NS * (s + 1) // n
            with torch.cuda.device(s):
                streams.append(torch.cuda.Stream(device=dev))
                sides.append(torch.cuda.Stream(device=dev))
                d_pods.append({k: torch.from_numpy(np.ascontiguousarray(v[lo:hi])).to(dev) for k, v in pods.items()})
                d_recs.append([torch.from_numpy(np.ascontiguousarray(t[slo:shi]).reshape(-1)).to(dev) for t in tabs])
                d_hash.append(torch.zeros(shi - slo, dtype=torch.int64, device=dev))
                d_n.append(torch.zeros(1, dtype=torch.int32, device=dev))
            eng.peer_bind(flags, s, shard=s)
        tab = oracle.StatusTable(NS)
        for step in range(5):
            par = step & 1
            if step == 1:   # only after one call has sized the scratch: an allocating call synchronises its device, and
                for s in range(n):  # with ONE host thread driving every shard a kernel that waits for a peer would never see it
                    eng.peer_inline_wait(inline, shard=s)
            waits = not (inline and step >= 1)
            for s in (range(n) if step & 1 else reversed(range(n))):   # issue order must not matter
                lo = P * s // n
                slo = NS * s // n
                with torch.cuda.device(s):
                    eng.status_diff_device_gather(d_recs[s][par], 32, d_hash[s], slo, xch[par], cap, s, d_n[s], shard=s, stream=sides[s].cuda_stream)
                    eng.select_device_gather(d_pods[s], vec[par], lo, shard=s, stream=streams[s].cuda_stream)
                    streams[s].wait_stream(sides[s])
                    if waits:
                        eng.peer_wait(3, shard=s, stream=streams[s].cuda_stream)
            for st in streams:
                st.synchronize()
            want = tab.diff(tabs[par])
            wcodes = oracle.record_codes(tabs[par])[want]
            for s in range(n):
                with torch.cuda.device(s):
                    dev = torch.device("cuda", s)
                    t = peer.as_int32_tensor(vec[par][s], P, dev).cpu().numpy()
                    assert np.array_equal(t, ob), f"step {step}: GPU {s} does not hold the oracle's vector"
                    xv = peer.as_int32_tensor(xch[par][s], xb // 4, dev).cpu().numpy()
                    counts = xv[:n].astype(np.int64)
                    idx = np.concatenate([xv[8 + r * cap: 8 + r * cap + counts[r]] for r in range(n)]).astype(np.uint32)
                    c16 = xv[8 + n * cap:].view(np.uint16)
                    codes = np.concatenate([c16[r * cap: r * cap + counts[r]] for r in range(n)])
                    assert np.array_equal(idx, want), f"step {step}: GPU {s} changed list"
                    assert np.array_equal(codes, wcodes), f"step {step}: GPU {s} codes"


def test_tick_over_two_gpus():
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs")
    offers = rpk.synth.make_offers(20_000, correlated=True)
    pods = rpk.synth.make_pods(300_007)
    ob, _ = oracle.select(offers, pods, want_top5=False, n_threads=8)
    N = 90_001
    with rpk.Engine(2) as eng:
        eng.upload_offers(offers)
        tab = oracle.StatusTable(N, 16)
        for sweep, frac in enumerate([0.0, 0.02, 0.4]):
            recs = rpk.synth.make_status_records(N, sweep, frac, stride=16)
            best, _, idx, codes = eng.tick(pods, recs)
            want = tab.diff(recs)
            assert np.array_equal(best, ob) and np.array_equal(idx, want) and np.array_equal(codes, oracle.record_codes(recs)[want])
