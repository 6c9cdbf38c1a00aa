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
            got, hashes = eng.status_diff(recs, want_hashes=True)
            assert np.array_equal(got, tab.diff(recs))
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
