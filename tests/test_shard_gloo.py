"""world_size-2 gloo test (CPU) of the N>1 host logic the bench uses: contiguous pod-row shards, synthetic
rows generated per shard (row0 offset), all-gather of the per-shard assignment vector.  The per-shard compute
here is the oracle (test-only checker); on the GPU box the same plumbing carries the CUDA results."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import oracle, rpk
from bench import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P, G = 1000, 300
offers = rpk.synth.make_offers(G, correlated=True)
lo, hi = shard(P, world, rank)
pods = rpk.synth.make_pods(hi - lo, row0=lo)          # shard rows generated in place
best_local, _ = oracle.select(offers, pods, want_top5=False)
full = torch.empty(P, dtype=torch.int32)
dist.all_gather_into_tensor(full, torch.from_numpy(best_local))
want, _ = oracle.select(offers, rpk.synth.make_pods(P), want_top5=False)
assert np.array_equal(full.numpy(), want), "gathered shard results differ from the unsharded run"
# sharded status sweep: each rank diffs its slot shard (global ids = shard offset + local), the changed lists are
# exchanged and concatenated in rank order -- that must be the unsharded sweep's list (what bench.py checks on the GPUs)
NS = 5001
slo, shi = shard(NS, world, rank)
tab = oracle.StatusTable(shi - slo)
mine = None
for sweep, frac in ((0, 0.0), (1, 0.2)):
    mine = tab.diff(rpk.synth.make_status_records(shi - slo, sweep, frac, row0=slo)) + slo
lists = [None] * world
dist.all_gather_object(lists, mine)
gtab = oracle.StatusTable(NS)
want_idx = None
for sweep, frac in ((0, 0.0), (1, 0.2)):
    want_idx = gtab.diff(rpk.synth.make_status_records(NS, sweep, frac))
assert np.array_equal(np.concatenate(lists), want_idx), "concatenated shard lists differ from the unsharded sweep"
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)               # max-over-ranks timing reduction
assert t.item() == world
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
""" % ROOT


def test_row_sharding_and_gather_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


def test_shard_ranges_cover_exactly():
    from bench import shard

    for total in (0, 1, 7, 1000, 1_000_000, 1_000_003):
        for n in (1, 2, 3, 4, 8):
            spans = [shard(total, n, r) for r in range(n)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_bench_parity_checkers_catch_corruption():
    """bench.py verifies the timed result with these two functions: they must accept the oracle's own answer and
    reject a vector / list with a single wrong entry."""
    import numpy as np

    import oracle
    import rpk
    from bench import check_select_parity, check_status_parity

    offers = rpk.synth.make_offers(400, correlated=True)
    pods = rpk.synth.make_pods(3000)
    best, _ = oracle.select(offers, pods, want_top5=False)
    assert check_select_parity(best, offers, pods, oracle, 2, sample=200)["ok"]
    bad = best.copy()
    bad[1234] = (bad[1234] + 1) if bad[1234] >= 0 else 0
    res = check_select_parity(bad, offers, pods, oracle, 2, sample=200)
    assert not res["ok"] and res["rows_wrong_by_class"] == 1
    a, b = rpk.synth.make_status_records(2000, 0), rpk.synth.make_status_records(2000, 1, 0.1)
    t = oracle.StatusTable(2000)
    t.diff(a)
    idx = t.diff(b)
    codes = oracle.record_codes(b)[idx]
    assert check_status_parity(idx, codes, a, b, oracle)["ok"]
    assert not check_status_parity(idx[:-1], codes[:-1], a, b, oracle)["ok"]
    wrong = codes.copy()
    wrong[0] ^= 1
    assert not check_status_parity(idx, wrong, a, b, oracle)["ok"]
