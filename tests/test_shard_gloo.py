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
