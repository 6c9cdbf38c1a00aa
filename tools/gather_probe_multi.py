#!/usr/bin/env python
"""N-GPU probe (one process per GPU, torchrun) that takes the N-GPU step of bench.py apart, per rank, to find where
an N-GPU step spends the time one GPU doing the same rows does not (43 us at N = 2):

    A  select into the rank's own vector only                      (n_out = 1: the 1-GPU cost of the shard)
    B  select with the fused per-block push + signal, no wait      (adds the NVLink push from inside the kernel)
    C  B + rpk_peer_wait                                           (adds the wait: NVLink latency and rank skew)
    D  C with the sharded status sweep (changed list exchanged) on a side stream
    E  B with rpk_peer_inline_wait: the last pusher waits itself, no wait launch
    F  D with the waits inside both kernels (= the bench step)

each as eager launches and as one CUDA graph replay, L2 flushed between iterations, CUDA events on the launching
stream.  Every rank prints its own median (no max-over-ranks), so an asymmetric rank shows up.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/gather_probe_multi.py [--pods 1000000] [--iters 20]

Every ingredient is the same call bench.py makes.
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pods", type=int, default=1_000_000, help="P total, sharded by row")
    ap.add_argument("--offers", type=int, default=100_000)
    ap.add_argument("--slots", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from bench import shard

    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    pkg = importlib.import_module("k8s-runpod-kubelet_b200")
    peer = importlib.import_module("k8s-runpod-kubelet_b200.peer")
    synth = pkg.synth
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    P, G, NS = args.pods, args.offers, args.slots
    lo, hi = shard(P, world, rank)
    slo, shi = shard(NS, world, rank)
    eng = pkg.Engine(1, device_ids=[local_rank])
    eng.upload_offers(synth.make_offers(G))
    d_pods = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_pods(hi - lo, row0=lo).items()}
    own, ptrs = peer.exchange_peer_buffer(eng, P * 4, rank, world)
    cap = -(-NS // world) + 1
    _, xptrs = peer.exchange_peer_buffer(eng, eng.xchg_bytes(world, cap), rank, world)
    flag_ptrs = peer.exchange_peer_buffer(eng, 64 * 4, rank, world)[1]
    own_only = [ptrs[rank]]
    recs = [torch.from_numpy(synth.make_status_records(shi - slo, i, 0.01 * i, row0=slo).reshape(-1)).to(dev) for i in range(2)]
    hash_prev = torch.zeros(shi - slo, dtype=torch.int64, device=dev)
    changed = torch.empty(max(shi - slo, 1), dtype=torch.int32, device=dev)
    n_changed = torch.zeros(1, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))  # as bench.py: the selection's CTAs are placed first
    side = torch.cuda.Stream(device=dev, priority=0)
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

    def case_a(i):
        eng.peer_inline_wait(False)
        eng.peer_bind(None, 0)
        eng.select_device_gather(d_pods, own_only, lo)

    def case_b(i):
        eng.peer_bind(flag_ptrs, rank)
        eng.select_device_gather(d_pods, ptrs, lo)

    def case_c(i):
        eng.peer_bind(flag_ptrs, rank)
        eng.select_device_gather(d_pods, ptrs, lo)
        eng.peer_wait(1)

    def case_d(i):
        eng.peer_bind(flag_ptrs, rank)
        ev_fork.record()
        side.wait_event(ev_fork)
        eng.status_diff_device_gather(recs[i & 1], 32, hash_prev, slo, xptrs, cap, rank, n_changed, stream=side.cuda_stream)
        ev_join.record(side)
        eng.select_device_gather(d_pods, ptrs, lo)
        torch.cuda.current_stream().wait_event(ev_join)
        eng.peer_wait(3)

    def case_e(i):
        eng.peer_bind(flag_ptrs, rank)
        eng.peer_inline_wait(True)
        eng.select_device_gather(d_pods, ptrs, lo)

    def case_f(i):
        eng.peer_bind(flag_ptrs, rank)
        eng.peer_inline_wait(True)
        ev_fork.record()
        side.wait_event(ev_fork)
        eng.status_diff_device_gather(recs[i & 1], 32, hash_prev, slo, xptrs, cap, rank, n_changed, stream=side.cuda_stream)
        ev_join.record(side)
        eng.select_device_gather(d_pods, ptrs, lo)
        torch.cuda.current_stream().wait_event(ev_join)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(run):
        ms = []
        for i in range(args.iters):
            flush.fill_(i & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(i)
            e1.record()
            ms.append((e0, e1))
        barrier()
        v = sorted(a.elapsed_time(b) for a, b in ms)
        return {"us_median": 1e3 * v[len(v) // 2], "us_min": 1e3 * v[0], "us_max": 1e3 * v[-1]}

    out = {"rank": rank, "world": world, "rows": hi - lo, "slots": shi - slo, "cases": {}}
    # B signals without anybody waiting: harmless (epochs only grow), but keep C / D last so that every wait sees all signals
    for name, fn in (("A select, own vector", case_a), ("B select + fused push + signal", case_b), ("C B + peer wait", case_c),
                     ("D C + sharded status sweep alongside", case_d),
                     ("E B with the wait inside the kernel", case_e), ("F D with the waits inside the kernels", case_f)):
        for i in range(3):
            fn(i)
        barrier()
        res = {"eager": timed(fn)}
        graphs = None
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            cap_stream = torch.cuda.Stream(device=dev, priority=-1)
            cap_stream.wait_stream(torch.cuda.current_stream())
            graphs = []
            for parity in (0, 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap_stream, capture_error_mode="relaxed"):
                    fn(parity)
                graphs.append(g)
            torch.cuda.current_stream().wait_stream(cap_stream)
        except Exception as e:  # all ranks must replay the same number of fences: agree on the mode
            ok.zero_()
            res["graph_error"] = f"{type(e).__name__}: {e}"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            for i in range(3):
                graphs[i & 1].replay()
            barrier()
            res["graph"] = timed(lambda i: graphs[i & 1].replay())
        out["cases"][name] = res
        barrier()
    lines = [None] * world
    dist.all_gather_object(lines, out)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(lines, open(os.path.join(ROOT, "gpurun_out", f"gather_probe_n{world}.json"), "w"), indent=1)
        for r in lines:
            for name, res in r["cases"].items():
                print(f"rank {r['rank']} rows {r['rows']}: {name}: " + ", ".join(
                    f"{mode} {v['us_median']:.1f} us (min {v['us_min']:.1f})" for mode, v in res.items() if isinstance(v, dict)), flush=True)
    barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
