#!/usr/bin/env python
"""BASELINE config 5 -- streaming reconcile: p50/p99 assignment latency under pod churn.

Poisson pod arrivals (lambda = 10^4/s for `--seconds` s, seed 0x52504B35), G = 10^4 resident offers; a
micro-batch is flushed when it holds 256 pods or its oldest pod has waited 1 ms; each flush is one
`rpk_select` through the host C-ABI (top-5 list included, as the deploy step needs gpuTypeIds).  Concurrently
10^4 status mutations/s hit N = 10^5 tracked slots; every 10 ms one `rpk_status_diff` sweep runs on the same
host thread.  Latency of a pod = (its batch's results are back on the host) - (its arrival time).
With --gpus N > 1 the micro-batches are round-robined over N single-GPU contexts (no collective on the latency
path: SURVEY.md 8e); the host loop stays single-threaded like the reference's one pod-sync worker.

Prints one JSON line; not part of the driver's bench contract (that is bench.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--rate", type=float, default=1e4)
    ap.add_argument("--offers", type=int, default=10_000)
    ap.add_argument("--slots", type=int, default=100_000)
    ap.add_argument("--max-batch", type=int, default=256)
    ap.add_argument("--max-wait-ms", type=float, default=1.0)
    ap.add_argument("--policy", default="window", choices=["window", "eager"],
                    help="window: SURVEY 8d flush rule (256 pods or 1 ms); eager: flush whenever the engine is free and a pod waits")
    args = ap.parse_args()

    rng = np.random.default_rng(0x52504B35)
    n_arr = int(args.rate * args.seconds)
    arrivals = np.cumsum(rng.exponential(1.0 / args.rate, n_arr))
    pods_all = rpk.synth.make_pods(n_arr, seed=0x52504B35)
    offers = rpk.synth.make_offers(args.offers)
    engines = [rpk.Engine(1, device_ids=[g]) for g in range(args.gpus)]
    for e in engines:
        e.upload_offers(offers)
    recs = rpk.synth.make_status_records(args.slots, 0)
    lut = rpk.synth.make_status_records(4096, 1, 1.0)  # pool of records to draw mutations from
    engines[0].status_seed(recs)
    best_buf = np.empty(args.max_batch, np.int32)
    top5_buf = np.empty((args.max_batch, 5), np.int32)
    # warm-up: every code path once
    for e in engines:
        e.select({k: np.ascontiguousarray(v[:32]) for k, v in pods_all.items()}, want_top5=True)
    engines[0].status_diff(recs)

    lat = np.empty(n_arr, np.float64)
    batch_sizes, service, sweep_ms, sweeps, changed_total = [], [], [], 0, 0
    mut_per_sweep = int(args.rate * 0.01)
    nxt, done, rr = 0, 0, 0
    t0 = time.perf_counter()
    next_sweep = 0.01
    max_wait = args.max_wait_ms * 1e-3
    while done < n_arr:
        now = time.perf_counter() - t0
        while nxt < n_arr and arrivals[nxt] <= now:
            nxt += 1
        pending = nxt - done
        if pending and (args.policy == "eager" or pending >= args.max_batch or now - arrivals[done] >= max_wait):
            b = min(pending, args.max_batch)
            sl = slice(done, done + b)
            batch = {k: v[sl] for k, v in pods_all.items()}  # contiguous views
            t_call = time.perf_counter()
            engines[rr % args.gpus].select(batch, want_top5=True, out_best=best_buf[:b], out_top5=top5_buf[:b])
            t_done = time.perf_counter() - t0
            service.append(time.perf_counter() - t_call)
            lat[sl] = t_done - arrivals[sl]
            batch_sizes.append(b)
            done += b
            rr += 1
            continue
        if now >= next_sweep:
            rows = rng.integers(0, args.slots, mut_per_sweep)
            recs[rows] = lut[rng.integers(0, lut.shape[0], mut_per_sweep)]
            t_call = time.perf_counter()
            idx, _ = engines[0].status_diff(recs)
            sweep_ms.append((time.perf_counter() - t_call) * 1e3)
            changed_total += len(idx)
            sweeps += 1
            next_sweep += 0.01
    wall = time.perf_counter() - t0
    out = {
        "config": "C5 streaming reconcile", "n_gpus": args.gpus, "arrival_rate_per_s": args.rate, "pods": n_arr,
        "offers": args.offers, "status_slots": args.slots, "flush": f"{args.max_batch} pods or {args.max_wait_ms} ms",
        "latency_ms": {"p50": float(np.percentile(lat, 50) * 1e3), "p90": float(np.percentile(lat, 90) * 1e3),
                       "p99": float(np.percentile(lat, 99) * 1e3), "max": float(lat.max() * 1e3), "mean": float(lat.mean() * 1e3)},
        "policy": args.policy,
        "select_call_ms": {"mean": float(np.mean(service) * 1e3), "p50": float(np.percentile(service, 50) * 1e3), "p99": float(np.percentile(service, 99) * 1e3)},
        "status_sweep_call_ms": {"mean": float(np.mean(sweep_ms)), "p99": float(np.percentile(sweep_ms, 99))} if sweep_ms else None,
        "batches": len(batch_sizes), "mean_batch": float(np.mean(batch_sizes)), "status_sweeps": sweeps,
        "status_changed_total": changed_total, "wall_s": wall,
        "api": "rpk_select (top-5) + rpk_status_diff through the host C-ABI, pageable numpy buffers, single host thread",
    }
    print(json.dumps(out), flush=True)
    for e in engines:
        e.close()


if __name__ == "__main__":
    main()
