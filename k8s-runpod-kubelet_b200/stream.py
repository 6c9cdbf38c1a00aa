"""BASELINE config 5 -- streaming reconcile: assignment latency under pod churn, through the host C-ABI.

Poisson pod arrivals (lambda pods/s, seed 0x52504B35) against G resident offers; a micro-batch is flushed when it
holds ``max_batch`` pods or its oldest pod has waited ``max_wait_ms`` (``window`` policy) or whenever a
pod waits and the engine is free (``eager``); each flush is ONE ``rpk_select`` with the top-5 list (the deploy step
needs gpuTypeIds, runpod_client.go:1339).  Concurrently ``rate * 0.01`` status mutations hit the tracked slots every
10 ms and one ``rpk_status_diff`` sweep runs on the same host thread -- the shape of the reference's loops
(kubelet.go:747-814 and 816-974), batched.  Latency of a pod = (its batch's results are back on the host) - (its
arrival).  With several engines the micro-batches are round-robined over them: the latency path has no collective
(SURVEY.md 8e); the host loop stays single-threaded like the reference's one pod-sync worker (main.go:263).

Returns the summary and every assignment (best, top5) so that the caller can check them.
"""
from __future__ import annotations

import time

import numpy as np

from . import synth

STREAM_SEED = 0x52504B35


def run_stream(engines, offers, seconds: float = 2.0, rate: float = 1e4, slots: int = 100_000, policy: str = "window",
               max_batch: int = 256, max_wait_ms: float = 1.0, stride: int = 16):
    rng = np.random.default_rng(STREAM_SEED)
    n_arr = int(rate * seconds)
    arrivals = np.cumsum(rng.exponential(1.0 / rate, n_arr))
    pods_all = synth.make_pods(n_arr, seed=STREAM_SEED)
    for e in engines:
        e.upload_offers(offers)
    recs = synth.make_status_records(slots, 0, stride=stride)
    lut = synth.make_status_records(4096, 1, 1.0, stride=stride)  # pool of records to draw mutations from
    engines[0].status_reset(slots)
    engines[0].status_seed(recs)
    best_all = np.full(n_arr, -9, np.int32)
    top5_all = np.full((n_arr, 5), -9, np.int32)
    for e in engines:  # warm-up: every code path once
        e.select({k: np.ascontiguousarray(v[:32]) for k, v in pods_all.items()}, want_top5=True)
    engines[0].status_diff(recs)

    lat = np.empty(n_arr, np.float64)
    batch_sizes, service, sweep_ms, sweeps, changed_total = [], [], [], 0, 0
    mut_per_sweep = max(1, int(rate * 0.01))
    nxt, done, rr = 0, 0, 0
    t0 = time.perf_counter()
    next_sweep = 0.01
    max_wait = max_wait_ms * 1e-3
    while done < n_arr:
        now = time.perf_counter() - t0
        while nxt < n_arr and arrivals[nxt] <= now:
            nxt += 1
        pending = nxt - done
        if pending and (policy == "eager" or pending >= max_batch or now - arrivals[done] >= max_wait):
            b = min(pending, max_batch)
            sl = slice(done, done + b)
            batch = {k: v[sl] for k, v in pods_all.items()}  # contiguous views
            t_call = time.perf_counter()
            engines[rr % len(engines)].select(batch, want_top5=True, out_best=best_all[sl], out_top5=top5_all[sl])
            t_done = time.perf_counter() - t0
            service.append(time.perf_counter() - t_call)
            lat[sl] = t_done - arrivals[sl]
            batch_sizes.append(b)
            done += b
            rr += 1
            continue
        if now >= next_sweep:
            rows = rng.integers(0, slots, mut_per_sweep)
            recs[rows] = lut[rng.integers(0, lut.shape[0], mut_per_sweep)]
            t_call = time.perf_counter()
            idx, _ = engines[0].status_diff(recs)
            sweep_ms.append((time.perf_counter() - t_call) * 1e3)
            changed_total += len(idx)
            sweeps += 1
            next_sweep += 0.01
    wall = time.perf_counter() - t0
    out = {
        "n_gpus": len(engines), "arrival_rate_per_s": rate, "pods": n_arr, "offers": int(offers["mem_gb"].shape[0]),
        "status_slots": slots, "policy": policy, "flush": f"{max_batch} pods or {max_wait_ms} ms" if policy == "window" else "whenever a pod waits",
        "latency_ms": {"p50": float(np.percentile(lat, 50) * 1e3), "p90": float(np.percentile(lat, 90) * 1e3),
                       "p99": float(np.percentile(lat, 99) * 1e3), "max": float(lat.max() * 1e3), "mean": float(lat.mean() * 1e3)},
        "select_call_ms": {"mean": float(np.mean(service) * 1e3), "p50": float(np.percentile(service, 50) * 1e3), "p99": float(np.percentile(service, 99) * 1e3)},
        "status_sweep_call_ms": {"mean": float(np.mean(sweep_ms)), "p99": float(np.percentile(sweep_ms, 99))} if sweep_ms else None,
        "batches": len(batch_sizes), "mean_batch": float(np.mean(batch_sizes)), "status_sweeps": sweeps,
        "status_changed_total": changed_total, "wall_s": wall,
    }
    return out, pods_all, best_all, top5_all
