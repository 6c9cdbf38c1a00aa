#!/usr/bin/env python
"""Where does the host C-ABI path spend its time?  rpk_select alone, rpk_status_diff_codes alone and rpk_tick on pinned
host buffers (wall clock, median of --iters), next to the PCIe floor of the bytes each one moves.
    python tools/e2e_probe.py [--gpus N]"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--pods", type=int, default=1_000_000)
    ap.add_argument("--slots", type=int, default=1_000_000)
    ap.add_argument("--offers", type=int, default=100_000)
    ap.add_argument("--iters", type=int, default=7)
    args = ap.parse_args()
    import torch

    pkg = importlib.import_module("k8s-runpod-kubelet_b200")
    synth = pkg.synth

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    eng = pkg.Engine(args.gpus, device_ids=list(range(args.gpus)))
    eng.upload_offers(synth.make_offers(args.offers))
    keep = {k: pinned(v) for k, v in synth.make_pods(args.pods).items()}
    pods = {k: t.numpy() for k, t in keep.items()}
    best = torch.empty(args.pods, dtype=torch.int32, pin_memory=True).numpy()
    out = {"n_gpus": args.gpus, "pods": args.pods, "slots": args.slots}
    for stride in (16, 32):
        recs = [pinned(synth.make_status_records(args.slots, i, 0.01 * i, stride=stride)) for i in range(2)]

        def med(fn):
            for i in range(2):
                fn(i)
            ts = []
            for i in range(args.iters):
                t0 = time.perf_counter()
                fn(i)
                ts.append(time.perf_counter() - t0)
            ts.sort()
            return 1e3 * ts[len(ts) // 2]

        out[f"stride{stride}"] = {
            "select_ms": med(lambda i: eng.select(pods, out_best=best)),
            "status_ms": med(lambda i: eng.status_diff(recs[i & 1].numpy(), want_codes=True)),
            "tick_ms": med(lambda i: eng.tick(pods, recs[i & 1].numpy(), out_best=best)),
            "select_h2d_mb": sum(v.nbytes for v in pods.values()) / 1e6, "status_h2d_mb": args.slots * stride / 1e6,
        }
        st = eng.stats()
        out[f"stride{stride}"]["last_select_total_ms_events"] = st["last_select_total_ms"]
        out[f"stride{stride}"]["last_status_total_ms_events"] = st["last_status_total_ms"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
