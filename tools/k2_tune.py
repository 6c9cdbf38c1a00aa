#!/usr/bin/env python
"""K2 timing (1 GPU): the status sweep alone, device-resident records, L2 flushed between launches.
    python tools/k2_tune.py [--out gpurun_out/k2_tune.json]
Algorithmic bytes per slot = stride + 8 read, + (8 + 4 + 2)*f written (hash, index, code per changed slot)."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--slots", default="125000,1000000,16777216")
    ap.add_argument("--strides", default="16,32")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "k2_tune.json"))
    args = ap.parse_args()
    import torch

    pkg = importlib.import_module("k8s-runpod-kubelet_b200")
    synth = pkg.synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    eng = pkg.Engine(1, device_ids=[0])
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for stride in [int(x) for x in args.strides.split(",")]:
        for N in [int(x) for x in args.slots.split(",")]:
            base = synth.make_status_records(min(N, 1 << 20), 0, stride=stride)
            reps = (N + base.shape[0] - 1) // base.shape[0]
            a = torch.from_numpy(base).to(dev).repeat(reps, 1)[:N].contiguous()
            b = a.clone()
            rows = torch.arange(0, N, 100, device=dev)  # every 100th slot takes its neighbour's record
            b[rows] = a[(rows + 1) % N]
            tabs = [a.reshape(-1), b.reshape(-1)]
            hp = torch.zeros(N, dtype=torch.int64, device=dev)
            chg = torch.empty(N, dtype=torch.int32, device=dev)
            code = torch.empty(N, dtype=torch.int16, device=dev)
            nch = torch.zeros(1, dtype=torch.int32, device=dev)
            for with_codes in (False, True):
                for i in range(3):
                    eng.status_diff_device(tabs[i & 1], stride, hp, chg, nch, d_changed_code=code if with_codes else None)
                torch.cuda.synchronize()
                ms, n_changed = [], 0
                for i in range(args.iters):
                    flush.fill_(i)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    eng.status_diff_device(tabs[(i + 1) & 1], stride, hp, chg, nch, d_changed_code=code if with_codes else None)
                    e1.record()
                    torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1))
                    n_changed = int(nch.item())
                ms.sort()
                med = ms[len(ms) // 2]
                algo = N * (stride + 8) + n_changed * (12 + (2 if with_codes else 0))
                row = {"stride": stride, "slots": N, "codes": with_codes, "changed": n_changed, "us_median": med * 1e3, "us_min": ms[0] * 1e3,
                       "GBps": algo / (med * 1e-3) / 1e9, "frac_of_measured_hbm": algo / (med * 1e-3) / 1e9 / peak,
                       "pods_per_s": N / (med * 1e-3)}
                print(json.dumps(row), flush=True)
                out.append(row)
            del a, b, tabs, hp, chg, code
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
