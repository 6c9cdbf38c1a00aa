#!/usr/bin/env python
"""1-GPU probe of the gather path's fixed costs: the same select into (1) a torch tensor, (2) a CUDA-IPC-exportable
vector (rpk_ipc_alloc), (3) two torch vectors (k_gather_push copies local -> local), (4) two IPC vectors.  Separates
what the kernels cost from what NVLink costs when the N-GPU bench is slower than P/N rows on one GPU."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    pkg = importlib.import_module("k8s-runpod-kubelet_b200")
    peer = importlib.import_module("k8s-runpod-kubelet_b200.peer")
    synth = pkg.synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    P, G = int(os.environ.get("PROBE_P", 500_000)), 100_000
    eng = pkg.Engine(1, device_ids=[0])
    eng.upload_offers(synth.make_offers(G))
    d_pods = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_pods(P).items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t_a, t_b = torch.empty(P, dtype=torch.int32, device=dev), torch.empty(P, dtype=torch.int32, device=dev)
    i_a, i_b = eng.ipc_alloc(P * 4)[0], eng.ipc_alloc(P * 4)[0]
    cases = {
        "torch vector": lambda: eng.select_device(d_pods, t_a),
        "ipc vector": lambda: eng.select_device_gather(d_pods, [i_a], 0),
        "two torch vectors (local push)": lambda: eng.select_device_gather(d_pods, [t_a.data_ptr(), t_b.data_ptr()], 0),
        "two ipc vectors (local push)": lambda: eng.select_device_gather(d_pods, [i_a, i_b], 0),
    }
    out = {}
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for i in range(10):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        out[name] = {"us_median": 1e3 * ms[5], "us_min": 1e3 * ms[0]}
        print(name, out[name], flush=True)
    same = bool((peer.as_int32_tensor(i_b, P, dev) == t_a).all()) and bool((t_b == t_a).all())
    out["all vectors equal"] = same
    print("all vectors equal:", same)
    json.dump({"P": P, "G": G, "cases": out}, open(os.path.join(ROOT, "gpurun_out", "gather_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
