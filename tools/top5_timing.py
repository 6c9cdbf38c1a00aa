import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, rpk
eng = rpk.Engine(1)
G, P = 100_000, 1_000_000
offers = rpk.synth.make_offers(G); eng.upload_offers(offers)
pods = rpk.synth.make_pods(P)
d = {k: torch.from_numpy(v).cuda() for k, v in pods.items()}
best = torch.empty(P, dtype=torch.int32, device="cuda"); t5 = torch.empty(P * 5, dtype=torch.int32, device="cuda")
for want in (False, True):
    for _ in range(3): eng.select_device(d, best, t5 if want else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): eng.select_device(d, best, t5 if want else None)
    e1.record(); torch.cuda.synchronize()
    print("top5" if want else "best only", e0.elapsed_time(e1) / 5, "ms")
t = t5.view(P, 5).cpu().numpy()
print("rows with <5 hits:", float((t[:, 4] < 0).mean()), "rows with 0 hits:", float((t[:, 0] < 0).mean()))
