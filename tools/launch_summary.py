#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (run here, no GPU needed):
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/x.csv python bench.py ...
    python tools/launch_summary.py gpurun_out/x.csv "header line" [k_pod_,k_select_,k_status_,k_peer_] > profiles/x.summary.txt
(third argument: name fragments of the kernels that make up one bench step; the share column is taken over those)
Per-launch times under ncu are serialised and cold-cache; what must agree with bench.py is each kernel's SHARE."""
import csv
import sys
from collections import OrderedDict


def main():
    path, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v
        key = (r["Kernel Name"], r.get("Grid Size", ""), r.get("Block Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
    if header:
        print(header)
    print("(per-launch times under ncu are serialised and cold-cache; the SHARE of the step is what must agree with bench.py)")
    frags = (sys.argv[3] if len(sys.argv) > 3 else "k_pod_,k_select_,k_status_,k_peer_,k_gather_").split(",")
    in_step = lambda name: name.startswith(("rpk::", "void rpk::")) and any(f in name for f in frags)
    tot_rpk = sum(a[1] for k, a in agg.items() if in_step(k[0]))
    print(f"{'kernel':<96}{'grid':>16}{'block':>14}{'launches':>9}{'avg us':>10}{'share of the step':>19}")
    for (name, grid, block), (n, us) in agg.items():
        mine = in_step(name)
        share = f"{100 * us / tot_rpk:.1f} %" if mine and tot_rpk else ""
        print(f"{name[:94]:<96}{grid:>16}{block:>14}{n:>9}{us / n:>10.2f}{share:>19}")


if __name__ == "__main__":
    main()
