#!/usr/bin/env python
"""Per-kernel static SASS mnemonic counts of lib/librpk.so (run here: cuobjdump needs no GPU).
    python tools/sass_evidence.py > profiles/rN_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "k8s-runpod-kubelet_b200", "lib", "librpk.so")
COLS = ["UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "LOP3", "LDS.128", "LDS", "CREDUX", "IMAD", "ATOMG", "ATOMS", "REDG", "RED", "MEMBAR", "LDG", "STG", "NANOSLEEP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur, i = collections.OrderedDict(), None, 0
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[i]).replace("void ", "")
            i += 1
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["instr"] += 1
            base = op.split(".")[0]
            kernels[cur][base] += 1
            if op.startswith("LDS") and ".128" in op:
                kernels[cur]["LDS.128"] += 1
    print("SASS evidence (cuobjdump -sass lib/librpk.so, sm_100a): per kernel, static instruction count and the mnemonics that show what it is built from.")
    print("UBLKCP = cp.async.bulk (TMA bulk copy into shared memory), SYNCS = mbarrier ops, ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents,")
    print("LOP3 = 3-input logic op (one LOP3 on mask words = 32 (pod, offer) pairs), LDS.128 = the 16-byte shared-memory loads of the persistent select kernel,")
    print("CREDUX = redux.sync, ATOMG / REDG / RED = global atomics (segment merge, tickets, class counts), ATOMS = shared-memory atomics (class histogram), MEMBAR = fences.\n")
    print(f"{'kernel':<64}{'instr':>7}" + "".join(f"{c:>10}" for c in COLS))
    for k in sorted(kernels):
        c = kernels[k]
        print(f"{k[:62]:<64}{c['instr']:>7}" + "".join(f"{c[x]:>10}" for x in COLS))


if __name__ == "__main__":
    main()
