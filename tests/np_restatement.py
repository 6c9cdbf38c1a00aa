"""Second, independently written restatement of the reference's hot path (numpy, vectorised over offers).

The reference has no test that pins selection results or the status changed-set, and Go
cannot run here, so the C oracle (oracle/rpk_oracle.c) is cross-checked against this file: two restatements
written separately from the same cited Go lines must agree on every input.
"""
import numpy as np


def get_gpu_types(offers, min_ram, max_price, cloud, req_vcpu=0, req_ram=0):
    """runpod_client.go:465-509 for one pod.  cloud: 'SECURE' | 'COMMUNITY' | anything else."""
    G = len(offers["mem_gb"])
    if cloud == "SECURE":  # :469-471
        price, ok = offers["secure_price"], (offers["flags"] & 1).astype(bool)
    elif cloud == "COMMUNITY":  # :472-475
        price, ok = offers["community_price"], (offers["flags"] & 2).astype(bool)
    else:  # neither: price 0, cloudCheck false
        price, ok = np.zeros(G), np.zeros(G, bool)
    vcpu = offers.get("vcpu")
    ram = offers.get("ram_gb")
    vcpu = np.zeros(G, np.int64) if vcpu is None else vcpu.astype(np.int64)
    ram = np.zeros(G, np.int64) if ram is None else ram.astype(np.int64)
    with np.errstate(invalid="ignore"):
        keep = ok & (price > 0) & (price < max_price) & (offers["mem_gb"].astype(np.int64) >= min_ram)  # :478
    keep &= (vcpu >= req_vcpu) & (ram >= req_ram)
    idx = np.nonzero(keep)[0]
    order = np.argsort(price[idx], kind="stable")  # :497-500 with the tie contract: lowest index first
    return [int(i) for i in idx[order][:5]]  # :503-509


def select(offers, pods):
    P = len(pods["req_mem_gb"])
    best = np.full(P, -1, np.int32)
    top5 = np.full((P, 5), -1, np.int32)
    names = {0: "SECURE", 1: "COMMUNITY"}
    for p in range(P):
        cloud = names.get(int(pods["cloud"][p]) if pods.get("cloud") is not None else 0, "OTHER")
        mp = float(pods["max_price"][p]) if pods.get("max_price") is not None else 0.5
        rv = int(pods["req_vcpu"][p]) if pods.get("req_vcpu") is not None else 0
        rr = int(pods["req_ram_gb"][p]) if pods.get("req_ram_gb") is not None else 0
        ids = get_gpu_types(offers, int(pods["req_mem_gb"][p]), mp, cloud, rv, rr)
        top5[p, : len(ids)] = ids
        if ids:
            best[p] = ids[0]
    return best, top5


def decode_record(rec):
    """[len | flag << 7][status][0][ports][pad] -> (status bytes, ports bool); the flag is not a compared field"""
    ln = int(rec[0]) & 0x7F
    if ln < 2:
        return b"", False
    return bytes(rec[1 : ln - 1]), bool(rec[ln])


def status_changed_set(records, prev_state):
    """kubelet.go:870-880 on decoded fields.  prev_state: list of (status, ports) or None (never seen)."""
    out = []
    for i in range(records.shape[0]):
        cur = decode_record(records[i])
        if prev_state[i] is None or cur[0] != prev_state[i][0] or cur[1] != prev_state[i][1]:
            prev_state[i] = cur
            out.append(i)
    return np.array(out, np.uint32)
