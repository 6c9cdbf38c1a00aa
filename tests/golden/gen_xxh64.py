"""Generates tests/golden/xxh64_kat.json from python-xxhash (binding of the canonical C xxHash, the same
algorithm github.com/cespare/xxhash/v2 v2.1.2 implements: reference go.mod:60).  Run in the build
container; the JSON travels, this script documents how it was made."""
import json
import os

import xxhash

vec = []
named = [b"", b"a", b"abc", b"RUNNING\x00\x01", b"RUNNING\x00\x00", b"STARTING\x00\x00", b"EXITED\x00\x01",
         b"x" * 31, b"x" * 32, b"x" * 33, bytes(range(64))]
for d in named:
    vec.append({"hex": d.hex(), "seed": 0, "xxh64": format(xxhash.xxh64(d, seed=0).intdigest(), "016x")})
# every length 0..255 of a fixed byte pattern (straddles the <32 path, the 4-lane stripes and all tails)
pat = bytes((i * 131 + 7) & 0xFF for i in range(255))
for n in range(256):
    d = pat[:n]
    vec.append({"hex": d.hex(), "seed": 0, "xxh64": format(xxhash.xxh64(d, seed=0).intdigest(), "016x")})
for seed in (1, 0x9E3779B185EBCA87):
    vec.append({"hex": pat[:77].hex(), "seed": seed, "xxh64": format(xxhash.xxh64(pat[:77], seed=seed).intdigest(), "016x")})
# slot prefixes as the status kernels hash them (include/rpk.h): byte 0 = len, data, zero padding up to a multiple of
# 8 bytes -- every len 0..127 of a fixed pattern; the hashed input is bytes [0, 8*ceil((1+len)/8))
slot = []
pat2 = bytes((i * 89 + 41) & 0xFF for i in range(127))
for ln in range(128):
    nbytes = 8 * ((1 + ln + 7) // 8)
    d = (bytes([ln]) + pat2[:ln]).ljust(nbytes, b"\x00")
    slot.append({"len": ln, "hex": d.hex(), "xxh64": format(xxhash.xxh64(d, seed=0).intdigest(), "016x")})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "xxh64_kat.json")
json.dump({"source": f"python-xxhash {xxhash.VERSION} (xxHash {xxhash.XXHASH_VERSION})", "vectors": vec, "slot_vectors": slot},
          open(out, "w"), indent=0)
print(len(vec), "+", len(slot), "vectors ->", out)
