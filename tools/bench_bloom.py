#!/usr/bin/env python
"""Throughput of the GPU bloom pre-filter next to flare's own
SaltedBloomFilter on the host when --cpu-library names a CPU build of the same C ABI (the
test oracle; this script never picks one up by itself).
Host buffers, copies included (yd_bloom_possibly_contains is synchronous)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from bloom_cases import tu_keys  # noqa: E402
from yadcc_b200 import TaskDispatcher  # noqa: E402


def run(lib, n, reps):
    d = TaskDispatcher(lib)
    keys = tu_keys(6124)
    d.bloom_reset()
    d.bloom_add(keys[:2000])
    m = d._key_matrix([keys[i % len(keys)] for i in range(n)])
    d.bloom_possibly_contains(m[:1000])
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        hit = d.bloom_possibly_contains(m)
        ts.append(time.perf_counter() - t0)
    return n / min(ts), float(hit.mean()), m.shape[1]


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("keys", nargs="?", type=int, default=1_000_000)
    ap.add_argument("--cpu-library", default=None, help="CPU library speaking the same C ABI, timed beside the GPU")
    args = ap.parse_args()
    n = args.keys
    gpu, hit, klen = run(None, n, 10)
    line = {"metric": "bloom_lookups_per_sec", "keys": n, "key_bytes": klen, "hashes": 10, "bits": 1 << 25,
            "gpu_e2e_keys_per_s": gpu, "gpu_e2e_GBps_of_keys": gpu * klen / 1e9, "hit_rate": hit}
    if args.cpu_library:
        cpu, _, _ = run(args.cpu_library, min(n, 200_000), 3)
        line.update({"cpu_reference_keys_per_s": cpu, "cpu_impl": Path(args.cpu_library).name})
    print(json.dumps(line))
