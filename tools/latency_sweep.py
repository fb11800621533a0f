#!/usr/bin/env python
"""Dispatch latency (request enqueue -> grant available to the caller) as a function of the batch
size, through the C-ABI call with pinned host buffers on BASELINE configs[1]'s cluster
(2 000 servants, 8 digests).  One line of JSON.  With --cpu-library (a CPU build of the same C ABI,
i.e. the test oracle -- never picked up implicitly) the reference's latency on the same host is
printed beside it (single thread: it serialises on allocation_lock_)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from yadcc_b200 import TaskDispatcher  # noqa: E402
from yadcc_b200 import streams as S  # noqa: E402


def sweep(lib, sizes, reps):
    d = TaskDispatcher(lib) if lib else TaskDispatcher()
    w = S.config2(100_000, 2000, 8, variant="mod")
    w.register(d)
    src = w.build_requests(d)
    rows = []
    for n in sizes:
        reqs = d.alloc_requests(n)
        reqs[...] = src[:n]
        out = d.alloc_grants(n)
        ts = []
        for it in range(reps + 5):
            t0 = time.perf_counter()
            g = d.wait_for_starting_new_tasks(reqs, 1.0 + it, out=out)
            t1 = time.perf_counter()
            d.free_tasks(g["task_id"][g["status"] == 2].copy())
            if it >= 5:
                ts.append(1e3 * (t1 - t0))
        ts = np.sort(np.asarray(ts))
        rows.append({"batch": n, "p50_ms": round(float(ts[len(ts) // 2]), 4), "p99_ms": round(float(ts[int(len(ts) * 0.99)]), 4),
                     "decisions_per_s": round(n / (float(ts.mean()) / 1e3))})
    return rows


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-library", default=None)
    args = ap.parse_args()
    gpu = sweep(None, [1, 16, 256, 1024, 4096, 16384, 65536, 100_000], 200)
    line = {"metric": "dispatch_latency_vs_batch", "cluster": "2000 servants x 8 digests (cfg2-mod)", "gpu": gpu,
            "note": "latency of the whole call = latency seen by every request of the batch; L2 not flushed between calls"}
    if args.cpu_library:
        line.update({"cpu_reference": sweep(args.cpu_library, [1, 256, 4096], 5), "cpu_impl": Path(args.cpu_library).name})
    print(json.dumps(line))
