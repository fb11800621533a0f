#!/usr/bin/env python
"""Throughput of the GPU in-flight task index next to the
RunningTaskKeeper loops on the host when --cpu-library names a CPU build of the same C ABI (the
test oracle; this script never picks one up by itself): Refresh() of a cluster-wide snapshot and TryFindTask for a pending queue.
Host buffers, copies included (both calls are synchronous)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from running_index_cases import populate, task_digests  # noqa: E402
from yadcc_b200 import TaskDispatcher  # noqa: E402


def run(lib, n_queries, reps):
    d = TaskDispatcher(lib)
    pool = task_digests(40_000, 3)
    populate(d, 2000, 32, pool, seed=1)  # 2000 servants x 16 grants each: 32 k in-flight tasks
    q = d._key_matrix([pool[i % len(pool)] for i in range(n_queries // 2)] + task_digests(n_queries - n_queries // 2, 4))
    tr, tf = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        n = d.running_index_refresh()
        tr.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        hits = d.find_running_tasks(q)
        tf.append(time.perf_counter() - t0)
    return n, d.running_index_size(), n / min(tr), len(q) / min(tf), float(hits["found"].mean())


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("queries", nargs="?", type=int, default=1_000_000)
    ap.add_argument("--cpu-library", default=None, help="CPU library speaking the same C ABI, timed beside the GPU")
    args = ap.parse_args()
    nq = args.queries
    n, distinct, g_refresh, g_find, hit = run(None, nq, 8)
    line = {"metric": "running_index", "snapshot_entries": n, "distinct_digests": distinct, "queries": nq, "key_bytes": 64,
            "hit_rate": hit, "gpu_refresh_entries_per_s": g_refresh, "gpu_find_keys_per_s": g_find}
    if args.cpu_library:
        _, _, c_refresh, c_find, _ = run(args.cpu_library, min(nq, 200_000), 3)
        line.update({"cpu_reference_refresh_entries_per_s": c_refresh, "cpu_reference_find_keys_per_s": c_find,
                     "cpu_impl": Path(args.cpu_library).name})
    print(json.dumps(line))
