#!/usr/bin/env python
"""Device / end-to-end time of one full-queue solve per workload (development aid; bench.py is the contract)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from yadcc_b200 import STATUS_GRANTED, TaskDispatcher  # noqa: E402
from bench import build_workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["cfg2-mod", "cfg2-random", "cfg-self", "cfg3"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="compare the grants with the CPU restatement (small enough workloads)")
    args = ap.parse_args()
    for name in args.workloads:
        w = build_workload(name, 0)
        d = TaskDispatcher()
        w.register(d, now=0.0, expires_in=3600.0)
        src = w.build_requests(d)
        n = len(src)
        reqs = d.alloc_requests(n)
        out = d.alloc_grants(n)
        reqs[...] = src
        dev, e2e = [], []
        for it in range(args.steps + 2):
            d.stage_requests(reqs)
            g = d.wait_for_staged_tasks(n, 1.0 + it, out=out)
            st = d.last_solve_stats()
            ok = g["status"] == STATUS_GRANTED
            d.free_tasks(g["task_id"][ok].copy())
            d.on_expiration_timer(now=1.0 + it)
            t0 = time.perf_counter()
            g = d.wait_for_starting_new_tasks(reqs, 1.5 + it, out=out)
            t1 = time.perf_counter()
            ok = g["status"] == STATUS_GRANTED
            granted = int(ok.sum())
            if args.check and it == 0:
                o = TaskDispatcher(str(ROOT / "oracle" / "libydoracle.so"))
                w.register(o, now=0.0, expires_in=3600.0)
                go = o.wait_for_starting_new_tasks(w.build_requests(o), 1.5)
                same = bool((go["status"] == g["status"]).all() and (go["servant_index"] == g["servant_index"]).all())
                print(f"{name}: parity vs port: {same}", flush=True)
                o.close()
            d.free_tasks(g["task_id"][ok].copy())
            d.on_expiration_timer(now=1.6 + it)
            if it >= 2:
                dev.append(st["prep_ms"] + st["solve_ms"] + st["final_ms"])
                e2e.append(1e3 * (t1 - t0))
        print(json.dumps({"workload": name, "n": n, "granted": granted, "dev_ms": round(float(np.median(dev)), 4),
                          "dev_Mdps": round(n / np.median(dev) / 1e3, 1), "e2e_ms": round(float(np.median(e2e)), 4),
                          "launches": st["kernel_launches"], "solver": st["solver"]}), flush=True)
        d.close()


if __name__ == "__main__":
    main()
