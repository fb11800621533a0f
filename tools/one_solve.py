#!/usr/bin/env python
"""A few solves of one workload (for ncu: `ncu ... python tools/one_solve.py cfg3 [n_solves]`)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import build_workload  # noqa: E402
from yadcc_b200 import STATUS_GRANTED, TaskDispatcher  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2-mod"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = build_workload(name)
d = TaskDispatcher()
w.register(d, now=0.0, expires_in=3600.0)
src = w.build_requests(d)
reqs = d.alloc_requests(len(src))
reqs[...] = src
out = d.alloc_grants(len(src))
for it in range(reps):
    d.stage_requests(reqs)
    g = d.wait_for_staged_tasks(len(src), 1.0 + it, out=out)
    d.free_tasks(g["task_id"][g["status"] == STATUS_GRANTED].copy())
    d.on_expiration_timer(now=1.5 + it)
print(name, d.last_solve_stats())
d.close()
