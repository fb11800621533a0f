import sys, time
sys.path.insert(0, '.')
import numpy as np
from yadcc_b200 import TaskDispatcher
from yadcc_b200 import streams as S
d = TaskDispatcher()
w = S.config2(100_000, 2000, 8, variant="mod")
w.register(d)
src = w.build_requests(d)
reqs = d.alloc_requests(len(src)); reqs[...] = src
out = d.alloc_grants(len(src))
prev = None
rows = []
for it in range(60):
    if prev is not None:
        d.free_tasks(prev)
    d.on_expiration_timer(now=1.0 + it)
    t0 = time.perf_counter()
    g = d.wait_for_starting_new_tasks(reqs, 1.0 + it, out=out)
    t1 = time.perf_counter()
    st = d.last_solve_stats()
    prev = g["task_id"].copy()
    if it >= 10:
        rows.append((1e3 * (t1 - t0), st["total_ms"], st["solve_ms"]))
a = np.asarray(rows)
print("host_ms total_ms(ev0..ev5) graph_ms(ev1..ev4):", a.mean(0).round(4), "min", a.min(0).round(4))
