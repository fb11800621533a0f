#!/usr/bin/env python
"""Phase stamps of the fused kernel (YDSCHED_FUSED_PROF=1), L2 flushed / warm: python tools/dev/phase_prof.py cfg2-mod"""
import os, sys, time
os.environ["YDSCHED_FUSED_PROF"] = "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch
from bench import build_workload
from yadcc_b200 import STATUS_GRANTED, TaskDispatcher, pack_requests, unpack_grants

for name in sys.argv[1:] or ["cfg2-mod"]:
    w = build_workload(name)
    d = TaskDispatcher()
    w.register(d, now=0.0, expires_in=3600.0)
    src = w.build_requests(d)
    reqs = d.alloc_requests(len(src)); reqs[...] = src
    out = d.alloc_grants(len(src))
    r16 = pack_requests(src, d.alloc_requests16(len(src)))
    o8 = d.alloc_grants8(len(src))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for it in range(10):
        mode = "staged" if it < 6 else "e2e-packed"
        cold = it >= 3
        if mode == "staged":
            d.stage_requests(reqs)
        if cold:
            flush.fill_(it); torch.cuda.synchronize()
        print(f"{name} {mode} {'cold' if cold else 'warm'}:", file=sys.stderr, end=" ", flush=True)
        t0 = time.perf_counter()
        if mode == "staged":
            g = d.wait_for_staged_tasks(len(src), 1.0 + it, out=out)
        else:
            g8, ids = d.wait_for_starting_new_tasks_packed(r16, 1.0 + it, out8=o8, unpack=False)
        t1 = time.perf_counter()
        if mode != "staged":
            g = unpack_grants(g8, ids)
        st = d.last_solve_stats()
        print(f"   device {1e3*(st['prep_ms']+st['solve_ms']+st['final_ms']):.1f} us total {st['total_ms']*1e3:.1f} host {1e6*(t1-t0):.1f} us launches {st['kernel_launches']}", file=sys.stderr, flush=True)
        d.free_tasks(g["task_id"][g["status"] == STATUS_GRANTED].copy())
        d.on_expiration_timer(now=1.5 + it)
    d.close()
