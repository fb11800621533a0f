#!/usr/bin/env python
"""bench.py -- task-assignment decisions/s on BASELINE.json's configs[1]
(100 k pending tasks x 2 k servants, 8 compiler digests, uniform slots).

A *step* is one pass of the hot path over one batch: the whole 100 k-request
FIFO queue is offered to the scheduler (n sequential WaitForStartingNewTask
decisions, zero-wait).  Between steps, untimed, the previous step's grants are
freed (so every step starts from the same servant state) and L2 is flushed by
writing a 256 MiB buffer.

  value   decisions/s with the request batch already resident in HBM
          (yd_stage_requests, untimed) when the timed region starts: device time of the
          slot-table + assignment + task-id kernels of yd_wait_for_staged_tasks, CUDA
          events on the library's solve stream (yd_last_solve_stats).  Every step runs
          the pass twice: once this way, once for e2e.
  e2e     the same metric through the C-ABI call a scheduler front-end makes
          (yd_wait_for_starting_new_tasks) with pinned HOST buffers: H2D of the
          24 B requests, all kernels, D2H of the 16 B grants, host clock around
          the synchronous call.
  roofline  the solve pipeline (one CUDA graph); achieved = SURVEY 8(d) algorithmic
          bytes (36*S + 32 per decision) / its CUDA-event duration, plus the compulsory
          traffic view; peak = measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the reference's own TaskDispatcher (oracle/_ref, compiled verbatim)
          or, if that build is absent, the CPU restatement, on the same stream,
          one thread (the reference serialises on one lock).

`--impl reference` times that CPU implementation instead (rank 0 only).

N > 1 (torchrun): ONE logical scheduler whose digest<->servant components are sharded
over the ranks (the sharding the reference's authors propose at
task_dispatcher.h:286-288): rank r owns 8 digests / 2 k servants / 100 k requests of
the global queue.  Decisions need no collective.  Task ids are local*world+rank by default
(opaque lease tokens: unique and routable, no exchange); `--ids fifo` reproduces the
single-scheduler numbering with one NCCL all-reduce of the per-request grant flags per
solve (yadcc_b200/sharded.py).  Weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from yadcc_b200 import STATUS_GRANTED, TaskDispatcher  # noqa: E402
from yadcc_b200 import streams as S  # noqa: E402

METRIC = "task_assignment_decisions_per_sec"
UNIT = "decisions/s"


def build_workload(name: str, rank: int):
    seed = 42 + 1000 * rank
    if name == "cfg2-mod":
        return S.config2(100_000, 2000, 8, seed=seed, variant="mod")
    if name == "cfg2-random":
        return S.config2(100_000, 2000, 8, seed=seed, variant="random")
    if name == "cfg1":
        return S.config1(seed=seed)
    if name == "cfg-self":
        return S.config_self(seed=seed)
    if name == "cfg3":
        return S.config3(1_000_000, 4000, 8, seed=seed)
    if name == "cfg5":  # BASELINE configs[4] shape on ONE GPU: 10 M x 8 k
        return S.config5(10_000_000, 8000, 8, seed=seed)
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows: list[list[str]] = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][2]) if self.rows[0][2].replace(".", "").isdigit() else None,
                "samples": len(self.rows), "reasons": sorted(reasons)}


def ncu_traffic_bytes() -> tuple[float | None, str]:
    """DRAM bytes per solve from the newest committed ncu launch list (profiles/), if present."""
    for name in ("r1h_launches_summary.csv", "r1d_launches_summary.csv"):
        p = ROOT / "profiles" / name
        try:
            last = p.read_text().strip().splitlines()[-1]
            mb = float(last.split("DRAM traffic per solve:")[1].split("MB")[0])
            return mb * 1e6, f"profiles/{name} (sum of dram__bytes_read+write over the solve kernels, cold caches)"
        except Exception:
            continue
    return None, "no ncu capture"


def measured_hbm_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_run(workload_name: str, rank: int, steps: int, warmup: int):
    """Times the reference TaskDispatcher (or the port) on the host, single thread."""
    ref = ROOT / "oracle" / "_ref" / "libydref.so"
    kind = "reference"
    if not ref.exists():
        ref = ROOT / "oracle" / "libydoracle.so"
        kind = "port"
        if not ref.exists():
            subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "libydoracle.so"])
    w = build_workload(workload_name, rank)
    d = TaskDispatcher(str(ref))
    w.register(d, now=0.0, expires_in=3600.0)
    reqs = w.build_requests(d)
    times, granted = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        g = d.wait_for_starting_new_tasks(reqs, 0.001)
        t1 = time.perf_counter()
        ok = g["status"] == STATUS_GRANTED
        granted = int(ok.sum())
        d.free_tasks(g["task_id"][ok])
        d.on_expiration_timer(now=1.0 + it)
        if it >= warmup:
            times.append(t1 - t0)
    d.close()
    return kind, len(reqs), granted, times, w


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    steps, warmup = args.steps, min(args.warmup, 1)
    kind, n, granted, times, w = cpu_reference_run(args.workload, 0, steps, warmup)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"{args.workload}: {w.meta}", "decisions_per_step": n, "granted_per_step": granted},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{len(times)} x the full {n}-request queue, single thread "
                                   f"(the reference serialises on allocation_lock_), host has {os.cpu_count()} cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from yadcc_b200.sharded import ShardedDispatcher

    # One logical scheduler: rank r owns the components of workload r (its 8 digests, 2 k
    # servants, 100 k requests); global request i*world + r is rank r's i-th request.
    w = build_workload(args.workload, rank)
    fifo_ids = world > 1 and args.ids == "fifo"
    d = TaskDispatcher(device=local, solver=args.solver, id_stride=0 if (world == 1 or fifo_ids) else world,
                       id_offset=0 if (world == 1 or fifo_ids) else rank)
    assert d.backend == "cuda-sm100a"
    owner_map = {}
    for r in range(world):
        for dg in build_workload(args.workload, r).digests:
            owner_map[dg] = r
    sd = ShardedDispatcher(d, rank, world, device=dev, digest_owner=lambda dg, _w: owner_map[dg],
                           id_mode="fifo" if fifo_ids else "strided")
    for sv in w.servants:
        sd.keep_servant_alive(sv, 3600.0, now=0.0)
    src = w.build_requests(d)
    n = len(src)
    S_count = len(w.servants)
    reqs = d.alloc_requests(n)  # pinned host memory
    out = d.alloc_grants(n)
    reqs[...] = src
    owners = (np.arange(n * world) % world).astype(np.int64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    dev_ms, e2e_ms, launches = [], [], 0
    n_solves = 0
    granted = 0
    h2d = d2h = 0
    solver_used = 0
    prev_ids = None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(args.warmup + args.steps):
        now = 1.0 + it
        # -- untimed: return the previous step's grants, 1 Hz expiration tick, flush L2 ------
        if prev_ids is not None:
            (sd.free_tasks if world > 1 else d.free_tasks)(prev_ids)
        d.on_expiration_timer(now=now)
        flush.fill_(it & 0xFF)
        if it == args.warmup:
            barrier()
            sampler.start()
            t_wall0 = time.perf_counter()
        torch.cuda.synchronize(dev)
        # -- timed (e2e): one pass of the hot path over the whole queue, HOST buffers ----------
        t0 = time.perf_counter()
        if world == 1:
            g = d.wait_for_starting_new_tasks(reqs, now, out=out)
        else:
            g = sd.wait_for_starting_new_tasks(None, owners, reqs, now)
            torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        st = d.last_solve_stats()
        ok = g["status"] == STATUS_GRANTED
        prev_ids = g["task_id"][ok].copy()
        granted = int(ok.sum())
        if it >= args.warmup:
            e2e_ms.append(1e3 * (t1 - t0))
            launches += st["kernel_launches"]
            n_solves += 1
            h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
            solver_used = st["solver"]
        if fifo_ids:
            # the exchange variant has no staged form: device time = the same call's pipeline
            # + the grant-flag all-reduce and prefix sum
            if it >= args.warmup:
                pipeline = st["prep_ms"] + st["solve_ms"] + st["final_ms"]
                dev_ms.append(pipeline + max(0.0, 1e3 * (t1 - t0) - st["total_ms"]))
            continue
        # -- timed (value): the same pass with the queue already resident in HBM ---------------
        (sd.free_tasks if world > 1 else d.free_tasks)(prev_ids)
        d.on_expiration_timer(now=now)
        d.stage_requests(reqs)  # untimed: inputs are in HBM when the timed region starts
        flush.fill_(~it & 0xFF)
        torch.cuda.synchronize(dev)
        g = d.wait_for_staged_tasks(n, now, out=out)
        st = d.last_solve_stats()
        ok = g["status"] == STATUS_GRANTED
        prev_ids = g["task_id"][ok].copy()
        assert int(ok.sum()) == granted
        if it >= args.warmup:
            dev_ms.append(st["prep_ms"] + st["solve_ms"] + st["final_ms"])  # CUDA events on the solve stream
            launches += st["kernel_launches"]
            n_solves += 1
    barrier()
    t_wall1 = time.perf_counter()
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # max over ranks of the summed step times
    tot = torch.tensor([sum(dev_ms), sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    tot_dev, tot_e2e = (float(x) for x in tot.tolist())
    K = args.steps
    decisions_all = n * K * world
    value = decisions_all / (tot_dev / 1e3)
    e2e_value = decisions_all / (tot_e2e / 1e3)

    if rank == 0:
        # phase breakdown: same workload, kernels launched one by one (no graph) so that CUDA
        # events on the solve stream can separate slot-table / assignment / task-id phases
        phases = None
        if world == 1:
            dp = TaskDispatcher(device=local, solver=args.solver, graphs=False)
            w.register(dp, now=0.0, expires_in=3600.0)
            rq = dp.alloc_requests(n)
            rq[...] = w.build_requests(dp)
            acc = np.zeros(3)
            for it in range(6):
                gg = dp.wait_for_starting_new_tasks(rq, 1.0 + it)
                dp.free_tasks(gg["task_id"][gg["status"] == STATUS_GRANTED])
                dp.on_expiration_timer(now=1.5 + it)
                if it >= 3:
                    s2 = dp.last_solve_stats()
                    acc += [s2["prep_ms"], s2["solve_ms"], s2["final_ms"]]
            phases = {"slot_table_ms": acc[0] / 3, "assignment_ms": acc[1] / 3, "task_ids_ms": acc[2] / 3,
                      "how": "un-graphed launches, CUDA events on the solve stream, mean of 3 steps"}
            dp.close()

        peak, peak_src = measured_hbm_peak()
        ms_step = tot_dev / K
        model_bytes = 36 * S_count + 32  # SURVEY.md 8(d) matrix-row model, per decision
        compulsory = 24 * n + 16 * n + 36 * S_count  # requests in, grants out, one servant-table read
        achieved_model = n * model_bytes / (ms_step / 1e3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            kind, cn, cgr, ctimes, _ = cpu_reference_run(args.workload, 0, max(1, min(3, K)), 0)
            cpu = {"value": cn * len(ctimes) / sum(ctimes), "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"{len(ctimes)} x the full {cn}-request queue ({sum(ctimes):.2f} s), single thread; "
                             f"the reference serialises on allocation_lock_ (host has {os.cpu_count()} cores)"}
        solver_name = {1: "row-scan", 2: "slot-stream"}.get(solver_used, str(solver_used))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w.meta} per GPU", "decisions_per_step_per_gpu": n,
                       "granted_per_step_per_gpu": granted,
                       "parallelism": f"digest<->servant components sharded over {world} rank(s)"
                                      + ("; task ids = local*world+rank, no collective" if world > 1 and not fifo_ids else "")
                                      + ("; one all-reduce of grant flags per solve for single-scheduler task ids" if fifo_ids else ""),
                       "l2": "flushed between steps (256 MiB write)", "solver": solver_name,
                       "between_steps_untimed": "FreeTask of the previous grants + OnExpirationTimer tick"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": tot_e2e / K,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": f"{solver_name} solve pipeline (one CUDA graph, {launches // max(1, n_solves)} kernels)",
                         "achieved": achieved_model, "peak": peak, "unit": "GB/s", "frac": achieved_model / peak,
                         "traffic": ncu_traffic_bytes()[0] if args.workload == "cfg2-mod" else None,
                         "traffic_source": ncu_traffic_bytes()[1], "peak_source": peak_src,
                         "algorithmic_bytes_per_decision": model_bytes, "kernel_ms_per_step": ms_step,
                         "compulsory": {"bytes_per_step": compulsory,
                                        "achieved": compulsory / (ms_step / 1e3) / 1e9, "unit": "GB/s",
                                        "frac": compulsory / (ms_step / 1e3) / 1e9 / peak},
                         "note": "achieved uses SURVEY 8(d)'s matrix-row model (36*S+32 B per decision = what the "
                                 "reference's O(S)-per-decision scan touches). The slot-stream solver is O(1) per "
                                 "decision, so the model over-counts by design and frac can exceed 1; 'compulsory' "
                                 "(requests in + grants out + one servant-table read) is the traffic a solve really "
                                 "needs. At 100k x 2k the pipeline is kernel-latency bound, not bandwidth bound: see "
                                 "DESIGN.md section 5 and profiles/."},
            "phases_ms": phases,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "wall_s_timed_loop": t_wall1 - t_wall0,
            "e2e_ms_steps": {"min": round(min(e2e_ms), 4), "median": round(sorted(e2e_ms)[len(e2e_ms) // 2], 4),
                             "max": round(max(e2e_ms), 4),
                             "note": "per-step host-clock times of the e2e call; nvidia-smi clock sampling during the timed region causes the rare ms-long outlier"},
            "latency_ms": {"p50": float(np.percentile(e2e_ms, 50)), "p99": float(np.percentile(e2e_ms, 99)),
                           "what": "enqueue->grant for every request of the batch (whole-batch call)"},
        }
        print(json.dumps(line))
    d.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2-mod")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--solver", type=int, default=0, help="0 auto, 1 row-scan, 2 slot-stream")
    ap.add_argument("--ids", default="strided", choices=["strided", "fifo"],
                    help="N>1 task-id space: strided (no exchange) or fifo (single-scheduler numbering, one all-reduce per solve)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
