"""Seeded synthetic event streams for the scheduler hot path.

A *stream* is a list of events; `Replayer.run` feeds them to any backend that
speaks the ydsched C ABI and returns a *trace* (one numpy array per event that
produces output).  Two backends are at parity iff their traces are equal
element-wise; `trace_digest` folds a trace into one SHA-256 for large runs.

Event kinds (tuples, first element is the kind):

  ("hb", now, Servant, expires_in)            one KeepServantAlive
  ("enqueue", REQ array)                      append to the FIFO pending queue
  ("solve", now)                              offer the whole pending queue in order
                                              (zero-wait); Timeout requests stay
                                              pending, everything else leaves
  ("wait", now, REQ array)                    one-shot batch, no pending queue
  ("free", ids)                               FreeTask per id
  ("free_frac", seed, frac)                   free a seeded subset of outstanding grants
  ("keepalive", now, ids | None, expires_in)  KeepTaskAlive (None = all outstanding)
  ("tick", now)                               OnExpirationTimer
  ("notify", location, [(servant_task_id, grant_id, digest)])
  ("notify_own", servant_index, drop_seed, extra_ids)
                                              heartbeat reporting the grants this
                                              servant holds (minus a seeded few, plus
                                              some bogus ids)
  ("running",)                                GetRunningTasks
  ("state",)                                  per-servant bookkeeping snapshot

Task ids are the ordinal of the grant (the reference starts at 0 and increments
per grant, task_dispatcher.h:218, .cc:127), so "free"/"keepalive" events can
name ids before the stream is run.

The request distributions for the five BASELINE.json configs are built by
`config1` .. `config5`; `fuzz_stream` mixes every quirk at small scale.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field, replace
from typing import Any, Callable, Sequence

import numpy as np

from . import _abi
from ._abi import GRANT_DTYPE, REQ_DTYPE, STATUS_GRANTED, STATUS_TIMEOUT
from .dispatcher import RunningTask, Servant, TaskDispatcher

GiB = 1 << 30


def hex_digest(rng: np.random.Generator) -> str:
    """A BLAKE3-looking compiler digest: 64 lowercase hex chars (env_desc.proto:27-28)."""
    return rng.bytes(32).hex()


def servant_ip(i: int) -> str:
    return f"10.{(i >> 16) & 255}.{(i >> 8) & 255}.{i & 255}"


@dataclass
class Stream:
    name: str
    events: list
    meta: dict = field(default_factory=dict)


class Replayer:
    """Drives a TaskDispatcher with a Stream and records everything it returns."""

    def __init__(self, dispatcher: TaskDispatcher, *, pinned: bool = False, on_solve: Callable | None = None,
                 batch_heartbeats: bool = False, packed: bool = False):
        """`batch_heartbeats`: runs of consecutive "hb" events with one timestamp go through
        keep_servants_alive, runs of consecutive "notify"/"notify_own" events through
        notify_servants_running_tasks (one call each); the trace is the same by definition."""
        self.d = dispatcher
        self.batch_heartbeats = batch_heartbeats
        self.packed = packed  # solves go through yd_wait_for_starting_new_tasks_packed (16-byte requests, 8-byte grants)
        self.pinned = pinned
        self.on_solve = on_solve
        self.pending = np.zeros(0, dtype=REQ_DTYPE)
        self.outstanding: dict[int, int] = {}  # task id -> servant index at grant time
        self.decisions = 0
        self.granted = 0
        self.solve_calls = 0

    def _wait(self, now: float, reqs: np.ndarray) -> np.ndarray:
        if self.packed:
            from .dispatcher import pack_requests
            if self.pinned:
                buf = pack_requests(reqs, self.d.alloc_requests16(len(reqs)))
                g = self.d.wait_for_starting_new_tasks_packed(buf, now, out8=self.d.alloc_grants8(len(reqs)))
            else:
                g = self.d.wait_for_starting_new_tasks_packed(pack_requests(reqs), now)
        elif self.pinned:
            buf = self.d.alloc_requests(len(reqs))
            buf[...] = reqs
            out = self.d.alloc_grants(len(reqs))
            g = self.d.wait_for_starting_new_tasks(buf, now, out=out).copy()
        else:
            g = self.d.wait_for_starting_new_tasks(np.ascontiguousarray(reqs), now).copy()
        self.decisions += len(reqs)
        ok = g["status"] == STATUS_GRANTED
        self.granted += int(ok.sum())
        self.solve_calls += 1
        for tid, sidx in zip(g["task_id"][ok].tolist(), g["servant_index"][ok].tolist()):
            self.outstanding[tid] = sidx
        if self.on_solve:
            self.on_solve(self.d, reqs, g)
        return g

    def _notify_args(self, ev):
        """(location, tasks) of a notify / notify_own event, or None if the servant index is gone."""
        d = self.d
        if ev[0] == "notify":
            _, loc, tasks = ev
            return loc, [RunningTask(a, b, loc, c) for a, b, c in tasks]
        _, sidx, drop_seed, extra = ev
        loc = d.servant_location(sidx)
        if loc is None:
            return None
        own = sorted(t for t, s in self.outstanding.items() if s == sidx)
        rng = np.random.default_rng(drop_seed)
        own = [t for t in own if rng.random() < 0.8]
        ids = own + list(extra)
        return loc, [RunningTask(1000 + k, t, loc, f"{t:064x}") for k, t in enumerate(ids)]

    def run(self, stream: Stream) -> list[np.ndarray]:
        d = self.d
        trace: list[np.ndarray] = []
        events = stream.events
        pos = 0
        while pos < len(events):
            ev = events[pos]
            pos += 1
            kind = ev[0]
            if self.batch_heartbeats and kind == "hb":
                run = [ev]
                while pos < len(events) and events[pos][0] == "hb" and events[pos][1] == ev[1]:
                    run.append(events[pos])
                    pos += 1
                d.keep_servants_alive([e[2] for e in run], [e[3] for e in run], now=ev[1])
                continue
            if self.batch_heartbeats and kind in ("notify", "notify_own"):
                run = [ev]
                while pos < len(events) and events[pos][0] in ("notify", "notify_own"):
                    run.append(events[pos])
                    pos += 1
                # (servant indices and outstanding grants do not change inside the run: arguments up front)
                args = [self._notify_args(e) for e in run]
                res = iter(d.notify_servants_running_tasks([a for a in args if a is not None]))
                for a in args:
                    trace.append(np.asarray(next(res) if a is not None else [], dtype=np.uint64))
                continue
            if kind == "hb":
                _, now, sv, exp = ev
                d.keep_servant_alive(sv, exp, now=now)
            elif kind == "enqueue":
                self.pending = np.concatenate([self.pending, ev[1]])
            elif kind == "solve":
                g = self._wait(ev[1], self.pending)
                self.pending = self.pending[g["status"] == STATUS_TIMEOUT]
                trace.append(g)
            elif kind == "wait":
                trace.append(self._wait(ev[1], ev[2]))
            elif kind == "free":
                ids = np.asarray(ev[1], dtype=np.uint64)
                d.free_tasks(ids)
                for i in ids.tolist():
                    self.outstanding.pop(i, None)
            elif kind == "free_frac":
                _, seed, frac = ev
                ids = np.fromiter(sorted(self.outstanding), dtype=np.uint64, count=len(self.outstanding))
                rng = np.random.default_rng(seed)
                pick = ids[rng.random(len(ids)) < frac]
                d.free_tasks(pick)
                for i in pick.tolist():
                    self.outstanding.pop(i, None)
                trace.append(pick.copy())
            elif kind == "keepalive":
                _, now, ids, exp = ev
                if ids is None:
                    ids = sorted(self.outstanding)
                ok = d.keep_tasks_alive(np.asarray(ids, dtype=np.uint64), exp, now=now)
                trace.append(ok.astype(np.uint8))
            elif kind == "tick":
                d.on_expiration_timer(now=ev[1])
            elif kind == "notify":
                _, loc, tasks = ev
                unknown = d.notify_servant_running_tasks(
                    loc, [RunningTask(a, b, loc, c) for a, b, c in tasks]
                )
                trace.append(np.asarray(unknown, dtype=np.uint64))
            elif kind == "notify_own":
                _, sidx, drop_seed, extra = ev
                loc = d.servant_location(sidx)
                if loc is None:
                    trace.append(np.zeros(0, dtype=np.uint64))
                    continue
                own = sorted(t for t, s in self.outstanding.items() if s == sidx)
                rng = np.random.default_rng(drop_seed)
                own = [t for t in own if rng.random() < 0.8]
                ids = own + list(extra)
                unknown = d.notify_servant_running_tasks(
                    loc, [RunningTask(1000 + k, t, loc, f"{t:064x}") for k, t in enumerate(ids)]
                )
                trace.append(np.asarray(unknown, dtype=np.uint64))
            elif kind == "running":
                rt = d.get_running_tasks()
                trace.append(
                    np.asarray([(t.servant_task_id, t.task_grant_id) for t in rt], dtype=np.uint64).reshape(-1, 2)
                )
            elif kind == "state":
                st = d.servant_state()
                trace.append(
                    np.stack([st["running_tasks"], st["ever_assigned_tasks"], st["capacity_available"]], axis=1)
                    if len(st)
                    else np.zeros((0, 3), dtype=np.uint64)
                )
                trace.append(np.asarray([d.next_task_id(), d.num_tasks(), d.num_servants()], dtype=np.uint64))
            else:  # pragma: no cover
                raise ValueError(f"unknown event {kind!r}")
        return trace


def trace_digest(trace: Sequence[np.ndarray]) -> str:
    h = hashlib.sha256()
    for a in trace:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype.descr).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def traces_equal(a: Sequence[np.ndarray], b: Sequence[np.ndarray]) -> bool:
    return len(a) == len(b) and all(x.shape == y.shape and x.dtype == y.dtype and (x == y).all() for x, y in zip(a, b))


def first_mismatch(a: Sequence[np.ndarray], b: Sequence[np.ndarray]) -> str:
    for k, (x, y) in enumerate(zip(a, b)):
        if x.shape != y.shape or x.dtype != y.dtype:
            return f"event-output {k}: shape/dtype {x.shape}/{x.dtype} vs {y.shape}/{y.dtype}"
        if not x.size:
            continue
        neq = np.nonzero(np.asarray(x != y).reshape(len(x), -1).any(axis=1))[0] if x.ndim else np.array([0])
        if (x != y).any():
            i = int(neq[0])
            return f"event-output {k}, row {i}: {x[i]!r} vs {y[i]!r} ({len(neq)} rows differ)"
    if len(a) != len(b):
        return f"trace lengths {len(a)} vs {len(b)}"
    return "equal"


# ---------------------------------------------------------------------------
# request / servant builders
# ---------------------------------------------------------------------------


def _requests(d: TaskDispatcher, env_ids: np.ndarray, ip_ids: np.ndarray, min_version, expires_in_s=15.0,
              prefetch=None) -> np.ndarray:
    r = np.zeros(len(env_ids), dtype=REQ_DTYPE)
    r["env_id"] = env_ids
    r["requestor_ip"] = ip_ids
    r["min_version"] = min_version
    r["expires_in_ns"] = int(expires_in_s * 1e9)
    if prefetch is not None:
        r["flags"] = np.where(prefetch, _abi.REQ_FLAG_PREFETCH, 0)
    return r


@dataclass
class Workload:
    """A config: servants to register and a function building its request queue."""

    name: str
    servants: list[Servant]
    digests: list[str]
    build_requests: Callable[[TaskDispatcher], np.ndarray]
    meta: dict = field(default_factory=dict)

    def register(self, d: TaskDispatcher, now: float = 0.0, expires_in: float = 10.0) -> None:
        for sv in self.servants:
            d.keep_servant_alive(sv, expires_in, now=now)

    def stream(self, d: TaskDispatcher) -> Stream:
        """Heartbeat x S, then one solve over the whole queue."""
        ev: list = [("hb", 0.0, sv, 10.0) for sv in self.servants]
        ev.append(("enqueue", self.build_requests(d)))
        ev.append(("solve", 0.001))
        ev.append(("state",))
        return Stream(self.name, ev, dict(self.meta))


def config1(n_tasks: int = 1000, n_servants: int = 64, seed: int = 42) -> Workload:
    """cfg 1: 1 k x 64, one digest held by everyone, uniform slots."""
    rng = np.random.default_rng(seed)
    dg = hex_digest(rng)
    servants = [
        Servant(f"{servant_ip(i)}:8335", None, [dg], 8, 32, 0, 64 * GiB, 50 * GiB, 16, _abi.PRIORITY_USER)
        for i in range(n_servants)
    ]

    def build(d: TaskDispatcher) -> np.ndarray:
        e = d.intern_env(dg)
        ips = np.asarray([d.intern_ip(f"172.16.{i >> 8}.{i & 255}") for i in range(256)], dtype=np.uint32)
        r = np.random.default_rng(seed + 1)
        return _requests(d, np.full(n_tasks, e, np.uint32), ips[r.integers(0, 256, n_tasks)], 8)

    return Workload("cfg1", servants, [dg], build, {"tasks": n_tasks, "servants": n_servants, "digests": 1})


def config2(n_tasks: int = 100_000, n_servants: int = 2000, n_digests: int = 8, seed: int = 42,
            variant: str = "mod", max_tasks: int = 64, nproc: int = 128) -> Workload:
    """cfg 2: 100 k x 2 k, 8 digests, uniform slots; all grant.

    variant "mod":    servant i holds digest i mod 8  (8 independent components)
    variant "random": 1-3 random digests each         (one coupled component)
    """
    rng = np.random.default_rng(seed)
    dgs = [hex_digest(rng) for _ in range(n_digests)]
    servants = []
    for i in range(n_servants):
        if variant == "mod":
            envs = [dgs[i % n_digests]]
        else:
            k = int(rng.integers(1, 4))
            envs = [dgs[j] for j in rng.choice(n_digests, size=k, replace=False)]
        servants.append(
            Servant(f"{servant_ip(i)}:8335", None, envs, 8, nproc, 0, 256 * GiB, 200 * GiB, max_tasks, _abi.PRIORITY_USER)
        )

    def build(d: TaskDispatcher) -> np.ndarray:
        env = np.asarray([d.intern_env(x) for x in dgs], dtype=np.uint32)
        ips = np.asarray([d.intern_ip(f"172.16.{i >> 8}.{i & 255}") for i in range(4096)], dtype=np.uint32)
        r = np.random.default_rng(seed + 1)
        return _requests(d, env[r.integers(0, n_digests, n_tasks)], ips[r.integers(0, 4096, n_tasks)], 8)

    return Workload(f"cfg2-{variant}", servants, dgs, build,
                    {"tasks": n_tasks, "servants": n_servants, "digests": n_digests, "variant": variant})


def config_self(n_tasks: int = 100_000, n_servants: int = 2000, seed: int = 44, run_len: int = 4) -> Workload:
    """Production-like: ONE compiler digest, every requestor is itself a servant (so the
    self-exclusion rule, task_dispatcher.cc:372-379, is live for every request) and requests
    arrive in runs of `run_len` from the same machine (immediate_reqs > 1 per RPC)."""
    rng = np.random.default_rng(seed)
    dg = hex_digest(rng)
    servants = [
        Servant(f"{servant_ip(i)}:8335", None, [dg], 8, 128, int(rng.integers(0, 20)), 256 * GiB, 200 * GiB, 51,
                _abi.PRIORITY_DEDICATED if i % 10 == 0 else _abi.PRIORITY_USER)
        for i in range(n_servants)
    ]

    def build(d: TaskDispatcher) -> np.ndarray:
        e = d.intern_env(dg)
        ips = np.asarray([d.intern_ip(servant_ip(i)) for i in range(n_servants)], dtype=np.uint32)
        r = np.random.default_rng(seed + 1)
        who = np.repeat(r.integers(0, n_servants, (n_tasks + run_len - 1) // run_len), run_len)[:n_tasks]
        return _requests(d, np.full(n_tasks, e, np.uint32), ips[who], 8)

    return Workload("cfg-self", servants, [dg], build,
                    {"tasks": n_tasks, "servants": n_servants, "digests": 1, "self": "every requestor is a servant"})


def _mixed_servants(n_servants: int, dgs: list[str], rng: np.random.Generator, envs_per: str = "mod") -> list[Servant]:
    """cfg 3/5 servant mix: nproc in {32,64,96,128}; USER 40% / DEDICATED 95% capacity
    (daemon/cloud/execution_engine.cc:132,153); load ~ U[0,nproc]; 15% low memory;
    5% not accepting; versions {7,8}."""
    out = []
    for i in range(n_servants):
        nproc = int(rng.choice([32, 64, 96, 128]))
        dedicated = rng.random() < 0.10
        mt = nproc * 95 // 100 if dedicated else nproc * 40 // 100
        if rng.random() < 0.05:
            mt = 0
        load = int(rng.integers(0, nproc + 1))
        lowmem = rng.random() < 0.15
        total = 64 * GiB
        avail = int(rng.integers(1, 10)) * GiB - 1 if lowmem else int(rng.integers(11, 60)) * GiB
        ver = 7 if rng.random() < 0.3 else 8
        if envs_per == "mod":
            envs = [dgs[i % len(dgs)]]
        else:
            k = int(rng.integers(1, 4))
            envs = [dgs[j] for j in rng.choice(len(dgs), size=min(k, len(dgs)), replace=False)]
        out.append(
            Servant(f"{servant_ip(i)}:8335", None, envs, ver, nproc, load, total, avail, mt,
                    _abi.PRIORITY_DEDICATED if dedicated else _abi.PRIORITY_USER)
        )
    return out


def config3(n_tasks: int = 1_000_000, n_servants: int = 4000, n_digests: int = 8, seed: int = 43,
            envs_per: str = "random") -> Workload:
    """cfg 3: 1 M x 4 k, mixed memory headroom / priority / versions / self-IP.

    Capacity < tasks, so `rounds_stream` interleaves solves with Free and Tick.
    """
    rng = np.random.default_rng(seed)
    dgs = [hex_digest(rng) for _ in range(n_digests)]
    servants = _mixed_servants(n_servants, dgs, rng, envs_per)

    def build(d: TaskDispatcher) -> np.ndarray:
        env = np.asarray([d.intern_env(x) for x in dgs], dtype=np.uint32)
        r = np.random.default_rng(seed + 1)
        outside = np.asarray([d.intern_ip(f"172.16.{i >> 8}.{i & 255}") for i in range(4096)], dtype=np.uint32)
        inside = np.asarray([d.intern_ip(servant_ip(i)) for i in range(n_servants)], dtype=np.uint32)
        shares = r.random(n_tasks) < 0.20  # 20% of requestors are servants themselves
        ip = np.where(shares, inside[r.integers(0, n_servants, n_tasks)], outside[r.integers(0, 4096, n_tasks)])
        mv = np.where(r.random(n_tasks) < 0.5, 7, 8).astype(np.uint32)
        return _requests(d, env[r.integers(0, n_digests, n_tasks)], ip.astype(np.uint32), mv)

    return Workload("cfg3", servants, dgs, build,
                    {"tasks": n_tasks, "servants": n_servants, "digests": n_digests, "envs_per": envs_per})


def config5(n_tasks: int = 10_000_000, n_servants: int = 8000, n_digests: int = 8, seed: int = 45) -> Workload:
    w = config3(n_tasks, n_servants, n_digests, seed)
    return replace(w, name="cfg5")


def rounds_stream(w: Workload, d: TaskDispatcher, max_rounds: int = 8, free_frac: float = 0.5) -> Stream:
    """cfg 3 interleaving: solve, free a seeded half of the outstanding grants,
    renew the rest, re-heartbeat, tick +1 s, re-offer what is still pending."""
    ev: list = [("hb", 0.0, sv, 10.0) for sv in w.servants]
    ev.append(("enqueue", w.build_requests(d)))
    t = 0.001
    for k in range(max_rounds):
        ev.append(("solve", t))
        ev.append(("free_frac", 1000 + k, free_frac))
        ev.append(("keepalive", t + 0.5, None, 15.0))
        t += 1.0
        for sv in w.servants:
            ev.append(("hb", t, sv, 10.0))
        ev.append(("tick", t))
        ev.append(("state",))
    return Stream(w.name + "-rounds", ev, dict(w.meta, rounds=max_rounds))


# ---------------------------------------------------------------------------
# fuzz: every quirk at small scale
# ---------------------------------------------------------------------------


def fuzz_stream(d: TaskDispatcher, seed: int, n_servants: int = 24, n_events: int = 60, max_batch: int = 40,
                wide: bool = False, unique_hosts: bool = False) -> Stream:
    """Random interleaving of all event kinds over a small cluster.

    Covers: several ports on one IP (only the first free one is 'self'),
    requestors that are servants, dedicated servants around the 50% mark, low
    memory, max_tasks == 0, load above nproc, version gating incl. negative
    versions, unknown environments, capacity shrinking below running_tasks,
    lease expiry -> zombies -> sweep on heartbeat, servant expiry -> orphans,
    freeing unknown / duplicate ids, heartbeats from unknown locations.
    `wide` adds capacities above 32768 (the wide-key solver path).  `unique_hosts` gives every
    servant its own IP (one daemon per machine): requestors that are servants then have exactly
    one "self" servant, which is the shape the merge solver takes on itself.
    """
    rng = np.random.default_rng(seed)
    dgs = [hex_digest(rng) for _ in range(int(rng.integers(1, 5)))]
    hosts = [f"10.0.{i >> 8}.{i & 255}" for i in range(n_servants if unique_hosts else max(2, n_servants // 2))]

    def rand_servant(i: int) -> Servant:
        host = hosts[i] if unique_hosts else hosts[int(rng.integers(0, len(hosts)))]
        nproc = int(rng.choice([0, 2, 4, 8, 16, 32, 40000 if wide else 24]))
        mt = int(rng.choice([0, 2, 3, 7, 8, 12, 70000 if wide else 30]))
        k = int(rng.integers(0 if rng.random() < 0.2 else 1, len(dgs) + 1))
        envs = [dgs[j] for j in rng.choice(len(dgs), size=k, replace=False)] if k else []
        if rng.random() < 0.1:
            envs = envs + envs[:1]  # duplicate digest in one heartbeat
        return Servant(
            f"{host}:{8000 + i}",
            None,
            envs,
            int(rng.choice([-1, 6, 7, 8, 8, 9, 9])),
            nproc,
            int(rng.integers(0, max(nproc, 1) + 3)) if rng.random() < 0.4 else int(rng.integers(0, 3)),
            int(rng.choice([0, 64 * GiB])),
            int(rng.choice([5 * GiB, 10 * GiB - 1, 10 * GiB, 40 * GiB, 40 * GiB, 40 * GiB])),
            mt,
            int(rng.choice([_abi.PRIORITY_USER, _abi.PRIORITY_DEDICATED, _abi.PRIORITY_USER])),
        )

    servants = [rand_servant(i) for i in range(n_servants)]
    ev: list = []
    now = 0.0
    for sv in servants:
        ev.append(("hb", now, sv, float(rng.choice([2.0, 5.0, 10.0]))))
    env_ids = [d.intern_env(x) for x in dgs] + [d.intern_env("not-a-known-digest")]
    ip_pool = hosts + ["172.16.0.1", "172.16.0.2", "10.0.0", ""]
    ip_ids = [d.intern_ip(x) for x in ip_pool]
    issued = 0  # upper bound on ids handed out so far (for picking ids to free / renew)
    for _ in range(n_events):
        now += float(rng.choice([0.0, 0.1, 0.4, 1.0]))
        kind = rng.choice(["wait", "wait", "wait", "free", "keepalive", "tick", "hb", "notify", "running", "solve"])
        if kind in ("wait", "solve"):
            n = int(rng.integers(0, max_batch + 1))
            # runs of identical requests, like one RPC with immediate_reqs > 1
            e = np.repeat(rng.choice(env_ids, size=n), 1)
            ip = rng.choice(ip_ids, size=n)
            if n and rng.random() < 0.5:
                run = int(rng.integers(1, n + 1))
                e[:run] = e[0]
                ip[:run] = ip[0]
            r = _requests(d, e.astype(np.uint32), ip.astype(np.uint32),
                          rng.choice([0, 7, 8, 9], size=n).astype(np.uint32),
                          expires_in_s=float(rng.choice([0.5, 1.5, 15.0])), prefetch=rng.random(n) < 0.3)
            if kind == "wait":
                ev.append(("wait", now, r))
            else:
                ev.append(("enqueue", r))
                ev.append(("solve", now))
            issued += n
        elif kind == "free":
            k = int(rng.integers(0, 12))
            ids = rng.integers(0, issued + 3, size=k)
            if k and rng.random() < 0.3:
                ids[-1] = ids[0]  # duplicate
            ev.append(("free", ids.astype(np.uint64)))
        elif kind == "keepalive":
            k = int(rng.integers(0, 12))
            ev.append(("keepalive", now, rng.integers(0, issued + 3, size=k).astype(np.uint64),
                       float(rng.choice([0.5, 2.0, 15.0]))))
        elif kind == "tick":
            ev.append(("tick", now))
        elif kind == "hb":
            i = int(rng.integers(0, n_servants))
            if rng.random() < 0.5:
                servants[i] = replace(rand_servant(i), observed_location=servants[i].observed_location)
            ev.append(("hb", now, servants[i], float(rng.choice([0.0, 2.0, 5.0, 10.0]))))
        elif kind == "notify":
            if rng.random() < 0.15:
                ev.append(("notify", "203.0.113.9:1", [(1, int(rng.integers(0, issued + 3)), "aa")]))
            else:
                extra = rng.integers(0, issued + 3, size=int(rng.integers(0, 3))).tolist()
                ev.append(("notify_own", int(rng.integers(0, n_servants)), int(rng.integers(0, 1 << 30)), extra))
        elif kind == "running":
            ev.append(("running",))
        if rng.random() < 0.25:
            ev.append(("state",))
    ev.append(("state",))
    return Stream(f"fuzz-{seed}", ev, {"seed": seed, "servants": n_servants})


# ---------------------------------------------------------------------------
# named streams (used by tests/golden and the parity tests)
# ---------------------------------------------------------------------------


def named_stream(name: str, d: TaskDispatcher) -> Stream:
    """Deterministic stream by name; the same names are keys in
    tests/golden/digests.json."""
    if name == "cfg1":
        return config1().stream(d)
    if name == "cfg2-mod-small":
        return config2(5000, 200, 8, variant="mod").stream(d)
    if name == "cfg2-random-small":
        return config2(5000, 200, 8, variant="random").stream(d)
    if name == "cfg2-mod":
        return config2(variant="mod").stream(d)
    if name == "cfg2-random":
        return config2(variant="random").stream(d)
    if name == "cfg-self-small":
        return config_self(6000, 150).stream(d)
    if name == "cfg-self":
        return config_self().stream(d)
    if name == "cfg3-small":
        return rounds_stream(config3(20000, 300, 8), d, max_rounds=4)
    if name == "cfg3-mod-small":
        return rounds_stream(config3(20000, 300, 8, envs_per="mod"), d, max_rounds=4)
    if name == "cfg3":
        return rounds_stream(config3(), d, max_rounds=3)
    if name == "cfg5-1m":  # BASELINE configs[4]'s servant pool and distributions, first 1 M requests of its queue
        return config5(1_000_000, 8000).stream(d)
    if name == "cfg5-1m-rounds":
        return rounds_stream(config5(1_000_000, 8000), d, max_rounds=2)
    if name.startswith("fuzz-"):
        seed = int(name.split("-")[1])
        return fuzz_stream(d, seed, n_servants=8 + seed % 30, wide=(seed % 5 == 0))
    raise KeyError(name)
