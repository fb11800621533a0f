"""Shared cases for the in-flight task index (RunningTaskKeeper,
yadcc/daemon/local/running_task_keeper.cc:40-75)."""
import numpy as np

from yadcc_b200 import RunningTask, Servant

GIB = 1 << 30


def _servant(i):
    return Servant(f"10.1.{i // 250}.{i % 250 + 1}:8335", None, ["d" * 64], version=8, num_processors=32,
                   max_tasks=16, total_memory_in_bytes=64 * GIB, memory_available_in_bytes=50 * GIB)


def task_digests(n, seed=7):
    """GetCxxTaskDigest = hex(BLAKE3(...)): 64 lowercase hex characters (yadcc/daemon/task_digest.cc:25-31)."""
    rng = np.random.default_rng(seed)
    return [rng.bytes(32).hex() for _ in range(n)]


def populate(d, n_servants, tasks_per_servant, pool, seed=0):
    """Heartbeats, grants, then running-task reports for the granted tasks (only tasks the
    dispatcher granted to that servant are kept, task_dispatcher.cc:256-273); digests drawn
    with repeats from `pool`.  A few reports carry grant ids nobody issued."""
    rng = np.random.default_rng(seed)
    svs = [_servant(i) for i in range(n_servants)]
    for sv in svs:
        d.keep_servant_alive(sv, 10, now=0.0)
    reqs = d.make_requests(n_servants * tasks_per_servant // 2, "d" * 64, "10.9.9.9", min_version=0, expires_in=300)
    g = d.wait_for_starting_new_tasks(reqs, 0.0)
    for i, sv in enumerate(svs):
        mine = g["task_id"][(g["status"] == 2) & (g["servant_index"] == i)]
        tasks = [RunningTask(int(rng.integers(1, 1 << 40)), int(t), sv.observed_location,
                             pool[int(rng.integers(0, len(pool)))]) for t in mine]
        if i % 5 == 0:
            tasks.insert(len(tasks) // 2, RunningTask(7, 1 << 50, sv.observed_location, pool[0]))  # unknown grant
        rng.shuffle(tasks)
        d.notify_servant_running_tasks(sv.observed_location, tasks)


def run_suite(d, seed=0, n_servants=40, n_queries=3000):
    """Everything observable about refresh + lookups, as arrays."""
    out = []
    rng = np.random.default_rng(100 + seed)
    # nothing refreshed yet: nothing is found
    pool = task_digests(300, seed)
    out.append(d.find_running_tasks(pool[:10]))
    populate(d, n_servants, 8, pool, seed)
    out.append(d.find_running_tasks(pool[:10]))  # reports are in, Refresh() has not run: still the old (empty) map
    n = d.running_index_refresh()
    snap = d.get_running_tasks()
    assert n == len(snap)
    out.append(np.asarray([n, d.running_index_size()]))
    absent = task_digests(200, 999 + seed)
    q = [pool[int(i)] for i in rng.integers(0, len(pool), n_queries)] + absent
    hits = d.find_running_tasks(q)
    out.append(hits)
    # the winning entry is a real snapshot entry with that digest, and it is the LAST one
    for key, h in list(zip(q, hits))[:400]:
        same = [i for i, t in enumerate(snap) if t.task_digest == key]
        if same:
            assert h["found"] == 1 and h["snapshot_index"] == same[-1]
            e = d.running_index_entry(int(h["snapshot_index"]))
            assert e.task_digest == key and e.servant_task_id == h["servant_task_id"] == snap[same[-1]].servant_task_id
            assert e.servant_location == snap[same[-1]].servant_location
        else:
            assert h["found"] == 0 and h["snapshot_index"] == 0xFFFFFFFF
    assert d.running_index_entry(n) is None
    # other key lengths never match 64-character digests; prefixes do not match either
    out.append(d.find_running_tasks([k[:63] for k in pool[:50]]))
    out.append(d.find_running_tasks([k + "0" for k in pool[:50]]))
    # a servant stops reporting; the map only changes at the next Refresh()
    sv0 = _servant(0)
    d.notify_servant_running_tasks(sv0.observed_location, [])
    out.append(d.find_running_tasks(q[:500]))
    d.running_index_refresh()
    out.append(d.find_running_tasks(q[:500]))
    # servants expire -> bookkeeper drops them (task_dispatcher.cc:510-511)
    for i in range(n_servants // 2):
        d.keep_servant_alive(_servant(i), 100, now=5.0)
    d.on_expiration_timer(now=20.0)
    out.append(np.asarray([d.running_index_refresh(), d.running_index_size()]))
    out.append(d.find_running_tasks(q[:500]))
    return out


def reference_test_case(d):
    """running_task_keeper_test.cc:36-66: three tasks 'task digest0..2' are found after a
    refresh and gone once the scheduler stops listing them.  (The mock scheduler of that test
    lists the tasks unconditionally; a real one only lists tasks it granted, so grant first.)"""
    sv = _servant(0)
    d.keep_servant_alive(sv, 10, now=0.0)
    g = d.wait_for_starting_new_tasks(d.make_requests(3, "d" * 64, "10.9.9.9", min_version=0, expires_in=300), 0.0)
    assert (g["status"] == 2).all()
    d.notify_servant_running_tasks(sv.observed_location,
                                   [RunningTask(i, int(g["task_id"][i]), "", "task digest" + str(i)) for i in range(3)])
    d.running_index_refresh()
    first = d.find_running_tasks(["task digest" + str(i) for i in range(3)])
    d.notify_servant_running_tasks(sv.observed_location, [])
    d.running_index_refresh()
    second = d.find_running_tasks(["task digest" + str(i) for i in range(3)])
    return first, second
