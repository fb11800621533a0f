"""Shared builder for the WaitForStartingTask-expansion tests."""
import numpy as np

from yadcc_b200 import Servant, PRIORITY_DEDICATED, PRIORITY_USER
from yadcc_b200 import _abi


def run_rpc_stream(d, seed: int):
    """Registers a small cluster and plays a few batches of RPCs; returns everything
    the backend answered."""
    rng = np.random.default_rng(seed)
    digs = [f"{i:064x}" for i in range(4)]
    for i in range(int(rng.integers(5, 60))):
        envs = [digs[j] for j in rng.choice(3, size=int(rng.integers(1, 4)), replace=False)]  # digs[3] is held by nobody
        nproc = int(rng.choice([4, 8, 16]))
        d.keep_servant_alive(
            Servant(f"10.6.0.{i}:8335", None, envs, int(rng.choice([7, 8])), nproc, int(rng.integers(0, 4)), 0, 64 << 30,
                    int(rng.integers(0, nproc)), PRIORITY_DEDICATED if i % 5 == 0 else PRIORITY_USER), 10.0, now=0.0)
    out = []
    for b in range(4):
        n = int(rng.integers(1, 40))
        rpcs = np.zeros(n, dtype=_abi.RPC_WAIT_DTYPE)
        rpcs["env_id"] = [d.intern_env(digs[j]) for j in rng.integers(0, 4, n)]
        rpcs["min_version"] = rng.choice([0, 7, 8, 9], n)
        rpcs["requestor_ip"] = [d.intern_ip(f"10.6.0.{j}") for j in rng.integers(0, 80, n)]
        rpcs["immediate_reqs"] = rng.choice([0, 1, 1, 2, 5], n)
        rpcs["prefetch_reqs"] = rng.choice([0, 0, 1, 3], n)
        rpcs["milliseconds_to_wait"] = rng.choice([0, 5000, 10000, 10001], n, p=[0.3, 0.4, 0.2, 0.1])
        rpcs["next_keep_alive_ns"] = rng.choice([15_000_000_000, 30_000_000_000, 30_000_000_001], n, p=[0.6, 0.3, 0.1])
        res, grants = d.wait_for_starting_task_rpcs(rpcs, now=0.1 * (b + 1))
        out += [np.stack([res["status"], res["n_grants"], res["first_grant"]], 1), grants.copy()]
        ok = grants["task_id"]
        if len(ok):
            d.free_tasks(ok[:: 2])
        st = d.servant_state()
        out.append(st["running_tasks"].copy())
    return out
