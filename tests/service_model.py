"""TEST-ONLY second restatement of SchedulerServiceImpl's handlers, in Python, written from
yadcc/scheduler/scheduler_service_impl.cc:55-333 and independent of include/ydservice_impl.inc /
include/ydsched_rpc_impl.inc (the C restatement every backend compiles): the two are compared on random
request streams in test_service.py, so that the C layer is not only compared with itself.

It sits on the Python mirror of `TaskDispatcher` (any backend) and makes ONE decision per
`WaitForStartingNewTask` call, literally as the handler's loops do (:234-264) -- it never touches the
batched RPC expansion.  Serving-daemon tokens are random (RAND_bytes, :46-51), so the model keeps symbolic
token ids: what is comparable is when the window of three rolls (:319-333).
"""
from __future__ import annotations

import socket
from dataclasses import dataclass

import numpy as np

from yadcc_b200 import Servant, STATUS_ENVIRONMENT_NOT_FOUND, STATUS_GRANTED
from yadcc_b200 import _abi

OK, NO_QUOTA, ACCESS_DENIED, INVALID_ARGUMENT, VERSION_TOO_OLD, ENV_NOT_AVAILABLE = 0, 1001, 1003, 1004, 1005, 1006
REASON_BEHIND_NAT, REASON_NOT_VERIFIED = 4, 100
PRIORITY_UNKNOWN, PRIORITY_USER, PRIORITY_DEDICATED = 0, 1, 2  # scheduler.proto ServantPriority


def token_set(flag: str) -> set[str]:
    """MakeTokenVerifierFromFlag, yadcc/common/token_verifier.cc:56-69: split on ',', empty entries kept."""
    return set(flag.split(","))


def parse_endpoint(s: str):
    """flare::TryParse<Endpoint> (flare/base/net/endpoint.cc:321-378): 'a.b.c.d:port', else '[v6]:port'.
    Returns (normalized text, port) or None."""
    def port_of(t: str):
        if not t or len(t) > 5 or not all("0" <= c <= "9" for c in t) or int(t) > 65535:
            return None
        return int(t)

    if ":" in s:
        host, _, p = s.partition(":")
        port = port_of(p)
        if port is not None:
            try:
                return socket.inet_ntop(socket.AF_INET, socket.inet_pton(socket.AF_INET, host)) + f":{port}", port
            except (OSError, ValueError):
                pass
    pos = s.rfind(":")
    if pos < 2:
        return None
    port = port_of(s[pos + 1:])
    if port is None:
        return None
    try:
        a6 = socket.inet_pton(socket.AF_INET6, s[1:pos - 1])
    except (OSError, ValueError):
        return None
    return "[" + socket.inet_ntop(socket.AF_INET6, a6) + f"]:{port}", port


@dataclass
class ModelHeartbeatResponse:
    status: int
    rolls: int          # how often the token window has rolled so far (0 on failure paths: not consulted)
    expired_tasks: list


class ServiceModel:
    def __init__(self, dispatcher, *, acceptable_user_tokens: str, acceptable_servant_tokens: str, min_daemon_version: int = 0,
                 serving_daemon_token_rollout_interval: int = 3600, now: float = 0.0):
        self.d = dispatcher
        self.user = token_set(acceptable_user_tokens)
        self.servant = token_set(acceptable_servant_tokens)
        self.min_version = min_daemon_version
        self.interval = serving_daemon_token_rollout_interval if serving_daemon_token_rollout_interval > 0 else 3600
        self.next_rollout = now + self.interval  # :62-64
        self.rolls = 0

    def _active_tokens(self, now: float) -> int:  # DetermineActiveServingDaemonTokens, :319-333
        if self.next_rollout < now:
            self.next_rollout = now + self.interval
            self.rolls += 1
        return self.rolls

    def heartbeat(self, req, *, now: float) -> ModelHeartbeatResponse:  # :67-194
        fail = lambda st: ModelHeartbeatResponse(st, 0, [])
        if req.token not in self.user and req.token not in self.servant:
            return fail(ACCESS_DENIED)
        if req.version < self.min_version:
            return fail(VERSION_TOO_OLD)
        ep = parse_endpoint(req.location)
        if ep is None:
            return fail(INVALID_ARGUMENT)
        reported, port = ep
        observed = f"[{req.remote_ip}]:{port}" if req.remote_is_ipv6 else f"{req.remote_ip}:{port}"
        if req.next_heartbeat_in_ms > 30_000:
            return fail(INVALID_ARGUMENT)
        nproc = req.num_processors or req.capacity
        prio = req.servant_priority
        if prio == PRIORITY_UNKNOWN or prio not in (PRIORITY_USER, PRIORITY_DEDICATED):
            prio = PRIORITY_USER
        max_tasks, reason = req.capacity, req.not_accepting_task_reason
        if observed != reported:
            max_tasks, reason = 0, REASON_BEHIND_NAT
        if req.token not in self.servant:
            max_tasks, reason = 0, REASON_NOT_VERIFIED
        if req.next_heartbeat_in_ms == 0:
            max_tasks = 0
        self.d.keep_servant_alive(
            Servant(observed, reported, list(req.env_digests), req.version, nproc, req.current_load, req.total_memory_in_bytes,
                    req.memory_available_in_bytes, max_tasks, prio, reason), req.next_heartbeat_in_ms / 1000.0, now=now)
        rolls = self._active_tokens(now)
        expired = self.d.notify_servant_running_tasks(req.location, list(req.running_tasks))  # (the REPORTED location, :181)
        return ModelHeartbeatResponse(OK, rolls, [int(x) for x in expired])

    def get_config(self, token: str, *, now: float):  # :196-208 -> (status, rolls)
        if token not in self.user:
            return ACCESS_DENIED, None
        return OK, self._active_tokens(now)

    def wait_for_starting_task(self, token: str, rpc, *, now: float):
        """:209-271 for ONE RPC (a row of RPC_WAIT_DTYPE): (status, [(task_id, servant_index)])."""
        if token not in self.user:
            return ACCESS_DENIED, []
        if int(rpc["milliseconds_to_wait"]) > 10_000 or int(rpc["next_keep_alive_ns"]) > 30_000_000_000:
            return INVALID_ARGUMENT, []
        r = np.zeros(1, dtype=_abi.REQ_DTYPE)
        r["env_id"], r["min_version"], r["requestor_ip"] = rpc["env_id"], rpc["min_version"], rpc["requestor_ip"]
        r["expires_in_ns"] = rpc["next_keep_alive_ns"]
        grants = []
        for _ in range(int(rpc["immediate_reqs"])):
            r["flags"] = 0
            g = self.d.wait_for_starting_new_tasks(r.copy(), now)[0]
            if g["status"] != STATUS_GRANTED:
                if g["status"] == STATUS_ENVIRONMENT_NOT_FOUND:
                    return ENV_NOT_AVAILABLE, []  # :242-246 (grants collected so far are dropped with the failed RPC)
                break
            grants.append((int(g["task_id"]), int(g["servant_index"])))
        for _ in range(int(rpc["prefetch_reqs"])):
            r["flags"] = _abi.REQ_FLAG_PREFETCH
            g = self.d.wait_for_starting_new_tasks(r.copy(), now)[0]
            if g["status"] != STATUS_GRANTED:
                break
            grants.append((int(g["task_id"]), int(g["servant_index"])))
        if not grants:
            return NO_QUOTA, []
        return OK, grants

    def keep_task_alive(self, token: str, ids, next_keep_alive_in_ms: int, *, now: float):  # :272-292
        if token not in self.user:
            return ACCESS_DENIED, []
        if next_keep_alive_in_ms > 30_000:
            return INVALID_ARGUMENT, []
        return OK, [bool(self.d.keep_task_alive(int(i), next_keep_alive_in_ms / 1000.0, now=now)) for i in ids]

    def free_task(self, token: str, ids) -> int:  # :294-308
        if token not in self.user:
            return ACCESS_DENIED
        for i in ids:
            self.d.free_task(int(i))
        return OK

    def get_running_tasks(self):  # :310-317
        return self.d.get_running_tasks()
