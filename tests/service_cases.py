"""Shared cases for the SchedulerServiceImpl restatement (include/ydservice.h).
The first three restate yadcc/scheduler/scheduler_service_impl_test.cc:38-172 with virtual
time; the rest walk the handler rules line by line (scheduler_service_impl.cc:67-333)."""
import numpy as np

from yadcc_b200 import RunningTask, _abi
from yadcc_b200.service import (REASON_BEHIND_NAT, REASON_NOT_VERIFIED, STATUS_ACCESS_DENIED,
                                STATUS_ENVIRONMENT_NOT_AVAILABLE, STATUS_INVALID_ARGUMENT,
                                STATUS_NO_QUOTA_AVAILABLE, STATUS_OK, STATUS_VERSION_TOO_OLD, HeartbeatRequest,
                                SchedulerService)

GIB = 1 << 30
DIGEST = "c" * 64


def token_case(d):
    """TEST(SchedulerServiceImpl, Token), :38-78 (roll-out interval 1 s, `sleep 2 s`)."""
    svc = SchedulerService(d, acceptable_user_tokens="token1,token2", acceptable_servant_tokens="token1,token2",
                           serving_daemon_token_rollout_interval=1, token_seed=7, now=0.0)
    assert svc.get_config("", now=0.0) == (STATUS_ACCESS_DENIED, None)
    st, first = svc.get_config("token1", now=0.0)
    assert st == STATUS_OK
    st, second = svc.get_config("token1", now=0.1)
    assert st == STATUS_OK and first == second
    st, third = svc.get_config("token1", now=2.1)
    assert st == STATUS_OK and third != first
    assert len(first) == 32 and all(c in "0123456789abcdef" for c in first)  # EncodeHex(16 bytes), :46-51
    return first, third


def _reported_locations(d):
    return {d.servant_personality(i).reported_location for i in range(d.num_servants())}


def token_with_intersection_case(d):
    """TEST(SchedulerServiceImpl, TokenWithIntersection), :80-129."""
    svc = SchedulerService(d, acceptable_user_tokens="token1,token2", acceptable_servant_tokens="token2,token3",
                           token_seed=1)
    req = HeartbeatRequest(servant_priority=2, remote_ip="192.0.2.1")
    for tok, loc, want in (("token1", "192.0.2.128:6666", STATUS_OK), ("token2", "192.0.2.128:7777", STATUS_OK),
                           ("token3", "192.0.2.128:8888", STATUS_OK),
                           ("token4", "192.0.2.128:9999", STATUS_ACCESS_DENIED)):
        req.token, req.location = tok, loc
        assert svc.heartbeat(req).status == want
    locs = _reported_locations(d)
    assert {"192.0.2.128:6666", "192.0.2.128:7777", "192.0.2.128:8888"} <= locs
    assert "192.0.2.128:9999" not in locs
    return sorted(locs)


def token_without_intersection_case(d):
    """TEST(SchedulerServiceImpl, TokenWithoutIntersection), :131-172."""
    svc = SchedulerService(d, acceptable_user_tokens="token1", acceptable_servant_tokens="token2", token_seed=1)
    req = HeartbeatRequest(servant_priority=2, remote_ip="192.0.2.1")
    for tok, loc, want in (("token1", "192.0.2.128:6666", STATUS_OK), ("token2", "192.0.2.128:7777", STATUS_OK),
                           ("token3", "192.0.2.128:8888", STATUS_ACCESS_DENIED)):
        req.token, req.location = tok, loc
        assert svc.heartbeat(req).status == want
    locs = _reported_locations(d)
    assert {"192.0.2.128:6666", "192.0.2.128:7777"} <= locs and "192.0.2.128:8888" not in locs
    return sorted(locs)


def _hb(**kw):
    base = dict(token="srv", location="10.0.0.5:8335", remote_ip="10.0.0.5", next_heartbeat_in_ms=5000, version=9,
                num_processors=32, current_load=1, servant_priority=2, capacity=12,
                total_memory_in_bytes=64 * GIB, memory_available_in_bytes=50 * GIB, env_digests=[DIGEST])
    base.update(kw)
    return HeartbeatRequest(**base)


def heartbeat_rules_case(d):
    """Every branch of Heartbeat (:67-194); returns what ended up in the registry."""
    svc = SchedulerService(d, acceptable_user_tokens="usr,both", acceptable_servant_tokens="srv,both",
                           min_daemon_version=5, token_seed=3)
    out = []
    # rejected before anything is registered
    assert svc.heartbeat(_hb(token="nobody")).status == STATUS_ACCESS_DENIED            # :73-77
    assert svc.heartbeat(_hb(version=4)).status == STATUS_VERSION_TOO_OLD                # :78-81
    for bad in ("", "10.0.0.5", "10.0.0.5:", "10.0.0.5:x", "10.0.0.5:70000", "10.0.0.256:80", "host:80", "[::1]"):
        assert svc.heartbeat(_hb(location=bad)).status == STATUS_INVALID_ARGUMENT, bad   # :87-93
    assert svc.heartbeat(_hb(next_heartbeat_in_ms=30001)).status == STATUS_INVALID_ARGUMENT  # :119-123
    assert d.num_servants() == 0
    # a plain, verified, directly reachable servant
    r = svc.heartbeat(_hb(), now=1.0)
    assert r.status == STATUS_OK and len(r.acceptable_tokens) == 3 and r.expired_tasks == []
    p = d.servant_personality(0)
    assert (p.observed_location, p.reported_location) == ("10.0.0.5:8335", "10.0.0.5:8335")
    assert (p.max_tasks, p.num_processors, p.current_load, p.version, p.priority, p.not_accepting_task_reason) == \
        (12, 32, 1, 9, 2, 0)
    assert list(p.environments) == [DIGEST]
    out.append(d.servant_state()[0]["expires_at_ns"])                                    # now + 5 s
    # exactly 30 s is allowed
    assert svc.heartbeat(_hb(next_heartbeat_in_ms=30000), now=1.0).status == STATUS_OK
    # behind NAT: observed (peer ip + reported port) != reported (:147-154)
    assert svc.heartbeat(_hb(location="192.168.1.9:7000", remote_ip="10.0.0.6", not_accepting_task_reason=1)).status == STATUS_OK
    p = d.servant_personality(1)
    assert (p.observed_location, p.reported_location) == ("10.0.0.6:7000", "192.168.1.9:7000")
    assert (p.max_tasks, p.not_accepting_task_reason) == (0, REASON_BEHIND_NAT)
    # a user token may report but not serve (:155-158); it also overrides the NAT reason
    assert svc.heartbeat(_hb(token="usr", location="10.0.0.7:8335", remote_ip="10.0.0.7")).status == STATUS_OK
    p = d.servant_personality(2)
    assert (p.max_tasks, p.not_accepting_task_reason) == (0, REASON_NOT_VERIFIED)
    assert svc.heartbeat(_hb(token="usr", location="192.168.1.9:7001", remote_ip="10.0.0.8")).status == STATUS_OK
    assert d.servant_personality(3).not_accepting_task_reason == REASON_NOT_VERIFIED
    # a token on both lists serves
    assert svc.heartbeat(_hb(token="both", location="10.0.0.9:8335", remote_ip="10.0.0.9")).status == STATUS_OK
    assert d.servant_personality(4).max_tasks == 12
    # older daemons: no processor count -> capacity; unknown / invalid priority -> USER (:131-143)
    assert svc.heartbeat(_hb(location="10.0.0.10:8335", remote_ip="10.0.0.10", num_processors=0, capacity=7,
                             servant_priority=0)).status == STATUS_OK
    p = d.servant_personality(5)
    assert (p.num_processors, p.max_tasks, p.priority) == (7, 7, 2)
    assert svc.heartbeat(_hb(location="10.0.0.11:8335", remote_ip="10.0.0.11", servant_priority=9)).status == STATUS_OK
    assert d.servant_personality(6).priority == 2
    assert svc.heartbeat(_hb(location="10.0.0.12:8335", remote_ip="10.0.0.12", servant_priority=1)).status == STATUS_OK
    assert d.servant_personality(7).priority == 1
    # leaving: next_heartbeat_in_ms == 0 -> capacity 0, reason as reported (:163-172)
    assert svc.heartbeat(_hb(location="10.0.0.13:8335", remote_ip="10.0.0.13", next_heartbeat_in_ms=0,
                             not_accepting_task_reason=2), now=3.0).status == STATUS_OK
    p = d.servant_personality(8)
    assert (p.max_tasks, p.not_accepting_task_reason) == (0, 2)
    # the reported IP is normalised, the observed one is taken as given (:95-107)
    assert svc.heartbeat(_hb(location="[2001:db8:0:0::1]:9000", remote_ip="2001:db8::1", remote_is_ipv6=True)).status == STATUS_OK
    p = d.servant_personality(9)
    assert (p.observed_location, p.reported_location) == ("[2001:db8::1]:9000", "[2001:db8::1]:9000") and p.max_tasks == 12
    for i in range(d.num_servants()):
        p = d.servant_personality(i)
        out.append((p.observed_location, p.reported_location, p.max_tasks, p.num_processors, p.priority,
                    p.not_accepting_task_reason, p.version))
    return out


def lease_flow_case(d):
    """WaitForStartingTask / KeepTaskAlive / FreeTask / GetRunningTasks through the service,
    including the expired-task answer of Heartbeat and its reported-location quirk (:182-186)."""
    svc = SchedulerService(d, acceptable_user_tokens="usr", acceptable_servant_tokens="srv", token_seed=5)
    out = []
    assert svc.heartbeat(_hb(), now=0.0).status == STATUS_OK                                     # 10.0.0.5, 12 tasks
    assert svc.heartbeat(_hb(location="192.168.1.9:7000", remote_ip="10.0.0.6"), now=0.0).status == STATUS_OK  # NAT-ed
    rpcs = np.zeros(5, dtype=_abi.RPC_WAIT_DTYPE)
    rpcs["env_id"] = d.intern_env(DIGEST)
    rpcs["requestor_ip"] = d.intern_ip("10.9.9.9")
    rpcs["immediate_reqs"] = [1, 2, 1, 1, 1]
    rpcs["prefetch_reqs"] = [0, 1, 0, 0, 30]
    rpcs["milliseconds_to_wait"] = [0, 100, 10001, 0, 0]
    rpcs["next_keep_alive_ns"] = 10_000_000_000
    rpcs["env_id"][3] = d.intern_env("e" * 64)  # nobody has it
    res, grants = svc.wait_for_starting_tasks(["usr", "usr", "usr", "usr", "bad"], rpcs, now=1.0)
    assert list(res["status"]) == [STATUS_OK, STATUS_OK, STATUS_INVALID_ARGUMENT, STATUS_ENVIRONMENT_NOT_AVAILABLE,
                                   STATUS_ACCESS_DENIED]
    assert list(res["n_grants"]) == [1, 3, 0, 0, 0] and len(grants) == 4
    out += [res.copy(), grants.copy()]
    ids = grants["task_id"]
    # the cloud is full for a 13-task request minus what is taken: NO_QUOTA only if nothing was granted
    rp2 = rpcs[:1].copy()
    rp2["immediate_reqs"], rp2["prefetch_reqs"] = 20, 0
    res2, g2 = svc.wait_for_starting_tasks(["usr"], rp2, now=1.0)
    assert res2["status"][0] == STATUS_OK and res2["n_grants"][0] == 8
    res3, g3 = svc.wait_for_starting_tasks(["usr"], rp2, now=1.0)
    assert res3["status"][0] == STATUS_NO_QUOTA_AVAILABLE and len(g3) == 0
    out += [res2.copy(), g2.copy(), res3.copy()]
    # KeepTaskAlive: token, 30 s limit, per-id statuses (:272-292)
    assert svc.keep_task_alive("bad", ids, 1000, now=2.0)[0] == STATUS_ACCESS_DENIED
    assert svc.keep_task_alive("usr", ids, 30001, now=2.0)[0] == STATUS_INVALID_ARGUMENT
    st, ok = svc.keep_task_alive("usr", list(ids) + [999999], 30000, now=2.0)
    assert st == STATUS_OK and list(ok) == [True] * 4 + [False]
    # heartbeat with running tasks: known ones are listed by GetRunningTasks, unknown ones come back as expired
    mine = [RunningTask(100 + i, int(t), "10.0.0.5:8335", "%064x" % i) for i, t in enumerate(ids[:3])]
    r = svc.heartbeat(_hb(running_tasks=mine + [RunningTask(7, 424242, "10.0.0.5:8335", "f" * 64)]), now=3.0)
    assert r.status == STATUS_OK and r.expired_tasks == [424242]
    rt = svc.get_running_tasks()
    assert sorted(t.task_grant_id for t in rt) == sorted(int(t) for t in ids[:3])
    out.append(np.asarray(sorted((t.servant_task_id, t.task_grant_id) for t in rt), dtype=np.uint64))
    # the NAT-ed servant is registered under its OBSERVED location but looked up by the REPORTED one:
    # not found -> everything it lists is returned as expired (:182-183, task_dispatcher.cc:241-243)
    r = svc.heartbeat(_hb(location="192.168.1.9:7000", remote_ip="10.0.0.6",
                          running_tasks=[RunningTask(1, int(ids[0]), "x", "a" * 64)]), now=3.0)
    assert r.status == STATUS_OK and r.expired_tasks == [int(ids[0])]
    # FreeTask (:294-308)
    assert svc.free_task("bad", ids[:2]) == STATUS_ACCESS_DENIED
    assert svc.free_task("usr", ids[:2]) == STATUS_OK
    st, ok = svc.keep_task_alive("usr", ids, 1000, now=4.0)
    assert list(ok) == [False, False, True, True]
    out.append(d.servant_state().copy())
    return out


def token_rollout_case(d):
    """DetermineActiveServingDaemonTokens (:319-333): three tokens, the window slides by one
    when `next_rollout < now`, at most once per call."""
    svc = SchedulerService(d, acceptable_user_tokens="u", acceptable_servant_tokens="s",
                           serving_daemon_token_rollout_interval=10, token_seed=11, now=100.0)
    a = svc.heartbeat(_hb(token="s"), now=100.0).acceptable_tokens
    assert len(set(a)) == 3 and svc.get_config("u", now=105.0)[1] == a[1]
    assert svc.heartbeat(_hb(token="s"), now=110.0).acceptable_tokens == a       # not yet: strict '<'
    b = svc.heartbeat(_hb(token="s"), now=110.5).acceptable_tokens
    assert b[:2] == a[1:] and b[2] not in a
    # a long silence still slides by ONE per call
    c = svc.heartbeat(_hb(token="s"), now=500.0).acceptable_tokens
    assert c[:2] == b[1:]
    assert svc.get_config("u", now=500.0)[1] == c[1]
    return a, b, c
