"""The reference's own unit tests for the hot path, re-expressed against the
C ABI with virtual time instead of sleeps (gtest / the fiber runtime are not
buildable here).  Each function cites the test it restates; expectations are
the reference's, verbatim.

Used three ways: against the CPU restatement (oracle/port.cc), against the
reference compiled verbatim (oracle/_ref) and -- on the GPU box -- against the
CUDA backend.
"""
from yadcc_b200 import RunningTask, Servant, TaskAllocation, WaitStatus
from yadcc_b200 import PRIORITY_DEDICATED, PRIORITY_USER

G50 = 50 * 1024**3


def case_all(d):
    """yadcc/scheduler/task_dispatcher_test.cc:29-144 `TaskDispatcher.All`."""
    servant = Servant(
        observed_location="127.0.0.1:1234",
        reported_location="127.0.0.1:1234",
        environments=["digest"],
        max_tasks=10,
        current_load=0,
        num_processors=10,
        priority=PRIORITY_USER,
        version=8,
        memory_available_in_bytes=G50,
    )
    t = 0.0
    d.keep_servant_alive(servant, 10.0, now=t)

    # :44-56 no environment available
    r = d.wait_for_starting_new_task("127.0.0.1", 8, "not found", 1.0, now=t)
    assert r == WaitStatus.EnvironmentNotFound

    # :58-72 allocate 10 tasks
    tasks = []
    for _ in range(10):
        a = d.wait_for_starting_new_task("127.0.0.1", 8, "digest", 5.0, now=t)
        assert isinstance(a, TaskAllocation)
        assert a.servant_location == "127.0.0.1:1234"
        tasks.append(a)
    assert [a.task_id for a in tasks] == list(range(10))  # next_task_id starts at 0 (h:218)

    # :74-89 no servant available -> Timeout (the reference waits ~1 s first)
    r = d.wait_for_starting_new_task("127.0.0.1", 8, "digest", 1.0, now=t)
    assert r == WaitStatus.Timeout
    t += 1.0

    # :91-92 unrecognised task id
    assert d.keep_task_alive(12345678, 1.0, now=t) is False
    # :94-97 keep existing tasks alive for 1 s
    for a in tasks:
        assert d.keep_task_alive(a.task_id, 1.0, now=t) is True

    # :99-107 1000002, 1000003 are unknown and must be returned
    unknown = d.notify_servant_running_tasks(
        "127.0.0.1:1234",
        [RunningTask(task_grant_id=tasks[0].task_id), RunningTask(task_grant_id=1000002), RunningTask(task_grant_id=1000003)],
    )
    assert unknown == [1000002, 1000003]

    # :109 sleep 2 s; the 1 Hz timer fires meanwhile
    d.on_expiration_timer(now=t + 1.0)
    d.on_expiration_timer(now=t + 2.0)
    t += 2.0
    # :112-114 all tasks expired by now
    for a in tasks:
        assert d.keep_task_alive(a.task_id, 1.0, now=t) is False

    # :116-127 every (zombie) id the servant still reports comes back as unknown
    ids = [a.task_id for a in tasks]
    assert d.notify_servant_running_tasks("127.0.0.1:1234", [RunningTask(task_grant_id=i) for i in ids]) == ids

    # :129-132 renew the servant for 1 s, sleep 2 s -> it expires
    d.keep_servant_alive(servant, 1.0, now=t)
    d.on_expiration_timer(now=t + 1.0)
    d.on_expiration_timer(now=t + 2.0)
    t += 2.0
    # :134-143 nothing can be allocated any more
    r = d.wait_for_starting_new_task("127.0.0.1", 8, "digest", 1.0, now=t)
    assert not isinstance(r, TaskAllocation)
    assert d.num_servants() == 0


def case_prefer_dedicated(d):
    """task_dispatcher_test.cc:146-186 `PreferDedicated`."""
    servant = Servant("127.0.0.1:1234", "127.0.0.1:1234", ["digest"], 8, 10, 0, 0, G50, 10, PRIORITY_USER)
    d.keep_servant_alive(servant, 1.0, now=0.0)
    a = d.wait_for_starting_new_task("127.0.0.1", 8, "digest", 1.0, now=0.0)
    assert isinstance(a, TaskAllocation) and a.servant_location == "127.0.0.1:1234"
    d.free_task(a.task_id)

    dedicated = Servant("192.168.0.1:1234", "192.168.0.1:1234", ["digest"], 8, 10, 2, 0, G50, 10, PRIORITY_DEDICATED)
    d.keep_servant_alive(dedicated, 1.0, now=0.0)
    a = d.wait_for_starting_new_task("127.0.0.1", 8, "digest", 1.0, now=0.0)
    # prefer the dedicated servant even though its load is higher
    assert isinstance(a, TaskAllocation) and a.servant_location == "192.168.0.1:1234"
    d.free_task(a.task_id)
    d.on_expiration_timer(now=1.5)
    assert d.num_servants() == 0


def case_load_balance(d):
    """task_dispatcher_test.cc:188-298 `LoadBalanceCase`: the known-answer test
    for the capacity / utilisation arithmetic."""

    def add(loc, max_tasks, nproc, load):
        s = Servant(loc, loc, ["Load Balance"], 8, nproc, load, 0, G50, max_tasks, PRIORITY_USER)
        d.keep_servant_alive(s, 10.0, now=0.0)
        return s

    def pick():
        return d.wait_for_starting_new_task("127.0.0.3", 8, "Load Balance", 1.0, now=0.0)

    add("192.168.0.0:0000", 7, 16, 16)  # overloaded: never picked (:220-230)
    assert not isinstance(pick(), TaskAllocation)

    s1 = add("192.168.0.1:1111", 7, 16, 1)
    s2 = add("192.168.0.2:2222", 8, 16, 5)
    s3 = add("192.168.0.3:3333", 6, 16, 12)
    # :239-291 the expected pick sequence with the fractions in the comments
    for expect in (s1, s2, s3, s2, s1, s2, s3):
        a = pick()
        assert isinstance(a, TaskAllocation)
        assert a.servant_location == expect.observed_location
        expect.current_load += 1  # "our task is running now"
        d.keep_servant_alive(expect, 10.0, now=0.0)
    st = d.servant_state()
    assert st["running_tasks"].tolist() == [0, 2, 3, 2]


def case_running_task_bookkeeper(d):
    """yadcc/scheduler/running_task_bookkeeper_test.cc:24-42, driven through
    NotifyServantRunningTasks / GetRunningTasks / servant expiry (the
    bookkeeper is private to the dispatcher)."""
    s = Servant("10.1.1.1:8335", None, ["d"], 8, 8, 0, 0, G50, 8, PRIORITY_USER)
    d.keep_servant_alive(s, 1.0, now=0.0)
    grants = [d.wait_for_starting_new_task("10.9.9.9", 0, "d", 10.0, now=0.0) for _ in range(3)]
    tasks = [RunningTask(i + 100, g.task_id, "10.1.1.1:8335", f"{i:064x}") for i, g in enumerate(grants)]
    assert d.notify_servant_running_tasks("10.1.1.1:8335", tasks) == []
    got = d.get_running_tasks()
    assert [t.servant_task_id for t in got] == [100, 101, 102]
    assert [t.task_digest for t in got] == [f"{i:064x}" for i in range(3)]
    # DropServant happens when the servant expires (task_dispatcher.cc:510-511)
    d.on_expiration_timer(now=2.0)
    assert d.get_running_tasks() == []


def case_parse_size(d):
    """yadcc/common/parse_size_test.cc:23-29."""
    assert d.parse_size("123") == 123
    assert d.parse_size("2K") == 2048
    assert d.parse_size("3M") == 3145728
    assert d.parse_size("1G") == 1073741824
    assert d.parse_size("3A") is None
    assert d.parse_size("10G") == 10737418240  # the flag default, task_dispatcher.cc:35


def case_token_gating(d):
    """scheduler_service_impl_test.cc:80-172: a heartbeat that fails servant-token
    verification (or is behind NAT / leaving) reaches the dispatcher with
    max_tasks = 0 (scheduler_service_impl.cc:146-157,168-170) and must never be
    granted, yet still makes the environment 'recognised'... no: max_tasks == 0
    servants are NOT eligible (task_dispatcher.cc:329-331), so the request fails
    with EnvironmentNotFound."""
    s = Servant("10.2.2.2:8335", None, ["d"], 8, 16, 0, 0, G50, 0, PRIORITY_USER, not_accepting_task_reason=100)
    d.keep_servant_alive(s, 10.0, now=0.0)
    assert d.wait_for_starting_new_task("10.9.9.9", 0, "d", 1.0, now=0.0) == WaitStatus.EnvironmentNotFound
    s.max_tasks = 4
    d.keep_servant_alive(s, 10.0, now=0.0)
    assert isinstance(d.wait_for_starting_new_task("10.9.9.9", 0, "d", 1.0, now=0.0), TaskAllocation)


ALL_CASES = [
    case_all,
    case_prefer_dedicated,
    case_load_balance,
    case_running_task_bookkeeper,
    case_parse_size,
    case_token_gating,
]
