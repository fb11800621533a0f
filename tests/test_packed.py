"""The packed interface (yd_wait_for_starting_new_tasks_packed: 16-byte requests, 8-byte grants)
makes the same decisions as the plain call.  CPU part: the checkers' generic implementation and
the Python pack / unpack helpers; the CUDA backend's native implementation is in test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import REF_LIB
from yadcc_b200 import streams as S
from yadcc_b200 import _abi, pack_requests, unpack_grants, STATUS_GRANTED


def test_pack_unpack_round_trip():
    r = np.zeros(5, dtype=_abi.REQ_DTYPE)
    r["env_id"] = [1, 2, 3, 4, 5]
    r["min_version"] = [0, 7, 8, 9, 0xFFFFFFFF]
    r["requestor_ip"] = [0, 1, 2, 3, 4]
    r["flags"] = [0, 1, 0, 1, 0]
    r["expires_in_ns"] = [0, 1_000_000, 500_000_000, 15_000_000_000, 30_000_000_000]
    p = pack_requests(r)
    assert p.dtype.itemsize == 16
    assert (p["lease"] & 0x7FFFFFFF).tolist() == [0, 1, 500, 15000, 30000]
    assert ((p["lease"] >> 31) & 1).tolist() == [0, 1, 0, 1, 0]
    g8 = np.zeros(3, dtype=_abi.GRANT8_DTYPE)
    g8["servant_index"] = [7, 0xFFFFFFFF, 9]
    g8["status_ordinal"] = [(2 << 30) | 0, (1 << 30), (2 << 30) | 1]
    ids = np.zeros(1, dtype=_abi.PACKED_IDS_DTYPE)
    ids["first_task_id"], ids["stride"] = 100, 3
    g = unpack_grants(g8, ids[0])
    assert g["task_id"].tolist() == [100, 0, 103] and g["status"].tolist() == [2, 1, 2]
    with pytest.raises(AssertionError):
        r["expires_in_ns"][0] = 1  # not a whole millisecond
        pack_requests(r)


@pytest.mark.parametrize("kind", ["port", "ref"])
@pytest.mark.parametrize("name", ["cfg1", "cfg2-random-small", "cfg-self-small", "fuzz-3", "fuzz-10", "fuzz-21"])
def test_packed_equals_plain_on_checkers(make_dispatcher, kind, name):
    traces = []
    for packed in (False, True):
        d = make_dispatcher(kind)
        traces.append(S.Replayer(d, packed=packed).run(S.named_stream(name, d)))
        d.close()
    assert S.traces_equal(*traces), S.first_mismatch(*traces)
    assert any((t["status"] == STATUS_GRANTED).any() for t in traces[1] if t.dtype == _abi.GRANT_DTYPE)


def test_prefiltered_solve_is_the_three_calls_in_order(make_dispatcher):
    """The checkers' yd_filter_and_wait_for_starting_new_tasks against the calls it is defined by."""
    from bloom_cases import tu_keys
    from running_index_cases import task_digests
    from yadcc_b200 import RunningTask

    keys, digests = tu_keys(400), task_digests(400, 11)
    n = 2000
    tu = np.arange(n) % len(keys)
    trace, trace_digests = [keys[i] for i in tu], [digests[i] for i in tu]
    out = []
    for fused in (True, False):
        d = make_dispatcher("port")
        w = S.config2(n, 64, 4, variant="mod", max_tasks=16)
        w.register(d)
        d.bloom_reset()
        d.bloom_add(keys[::3])
        reqs = w.build_requests(d)
        early = d.wait_for_starting_new_tasks(reqs[:100].copy(), 0.25)
        locs = [d.servant_location(i) for i in range(64)]
        for j, gr in enumerate(early):
            if gr["status"] == STATUS_GRANTED:
                d.notify_servant_running_tasks(locs[int(gr["servant_index"])],
                                               [RunningTask(j + 1, int(gr["task_id"]), locs[int(gr["servant_index"])], digests[399 - j])])
        d.running_index_refresh()
        if fused:
            verdict, hits, g = d.filter_and_wait_for_starting_new_tasks(reqs, trace, trace_digests, 0.5)
        else:
            hit = d.bloom_possibly_contains(trace)
            hits = d.find_running_tasks(trace_digests)
            verdict = np.where(hit, 1, np.where(hits["found"] != 0, 2, 0)).astype(np.uint8)
            g = d.wait_for_starting_new_tasks(reqs[verdict == 0], 0.5)
        out.append((verdict.copy(), hits.copy(), g.copy()))
    for a, b in zip(*out):
        assert a.shape == b.shape and (a == b).all()
