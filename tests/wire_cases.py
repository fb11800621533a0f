"""Shared scenario for the FlareStd wire front end (include/ydwire.h): a small cluster driven
ONLY through request frames encoded by the google.protobuf runtime; every response frame is
decoded by that runtime, compared with what the handlers return when called directly on a twin
scheduler, and re-encoded to check the hand-written encoder byte for byte."""
import struct

import numpy as np

from yadcc_b200 import RunningTask, _abi
from yadcc_b200.service import HeartbeatRequest, SchedulerService

import wire_protos as W

PB = W.PB
DIGEST = "a1" * 32
GIB = 1 << 30


def _services(make_dispatcher, kind):
    mk = lambda: SchedulerService(make_dispatcher(kind), acceptable_user_tokens="usr", acceptable_servant_tokens="srv",
                                  token_seed=9, now=0.0)
    return mk(), mk()  # one driven through frames, its twin through direct handler calls


def _hb_msg(i, token="srv", running=()):
    m = PB["HeartbeatRequest"](token=token, next_heartbeat_in_ms=5000, version=9, location=f"10.0.1.{i}:8335",
                               num_processors=16, current_load=1, servant_priority=2, capacity=6,
                               total_memory_in_bytes=64 * GIB, memory_available_in_bytes=50 * GIB)
    m.env_descs.add().compiler_digest = DIGEST
    for t in running:
        r = m.running_tasks.add()
        r.servant_task_id, r.task_grant_id, r.servant_location, r.task_digest = t
    return m


def _hb_direct(i, token="srv", running=()):
    return HeartbeatRequest(token=token, location=f"10.0.1.{i}:8335", remote_ip=f"10.0.1.{i}", next_heartbeat_in_ms=5000,
                            version=9, num_processors=16, current_load=1, servant_priority=2, capacity=6,
                            total_memory_in_bytes=64 * GIB, memory_available_in_bytes=50 * GIB, env_digests=[DIGEST],
                            running_tasks=[RunningTask(*t) for t in running])


def _check_frame(frame, resp_cls, corr, status, description=None):
    meta, body = W.parse_response_frame(frame, resp_cls)
    assert meta.correlation_id == corr and meta.method_type == 1 and meta.HasField("response_meta")
    assert not meta.HasField("request_meta") and not meta.HasField("flags")
    assert meta.response_meta.status == status
    assert meta.response_meta.HasField("description") == (status != 0)
    if description is not None:
        assert meta.response_meta.description == description
    # byte-exact: the runtime's own serialisation of what it just parsed
    ms, bs = struct.unpack("<II", frame[4:12])
    assert frame[16:16 + ms] == meta.SerializeToString()
    assert frame[16 + ms:] == body.SerializeToString()
    return body


def run_wire_scenario(make_dispatcher, kind):
    wire, twin = _services(make_dispatcher, kind)
    trace = []
    corr = 1000

    def one(method, msg, ip, resp_cls, status=0, description=None, now=0.0):
        nonlocal corr
        corr += 1
        (verdict, consumed, st, frame), = wire.handle_frames([(W.request_frame(method, msg, corr), ip)], now=now)
        assert verdict == 1 and st == status and consumed == len(W.request_frame(method, msg, corr))
        body = _check_frame(frame, resp_cls, corr, status, description)
        trace.append(frame[16:])
        return body

    # -- heartbeats: 8 servants, one behind NAT, one reporting with a user token -----------------
    for i in range(8):
        token = "usr" if i == 7 else "srv"
        ip = "10.9.9.9" if i == 6 else f"10.0.1.{i}"
        body = one("Heartbeat", _hb_msg(i, token), ip, PB["HeartbeatResponse"])
        d = _hb_direct(i, token)
        d.remote_ip = ip
        r = twin.heartbeat(d)
        assert list(body.acceptable_tokens) == r.acceptable_tokens and list(body.expired_tasks) == r.expired_tasks
    assert one("Heartbeat", _hb_msg(9, "nobody"), "10.0.1.9", PB["HeartbeatResponse"], 1003, "").ByteSize() == 0
    bad = _hb_msg(9)
    bad.location = "not-an-endpoint"
    one("Heartbeat", bad, "10.0.1.9", PB["HeartbeatResponse"], 1004, "")
    for i in range(wire.dispatcher.num_servants()):
        assert wire.dispatcher.servant_personality(i) == twin.dispatcher.servant_personality(i)

    # -- GetConfig ------------------------------------------------------------------------------
    assert one("GetConfig", PB["GetConfigRequest"](token="usr"), "10.2.0.1", PB["GetConfigResponse"]).serving_daemon_token == \
        twin.get_config("usr")[1]
    one("GetConfig", PB["GetConfigRequest"](token="srv"), "10.2.0.1", PB["GetConfigResponse"], 1003, "")

    # -- WaitForStartingTask: five frames in ONE call = one batched solve ---------------------------
    def wait_msg(token="usr", imm=1, pre=0, ms=0, ka=10000, digest=DIGEST, mv=0):
        m = PB["WaitForStartingTaskRequest"](token=token, milliseconds_to_wait=ms, immediate_reqs=imm, prefetch_reqs=pre,
                                             next_keep_alive_in_ms=ka, min_version=mv)
        m.env_desc.compiler_digest = digest
        return m

    msgs = [wait_msg(imm=2, pre=1), wait_msg(token="bad"), wait_msg(ka=30001), wait_msg(digest="ff" * 32), wait_msg(imm=4),
            wait_msg(imm=0, pre=2, mv=10)]
    ips = ["10.2.0.1", "10.2.0.2", "10.2.0.3", "10.2.0.4", "10.0.1.2", "10.2.0.5"]  # the fifth requestor is servant 2 itself
    frames = []
    for m, ip in zip(msgs, ips):
        corr += 1
        frames.append((W.request_frame("WaitForStartingTask", m, corr), ip))
    outs = wire.handle_frames(frames, now=1.0)
    # the twin: the same RPCs through the service's batched call
    d = twin.dispatcher
    rpcs = np.zeros(len(msgs), dtype=_abi.RPC_WAIT_DTYPE)
    for k, (m, ip) in enumerate(zip(msgs, ips)):
        rpcs[k] = (d.intern_env(m.env_desc.compiler_digest), m.min_version, d.intern_ip(ip), m.immediate_reqs, m.prefetch_reqs,
                   m.milliseconds_to_wait, m.next_keep_alive_in_ms * 1_000_000)
    res, grants = twin.wait_for_starting_tasks([m.token for m in msgs], rpcs, now=1.0)
    want_desc = {0: None, 1001: "The compilation cloud is busy now.", 1003: "", 1004: "",
                 1006: "No matched servant environment."}
    all_ids = []
    for k, (verdict, consumed, st, frame) in enumerate(outs):
        assert verdict == 1 and st == int(res["status"][k]), (k, st, res["status"][k])
        body = _check_frame(frame, PB["WaitForStartingTaskResponse"], corr - len(msgs) + 1 + k, st, want_desc[st])
        trace.append(frame[16:])
        g = grants[res["first_grant"][k]: res["first_grant"][k] + res["n_grants"][k]] if st == 0 else grants[:0]
        assert [x.task_grant_id for x in body.grants] == [int(t) for t in g["task_id"]]
        assert [x.servant_location for x in body.grants] == [d.servant_location(int(s)) for s in g["servant_index"]]
        all_ids += [x.task_grant_id for x in body.grants]
    assert list(res["status"]) == [0, 1003, 1004, 1006, 0, 1001] and len(all_ids) == 7 and all_ids[0] == 0

    # -- KeepTaskAlive / FreeTask / Heartbeat with running tasks / GetRunningTasks ------------------
    ka = PB["KeepTaskAliveRequest"](token="usr", next_keep_alive_in_ms=8000)
    ka.task_grant_ids.extend(all_ids + [424242])
    body = one("KeepTaskAlive", ka, "10.2.0.1", PB["KeepTaskAliveResponse"], now=2.0)
    assert list(body.statuses) == list(twin.keep_task_alive("usr", all_ids + [424242], 8000, now=2.0)[1])
    ka.next_keep_alive_in_ms = 30001
    one("KeepTaskAlive", ka, "10.2.0.1", PB["KeepTaskAliveResponse"], 1004, "", now=2.0)
    # a servant reports two of its tasks (one unknown): the unknown id comes back as expired
    st_w = wire.dispatcher.servant_state()["running_tasks"]
    sv = int(np.argmax(st_w))
    loc = wire.dispatcher.servant_location(sv)
    mine = [int(all_ids[k]) for k in range(len(all_ids))]
    owner = {}
    for k, (verdict, consumed, st, frame) in enumerate(outs):
        if st == 0:
            _, b = W.parse_response_frame(frame, PB["WaitForStartingTaskResponse"])
            for x in b.grants:
                owner[x.task_grant_id] = x.servant_location
    held = [t for t in mine if owner[t] == loc][:2]
    running = [(70 + j, t, loc, "%064x" % (j + 1)) for j, t in enumerate(held)] + [(99, 555555, loc, "e" * 64)]
    i = int(loc.split(":")[0].split(".")[-1])
    body = one("Heartbeat", _hb_msg(i, running=running), loc.split(":")[0], PB["HeartbeatResponse"], now=3.0)
    r = twin.heartbeat(_hb_direct(i, running=running), now=3.0)
    assert list(body.expired_tasks) == r.expired_tasks == [555555]
    body = one("GetRunningTasks", PB["GetRunningTasksRequest"](), "10.2.0.1", PB["GetRunningTasksResponse"], now=3.0)
    got = [(t.servant_task_id, t.task_grant_id, t.servant_location, t.task_digest) for t in body.running_tasks]
    want = [(t.servant_task_id, t.task_grant_id, t.servant_location, t.task_digest) for t in twin.get_running_tasks()]
    assert got == want and len(got) == len(held)
    fr = PB["FreeTaskRequest"](token="usr")
    fr.task_grant_ids.extend(all_ids[:3])
    assert one("FreeTask", fr, "10.2.0.1", PB["FreeTaskResponse"], now=4.0).ByteSize() == 0
    assert twin.free_task("usr", all_ids[:3]) == 0
    assert (wire.dispatcher.servant_state() == twin.dispatcher.servant_state()).all()

    # -- protocol corner cases -----------------------------------------------------------------------
    good = W.request_frame("GetConfig", PB["GetConfigRequest"](token="usr"), 77)
    unknown = W.request_frame("yadcc.scheduler.SchedulerService.Nope", None, 78)
    stream = W.request_frame("GetConfig", PB["GetConfigRequest"](token="usr"), 79)
    stream = stream[:16] + stream[16:].replace(b"\x38\x01", b"\x38\x02", 1)  # method_type = STREAM
    gz = W.request_frame("GetConfig", PB["GetConfigRequest"](token="usr"), 80, compression_algorithm=2)
    nopayload = W.request_frame("GetRunningTasks", None, 81, flags=4)
    trailing = good + b"\x01\x02\x03"
    # packed and unpacked repeated ids, an unknown field, fields out of order: all must parse
    raw = bytes([0x2a, 0x03]) + b"xyz" + bytes([0x08, 0x05, 0x08, 0x06, 0x32, 0x03]) + b"usr" + bytes([0x0a, 0x02, 0x07, 0x08, 0x28, 0x10])
    meta = PB["RpcMeta"](correlation_id=82, method_type=1)
    meta.request_meta.method_name = W.SERVICE + "KeepTaskAlive"
    mb = meta.SerializeToString()
    odd = struct.pack("<IIII", W.MAGIC, len(mb), len(raw), 0) + mb + raw
    corrupt_body = struct.pack("<IIII", W.MAGIC, len(mb), 2, 0) + mb + b"\x0a\x7f"  # length runs past the end
    outs = wire.handle_frames([(good[:10], "1.1.1.1"), (good[:-1], "1.1.1.1"), (b"GET / HTTP/1.1\r\n\r\n", "1.1.1.1"),
                               (unknown, "1.1.1.1"), (stream, "1.1.1.1"), (gz, "1.1.1.1"), (nopayload, "1.1.1.1"),
                               (trailing, "1.1.1.1"), (odd, "1.1.1.1"), (corrupt_body, "1.1.1.1")], now=5.0)
    assert [o[0] for o in outs] == [0, 0, -1, 1, -1, 1, 1, 1, 1, -1]
    _check_frame(outs[3][3], PB["GetConfigResponse"], 78, 10, "Method [yadcc.scheduler.SchedulerService.Nope] is not implemented.")
    _check_frame(outs[5][3], PB["GetConfigResponse"], 80, 101)
    _check_frame(outs[6][3], PB["GetRunningTasksResponse"], 81, 0)
    assert outs[7][1] == len(good)  # only the frame is consumed, the trailing bytes stay
    body = _check_frame(outs[8][3], PB["KeepTaskAliveResponse"], 82, 0)
    assert len(body.statuses) == 4  # ids 5, 6 (unpacked) and 7, 8 (packed)
    trace += [o[3][16:] for o in outs]
    # body-level entry point
    st, desc, resp = wire.call(W.SERVICE + "GetConfig", PB["GetConfigRequest"](token="usr").SerializeToString(), "1.1.1.1", now=5.0)
    r = PB["GetConfigResponse"]()
    r.ParseFromString(resp)
    assert st == 0 and r.serving_daemon_token == twin.get_config("usr", now=5.0)[1]
    assert wire.call("nope", b"", "1.1.1.1")[0] == 10
    return trace
