"""tools/scheduler_server.py over a real TCP socket on 127.0.0.1 (CPU oracle as backend): a
client speaking FlareStd through the google.protobuf runtime heartbeats three servants, takes
grants from two connections at once, renews and frees them."""
import asyncio
import struct
import sys

import pytest

pytest.importorskip("google.protobuf")

from conftest import PORT_LIB, ROOT

sys.path.insert(0, str(ROOT / "tools"))
import wire_protos as W

PB = W.PB


async def _rpc(reader, writer, method, msg, corr, resp_cls):
    writer.write(W.request_frame(method, msg, corr))
    await writer.drain()
    hdr = await reader.readexactly(16)
    _, ms, bs, att = struct.unpack("<IIII", hdr)
    rest = await reader.readexactly(ms + bs + att)
    return W.parse_response_frame(hdr + rest, resp_cls)


async def _scenario(port):
    import scheduler_server as srv

    args = srv.parse_args(["--port", str(port), "--user-tokens", "usr", "--servant-tokens", "srv", "--library", str(PORT_LIB),
                           "--window-ms", "5"])
    ready = asyncio.Event()
    task = asyncio.create_task(srv.serve(args, ready))
    await asyncio.wait_for(ready.wait(), 10)
    try:
        r1, w1 = await asyncio.open_connection("127.0.0.1", port)
        r2, w2 = await asyncio.open_connection("127.0.0.1", port)
        for i in range(3):
            hb = PB["HeartbeatRequest"](token="srv", next_heartbeat_in_ms=5000, version=9, location=f"127.0.0.1:{9000 + i}",
                                        num_processors=8, capacity=2, servant_priority=2)
            hb.env_descs.add().compiler_digest = "c" * 64
            meta, body = await _rpc(r1, w1, "Heartbeat", hb, 10 + i, PB["HeartbeatResponse"])
            assert meta.response_meta.status == 0 and len(body.acceptable_tokens) == 3
        wq = PB["WaitForStartingTaskRequest"](token="usr", immediate_reqs=2, next_keep_alive_in_ms=5000)
        wq.env_desc.compiler_digest = "c" * 64
        # two connections ask at the same time: one batching window, one solve
        (m1, b1), (m2, b2) = await asyncio.gather(_rpc(r1, w1, "WaitForStartingTask", wq, 21, PB["WaitForStartingTaskResponse"]),
                                                  _rpc(r2, w2, "WaitForStartingTask", wq, 22, PB["WaitForStartingTaskResponse"]))
        assert m1.correlation_id == 21 and m2.correlation_id == 22
        ids = [g.task_grant_id for g in list(b1.grants) + list(b2.grants)]
        assert m1.response_meta.status == m2.response_meta.status == 0 and sorted(ids) == [0, 1, 2, 3]
        # requestors are on 127.0.0.1 like the servants: the self rule keeps them off the first servant
        # until nothing else is free (task_dispatcher.cc:372-396)
        assert all(g.servant_location.startswith("127.0.0.1:900") for g in list(b1.grants) + list(b2.grants))
        ka = PB["KeepTaskAliveRequest"](token="usr", next_keep_alive_in_ms=5000)
        ka.task_grant_ids.extend(ids + [99])
        meta, body = await _rpc(r2, w2, "KeepTaskAlive", ka, 23, PB["KeepTaskAliveResponse"])
        assert list(body.statuses) == [True] * 4 + [False]
        fr = PB["FreeTaskRequest"](token="usr")
        fr.task_grant_ids.extend(ids)
        meta, _ = await _rpc(r1, w1, "FreeTask", fr, 24, PB["FreeTaskResponse"])
        assert meta.response_meta.status == 0
        meta, body = await _rpc(r2, w2, "KeepTaskAlive", ka, 25, PB["KeepTaskAliveResponse"])
        assert not any(body.statuses)
        meta, _ = await _rpc(r1, w1, "GetConfig", PB["GetConfigRequest"](token="nope"), 26, PB["GetConfigResponse"])
        assert meta.response_meta.status == 1003
        w1.close()
        w2.close()
    finally:
        task.cancel()
        try:
            await task
        except (asyncio.CancelledError, Exception):
            pass


def test_scheduler_server_over_tcp():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    asyncio.run(asyncio.wait_for(_scenario(port), 60))
