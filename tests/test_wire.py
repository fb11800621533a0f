"""FlareStd wire front end on the CPU backends: frames encoded/decoded by the google.protobuf
runtime against the hand-written codec, handlers checked against direct calls, and the verbatim
reference dispatcher (`ref`) against the restatement (`port`) frame by frame."""
import pytest

pytest.importorskip("google.protobuf")

from wire_cases import run_wire_scenario


def test_wire_scenario_port_equals_reference(make_dispatcher):
    a = run_wire_scenario(make_dispatcher, "ref")
    b = run_wire_scenario(make_dispatcher, "port")
    assert a == b


def test_wire_parser_survives_mutated_frames(make_dispatcher):
    """A network-facing parser must not crash or over-read on garbage: 6 000 mutated frames
    (bit flips, truncations, length-field corruption, random tails) only ever yield a verdict in
    {-1, 0, 1}; handled frames produce a well-formed response frame."""
    import struct

    import numpy as np

    import wire_protos as W
    from yadcc_b200.service import SchedulerService

    PB = W.PB
    svc = SchedulerService(make_dispatcher("port"), acceptable_user_tokens="usr", acceptable_servant_tokens="srv", token_seed=2)
    hb = PB["HeartbeatRequest"](token="srv", next_heartbeat_in_ms=1000, version=3, location="10.0.0.1:8335", num_processors=8,
                                capacity=4, servant_priority=2, total_memory_in_bytes=1 << 36, memory_available_in_bytes=1 << 35)
    hb.env_descs.add().compiler_digest = "d" * 64
    t = hb.running_tasks.add()
    t.servant_task_id, t.task_grant_id, t.servant_location, t.task_digest = 5, 6, "10.0.0.1:8335", "e" * 64
    wq = PB["WaitForStartingTaskRequest"](token="usr", immediate_reqs=2, prefetch_reqs=1, next_keep_alive_in_ms=1000, min_version=1)
    wq.env_desc.compiler_digest = "d" * 64
    ka = PB["KeepTaskAliveRequest"](token="usr", next_keep_alive_in_ms=1000)
    ka.task_grant_ids.extend(range(40))
    seeds = [W.request_frame("Heartbeat", hb, 1), W.request_frame("WaitForStartingTask", wq, 2),
             W.request_frame("KeepTaskAlive", ka, 3), W.request_frame("GetRunningTasks", None, 4, flags=4),
             W.request_frame("FreeTask", PB["FreeTaskRequest"](token="usr", task_grant_ids=[1, 2, 3]), 5)]
    rng = np.random.default_rng(0)
    batch, verdicts = [], {-1: 0, 0: 0, 1: 0}
    for it in range(6000):
        f = bytearray(seeds[it % len(seeds)])
        kind = it % 6
        if kind == 0:  # bit flips anywhere
            for _ in range(int(rng.integers(1, 6))):
                f[int(rng.integers(0, len(f)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:  # truncation
            f = f[: int(rng.integers(0, len(f)))]
        elif kind == 2:  # corrupt a header length
            struct.pack_into("<I", f, 4 * int(rng.integers(1, 4)), int(rng.integers(0, 1 << 32)))
        elif kind == 3:  # random bytes in the body
            lo = 16 + struct.unpack_from("<I", f, 4)[0]
            for k in range(lo, len(f)):
                if rng.random() < 0.2:
                    f[k] = int(rng.integers(0, 256))
        elif kind == 4:  # random bytes in the meta
            for k in range(16, min(len(f), 16 + struct.unpack_from("<I", f, 4)[0])):
                if rng.random() < 0.2:
                    f[k] = int(rng.integers(0, 256))
        else:  # junk appended
            f += bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
        batch.append((bytes(f), "10.0.0.%d" % (it % 200)))
        if len(batch) == 64:
            for (verdict, consumed, status, frame), (data, _) in zip(svc.handle_frames(batch, now=1.0 + it * 1e-3), batch):
                assert verdict in (-1, 0, 1)
                verdicts[verdict] += 1
                if verdict == 1:
                    assert 16 <= consumed <= len(data)
                    magic, ms, bs, att = struct.unpack("<IIII", frame[:16])
                    assert magic == W.MAGIC and att == 0 and len(frame) == 16 + ms + bs
                    meta = PB["RpcMeta"]()
                    meta.ParseFromString(frame[16:16 + ms])
                    assert meta.response_meta.status == status
                else:
                    assert consumed == 0 and frame == b""
            batch = []
    assert min(verdicts.values()) > 100  # all three outcomes really occurred
