"""google.protobuf message classes for the scheduler's wire messages, built at run time from
hand-written descriptors (there is no protoc in this image).  Field names, numbers and types are
transcribed from yadcc/api/scheduler.proto, yadcc/api/env_desc.proto and
flare/rpc/protocol/protobuf/rpc_meta.proto; the google.protobuf runtime is the independent
encoder/decoder the hand-written C++ codec (include/ydwire_impl.inc) is pinned against."""
import struct

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

F = descriptor_pb2.FieldDescriptorProto
_T = {"string": F.TYPE_STRING, "bytes": F.TYPE_BYTES, "uint32": F.TYPE_UINT32, "uint64": F.TYPE_UINT64,
      "int32": F.TYPE_INT32, "bool": F.TYPE_BOOL}


def _msg(fd, name, fields):
    """fields: (name, number, type[, label]) with type a scalar name, '.pkg.Message' or 'enum:.pkg.Enum'."""
    m = fd.message_type.add()
    m.name = name
    for f in fields:
        fname, num, typ = f[:3]
        label = f[3] if len(f) > 3 else "optional"
        fld = m.field.add()
        fld.name, fld.number = fname, num
        fld.label = {"optional": F.LABEL_OPTIONAL, "repeated": F.LABEL_REPEATED, "required": F.LABEL_REQUIRED}[label]
        if typ.startswith("enum:"):
            fld.type, fld.type_name = F.TYPE_ENUM, typ[5:]
        elif typ.startswith("."):
            fld.type, fld.type_name = F.TYPE_MESSAGE, typ
        else:
            fld.type = _T[typ]
    return m


def _enum(fd, name, values):
    e = fd.enum_type.add()
    e.name = name
    for k, v in values:
        x = e.value.add()
        x.name, x.number = k, v


def build():
    pool = descriptor_pool.DescriptorPool()
    env = descriptor_pb2.FileDescriptorProto(name="yadcc/api/env_desc.proto", package="yadcc", syntax="proto3")
    _msg(env, "EnvironmentDesc", [("compiler_digest", 1, "string")])  # env_desc.proto:24-29
    pool.Add(env)

    s = descriptor_pb2.FileDescriptorProto(name="yadcc/api/scheduler.proto", package="yadcc.scheduler", syntax="proto3")
    s.dependency.append("yadcc/api/env_desc.proto")
    _enum(s, "ServantPriority", [("SERVANT_PRIORITY_UNKNOWN", 0), ("SERVANT_PRIORITY_DEDICATED", 1),
                                 ("SERVANT_PRIORITY_USER", 2)])  # scheduler.proto:39-48
    _msg(s, "RunningTask", [("servant_task_id", 1, "uint64"), ("task_grant_id", 2, "uint64"),
                            ("servant_location", 3, "string"), ("task_digest", 7, "string")])  # :233-238
    _msg(s, "HeartbeatRequest", [  # :63-118
        ("token", 13, "string"), ("next_heartbeat_in_ms", 7, "uint32"), ("version", 11, "uint32"),
        ("location", 1, "string"), ("num_processors", 10, "uint32"), ("current_load", 4, "uint32"),
        ("servant_priority", 14, "enum:.yadcc.scheduler.ServantPriority"), ("not_accepting_task_reason", 12, "uint32"),
        ("capacity", 3, "uint32"), ("total_memory_in_bytes", 15, "uint64"), ("memory_available_in_bytes", 16, "uint64"),
        ("env_descs", 6, ".yadcc.EnvironmentDesc", "repeated"), ("running_tasks", 17, ".yadcc.scheduler.RunningTask", "repeated")])
    _msg(s, "HeartbeatResponse", [("acceptable_tokens", 2, "string", "repeated"),
                                  ("expired_tasks", 1, "uint64", "repeated")])  # :120-141
    _msg(s, "GetConfigRequest", [("token", 1, "string")])
    _msg(s, "GetConfigResponse", [("serving_daemon_token", 1, "string")])
    _msg(s, "StartingTaskGrant", [("task_grant_id", 1, "uint64"), ("servant_location", 2, "string")])  # :171-179
    _msg(s, "WaitForStartingTaskRequest", [  # :181-201
        ("token", 6, "string"), ("milliseconds_to_wait", 1, "uint32"), ("env_desc", 2, ".yadcc.EnvironmentDesc"),
        ("immediate_reqs", 3, "uint32"), ("prefetch_reqs", 4, "uint32"), ("next_keep_alive_in_ms", 5, "uint32"),
        ("min_version", 7, "uint32")])
    _msg(s, "WaitForStartingTaskResponse", [("grants", 1, ".yadcc.scheduler.StartingTaskGrant", "repeated")])
    _msg(s, "KeepTaskAliveRequest", [("token", 6, "string"), ("task_grant_ids", 1, "uint64", "repeated"),
                                     ("next_keep_alive_in_ms", 5, "uint32")])
    _msg(s, "KeepTaskAliveResponse", [("statuses", 1, "bool", "repeated")])
    _msg(s, "FreeTaskRequest", [("token", 2, "string"), ("task_grant_ids", 1, "uint64", "repeated")])
    _msg(s, "FreeTaskResponse", [])
    _msg(s, "GetRunningTasksRequest", [])
    _msg(s, "GetRunningTasksResponse", [("running_tasks", 1, ".yadcc.scheduler.RunningTask", "repeated")])
    pool.Add(s)

    r = descriptor_pb2.FileDescriptorProto(name="flare/rpc/protocol/protobuf/rpc_meta.proto", package="flare.rpc",
                                           syntax="proto2")
    _enum(r, "MethodType", [("METHOD_TYPE_UNKNOWN", 0), ("METHOD_TYPE_SINGLE", 1), ("METHOD_TYPE_STREAM", 2)])
    _enum(r, "CompressionAlgorithm", [("COMPRESSION_ALGORITHM_UNKNOWN", 0), ("COMPRESSION_ALGORITHM_NONE", 1),
                                      ("COMPRESSION_ALGORITHM_GZIP", 2), ("COMPRESSION_ALGORITHM_LZ4_FRAME", 3),
                                      ("COMPRESSION_ALGORITHM_SNAPPY", 4), ("COMPRESSION_ALGORITHM_ZSTD", 5)])
    _msg(r, "RpcRequestMeta", [("method_name", 2, "string", "required"), ("request_id", 3, "uint32"), ("timeout", 4, "uint32"),
                               ("tracing_context", 5, "bytes"), ("acceptable_compression_algorithms", 6, "uint64")])
    _msg(r, "RpcResponseMeta", [("status", 1, "int32", "required"), ("description", 2, "string"),
                                ("trace_forcibly_sampled", 3, "bool")])
    _msg(r, "RpcMeta", [("correlation_id", 1, "uint64", "required"), ("method_type", 7, "enum:.flare.rpc.MethodType", "required"),
                        ("flags", 8, "uint64"), ("compression_algorithm", 9, "enum:.flare.rpc.CompressionAlgorithm"),
                        ("attachment_compressed", 10, "bool"), ("request_meta", 5, ".flare.rpc.RpcRequestMeta"),
                        ("response_meta", 6, ".flare.rpc.RpcResponseMeta")])
    pool.Add(r)

    def cls(name):
        return message_factory.GetMessageClass(pool.FindMessageTypeByName(name))

    names = ["yadcc.EnvironmentDesc"] + ["yadcc.scheduler." + n for n in (
        "RunningTask", "HeartbeatRequest", "HeartbeatResponse", "GetConfigRequest", "GetConfigResponse", "StartingTaskGrant",
        "WaitForStartingTaskRequest", "WaitForStartingTaskResponse", "KeepTaskAliveRequest", "KeepTaskAliveResponse",
        "FreeTaskRequest", "FreeTaskResponse", "GetRunningTasksRequest", "GetRunningTasksResponse")] + [
        "flare.rpc.RpcMeta", "flare.rpc.RpcRequestMeta", "flare.rpc.RpcResponseMeta"]
    return {n.split(".")[-1]: cls(n) for n in names}


PB = build()
MAGIC = (ord("F") << 24) | (ord("R") << 16) | (ord("P") << 8) | ord("C")  # std_protocol.cc:63
SERVICE = "yadcc.scheduler.SchedulerService."


def request_frame(method: str, body_msg, correlation_id: int, **meta_kw) -> bytes:
    """What flare's client side puts on the wire for a unary call (std_protocol.cc:248-311)."""
    meta = PB["RpcMeta"](correlation_id=correlation_id, method_type=1, **meta_kw)
    meta.request_meta.method_name = method if "." in method else SERVICE + method
    meta.request_meta.timeout = 5000
    mb = meta.SerializeToString()
    body = body_msg.SerializeToString() if body_msg is not None else b""
    return struct.pack("<IIII", MAGIC, len(mb), len(body), 0) + mb + body


def parse_response_frame(frame: bytes, resp_cls):
    magic, ms, bs, att = struct.unpack("<IIII", frame[:16])
    assert magic == MAGIC and att == 0 and len(frame) == 16 + ms + bs
    meta = PB["RpcMeta"]()
    meta.ParseFromString(frame[16:16 + ms])
    body = resp_cls()
    body.ParseFromString(frame[16 + ms:])
    return meta, body
