/* ydwire.h -- flare "std" wire front end for the scheduler service.
 *
 * An unmodified yadcc daemon talks to its scheduler in FlareStd frames
 * (flare/rpc/protocol/protobuf/std_protocol.cc:52-66):
 *
 *     u32le magic 'FRPC' | u32le meta_size | u32le msg_size | u32le att_size
 *     RpcMeta (proto2, flare/rpc/protocol/protobuf/rpc_meta.proto:151-191)
 *     message body (proto3, yadcc/api/scheduler.proto)   [attachment: unused by this service]
 *
 * This layer decodes request frames, runs the handlers of ydservice.h and encodes the
 * response frames exactly as flare's server side does (service.cc:657-681: correlation id
 * echoed, METHOD_TYPE_SINGLE, response_meta.status = controller.ErrorCode(), description set
 * iff the call failed).  Hand-written protobuf codec (no protoc / libprotobuf in the build);
 * the tests pin every message against the google.protobuf runtime.  Socket handling stays
 * with the caller: bytes in, bytes out.
 *
 * Consecutive WaitForStartingTask frames of one call are decided as ONE batch (one GPU solve);
 * the answers are what handling the frames one after the other would give.
 *
 * Not supported (answered like flare answers an unknown method / dropped like a corrupt
 * frame): compressed bodies, streaming RPCs, attachments.
 */
#ifndef YDWIRE_H_
#define YDWIRE_H_

#include "ydservice.h"

#ifdef __cplusplus
extern "C" {
#endif

/* rpc::Status values this layer produces itself (rpc_meta.proto:31-86). */
#define YD_RPC_STATUS_METHOD_NOT_FOUND 10
#define YD_RPC_STATUS_NOT_SUPPORTED 101

typedef struct yd_wire_in {
  const uint8_t* data;   /* bytes received from one connection, starting at a frame boundary */
  size_t len;
  const char* remote_ip; /* the connection's peer address (text, no port) */
  uint32_t remote_is_ipv6;
  uint32_t reserved;
} yd_wire_in;

typedef struct yd_wire_out {
  size_t consumed; /* bytes of `data` that made up the frame (0 unless verdict == 1) */
  size_t offset;   /* response frame = out[offset .. offset + len) */
  size_t len;
  int32_t verdict; /* 1 = frame handled; 0 = incomplete frame, read more; -1 = not FlareStd / corrupt: close */
  int32_t status;  /* status put into the response meta (0 = success) */
} yd_wire_out;

/* Handles the FIRST frame of each of the n inputs, in array order.  Response frames are written
 * back to back into out[0 .. out_cap).  Returns the number of bytes written, or (size_t)-1 if
 * out_cap is too small (nothing is handled in that case is NOT guaranteed: size the buffer
 * generously -- 64 KiB per input is ample). */
size_t yd_wire_handle_frames(yd_service* svc, int64_t now_ns, const yd_wire_in* in, size_t n, uint8_t* out,
                             size_t out_cap, yd_wire_out* outs);

/* Body level, one call: `method` = MethodDescriptor::full_name(), e.g.
 * "yadcc.scheduler.SchedulerService.Heartbeat".  Writes the serialized response message to
 * resp[0 .. *resp_len) and returns the status for the response meta; *description (may be
 * NULL) receives the error text (valid until the next call on the service). */
int yd_wire_call(yd_service* svc, int64_t now_ns, const char* method, const char* remote_ip,
                 uint32_t remote_is_ipv6, const uint8_t* req, size_t req_len, uint8_t* resp, size_t resp_cap,
                 size_t* resp_len, const char** description);

#ifdef __cplusplus
}
#endif
#endif /* YDWIRE_H_ */
