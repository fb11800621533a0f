/* ydservice.h -- SchedulerServiceImpl over the ydsched C ABI.
 *
 * The RPC handlers of yadcc/scheduler/scheduler_service_impl.cc:67-333 minus the wire:
 * token verification, argument limits, the NAT / unverified-servant -> max_tasks = 0 rules,
 * status codes (yadcc/api/scheduler.proto:23-35) and the rolling window of three
 * serving-daemon tokens.  A front end decodes a request, fills one of the structs below and
 * gets back exactly what the reference handler would have put on the wire.  Pure host
 * code: it is compiled into every library that exports the ydsched ABI and only calls that
 * ABI, so it behaves the same over the CUDA backend and over the reference oracle.
 *
 * Time is explicit (`now_ns`, steady clock), as everywhere in ydsched.h.  Calls on one
 * service (and its yd_sched) must be serialised by the caller.
 */
#ifndef YDSERVICE_H_
#define YDSERVICE_H_

#include "ydsched.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Status (scheduler.proto:23-35); 0 = success (rpc.STATUS_SUCCESS). */
#define YD_STATUS_OK 0
#define YD_STATUS_NO_QUOTA_AVAILABLE 1001
#define YD_STATUS_NOT_IMPLEMENTED 1002
#define YD_STATUS_ACCESS_DENIED 1003
#define YD_STATUS_INVALID_ARGUMENT 1004
#define YD_STATUS_VERSION_TOO_OLD 1005
#define YD_STATUS_ENVIRONMENT_NOT_AVAILABLE 1006

/* NotAcceptingTaskReason values the service itself sets (scheduler.proto:58,61). */
#define YD_REASON_BEHIND_NAT 4
#define YD_REASON_NOT_VERIFIED 100

typedef struct yd_service yd_service;

typedef struct yd_service_config {
  const char* acceptable_user_tokens;    /* --acceptable_user_tokens: comma separated, empty entries kept
                                          * (yadcc/common/token_verifier.cc:56-69) */
  const char* acceptable_servant_tokens; /* --acceptable_servant_tokens */
  int32_t min_daemon_version;            /* --min_daemon_version, default 0 (scheduler_service_impl.cc:35) */
  int32_t serving_daemon_token_rollout_interval_s; /* default 3600 (:38); <= 0 selects the default */
  uint64_t token_seed; /* serving-daemon tokens are 16 random bytes in hex (:46-51, RAND_bytes);
                        * 0 = std::random_device, anything else = reproducible sequence */
} yd_service_config;

/* SchedulerServiceImpl() (:55-65).  NULL if a token list is empty (the reference CHECK-fails,
 * token_verifier.cc:58-59) or the dispatcher handle is NULL. */
yd_service* yd_service_create(yd_sched* dispatcher, int64_t now_ns, const yd_service_config* cfg);
void yd_service_destroy(yd_service* svc);

/* HeartbeatRequest (scheduler.proto:63-118) + what the handler reads from the RPC controller. */
typedef struct yd_heartbeat_request {
  const char* token;
  const char* location;   /* "ip:port" / "[v6]:port" as reported by the servant */
  const char* remote_ip;  /* controller->GetRemotePeer(): the peer's IP as text, no port */
  uint32_t remote_is_ipv6;
  uint32_t next_heartbeat_in_ms;
  uint32_t version;
  uint32_t num_processors;
  uint32_t current_load;
  uint32_t servant_priority;          /* ServantPriority as received (may be out of range) */
  uint32_t not_accepting_task_reason;
  uint32_t capacity;
  uint32_t n_env_digests;
  uint64_t total_memory_in_bytes;
  uint64_t memory_available_in_bytes;
  const char* const* env_digests;
  const yd_running_task* running_tasks;
  size_t n_running_tasks;
} yd_heartbeat_request;

typedef struct yd_heartbeat_response {
  const char* acceptable_tokens[3]; /* valid until the next call on the service */
  uint64_t* expired_tasks;          /* caller's buffer, capacity >= n_running_tasks */
  size_t n_expired_tasks;
} yd_heartbeat_response;

/* Heartbeat (:67-194).  Returns the status the handler would set. */
int yd_service_heartbeat(yd_service* svc, int64_t now_ns, const yd_heartbeat_request* req,
                         yd_heartbeat_response* resp);

/* GetConfig (:196-208): *serving_daemon_token = the middle one of the three active tokens. */
int yd_service_get_config(yd_service* svc, int64_t now_ns, const char* token, const char** serving_daemon_token);

/* WaitForStartingTask (:209-271) for n RPCs at once: token check per RPC (tokens[i]), then
 * yd_wait_for_starting_task_rpcs for the accepted ones as ONE batch.  Rejected RPCs get
 * YD_STATUS_ACCESS_DENIED and no grants.  Returns the number of grants written. */
size_t yd_service_wait_for_starting_tasks(yd_service* svc, int64_t now_ns, const char* const* tokens,
                                          const yd_rpc_wait* rpcs, size_t n_rpcs, yd_rpc_wait_result* results,
                                          yd_grant* grants_out, size_t cap);

/* KeepTaskAlive (:272-292): statuses[i] = 0/1 per id. */
int yd_service_keep_task_alive(yd_service* svc, int64_t now_ns, const char* token, uint32_t next_keep_alive_in_ms,
                               const uint64_t* task_grant_ids, size_t n, uint8_t* statuses);

/* FreeTask (:294-308). */
int yd_service_free_task(yd_service* svc, const char* token, const uint64_t* task_grant_ids, size_t n);

/* GetRunningTasks (:310-317): no token check in the reference either. */
size_t yd_service_get_running_tasks(yd_service* svc, yd_running_task* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* YDSERVICE_H_ */
