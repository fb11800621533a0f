/* ydshard.h -- ONE scheduler whose pending queue is range-sharded over the GPUs of a node.
 *
 * BASELINE.json north_star / SURVEY.md 8(e) option 2: the FIFO queue is cut into `world` contiguous
 * ranges, rank g keeps range g in its HBM; the servant table (<= 288 KB) and all scheduler state
 * that decisions depend on (running_tasks per servant) are REPLICATED: every rank receives the same
 * heartbeats / expiration ticks through the ordinary calls of ydsched.h.  A solve produces exactly the
 * decisions of one TaskDispatcher fed the concatenated queue (task_dispatcher.cc:93-140; the coupling
 * that must survive the sharding is `++running_tasks` at :123-124).  Per solve the ranks exchange, over
 * NCCL (NVLink / NVSwitch):
 *
 *   1. all-gather   the class tables (which (digest, min_version) classes occur, 16 KB per rank)
 *   2. all-gather   per-class request counts  -> every request's FIFO rank inside its class
 *   3. all-reduce   the per-class request records the slot lists can reach (a disjointly written
 *                   buffer: an all-gather-v; <= one 8-byte record per (class, eligible slot))
 *   4. all-reduce   the per-servant claimed-slot counts (u32[S]) + per-rank grant counts + flags
 *
 * Between 3 and 4 every rank runs the same slot-side merge (solve_merge.cuh) on the same data and
 * keeps the verdicts of its own requests.  Task ids are the batch's FIFO ordinals (rank g's grants
 * follow those of the lower ranks).  Leases (TaskDesc, task_dispatcher.h:199-215) live on the rank
 * that holds the request; FreeTask is collective so that running_tasks stays replicated.
 *
 * The library dlopen()s libnccl.so.2 on yd_shard_init: a single-GPU process never needs it.
 */
#ifndef YDSHARD_H_
#define YDSHARD_H_

#include "ydsched.h"

#ifdef __cplusplus
extern "C" {
#endif

#define YD_SHARD_UNIQUE_ID_BYTES 128

/* Rank 0: a fresh ncclUniqueId to hand to every rank (any transport: a file, MPI, torch.distributed).
 * Returns 0, or non-zero if NCCL cannot be loaded. */
int yd_shard_unique_id(uint8_t out[YD_SHARD_UNIQUE_ID_BYTES]);

/* Collective: joins `s` (one handle per process, on its own GPU) to a communicator of `world` ranks. */
int yd_shard_init(yd_sched* s, int rank, int world, const uint8_t unique_id[YD_SHARD_UNIQUE_ID_BYTES]);
void yd_shard_finalize(yd_sched* s);

/* Collective, THE HOT PATH: the whole queue = ranks' `reqs_local` concatenated in rank order.  out_local[i]
 * is what the i-th request of this rank's range gets (same fields as yd_wait_for_starting_new_tasks;
 * task ids number the grants of the whole batch in FIFO order).  reqs_local == NULL: the range staged with
 * yd_stage_requests.  Returns 0; 2 if some digest component needs the sequential solver (several servants
 * behind one requestor IP, more than 32 classes on one component, or the last-resort rule fired) -- nothing
 * was decided then and every rank gets the same answer, the caller solves that batch on one rank. */
int yd_shard_wait_for_starting_new_tasks(yd_sched* s, int64_t now_ns, const yd_task_req* reqs_local, size_t n_local,
                                         yd_grant* out_local);

/* Collective FreeTask x n (task_dispatcher.cc:167-188): every rank passes the ids IT wants freed (any
 * subset, also none); a lease is released by the rank that holds it and the running_tasks decrements
 * reach every replica through one all-reduce. */
int yd_shard_free_tasks(yd_sched* s, const uint64_t* ids, size_t n);

/* Device time (ms, CUDA events on the solve stream) of the last sharded solve's phases: local kernels and
 * the four exchanges.  Returns 0 if there was none. */
typedef struct yd_shard_stats {
  float total_ms;       /* first kernel .. grants ready on the device */
  float exchange_ms[4]; /* the four NCCL exchanges, including waiting for the slowest rank */
  uint64_t exchange_bytes[4];
  uint64_t decisions_local, granted_local, granted_total;
  uint32_t merge_rounds, kernel_launches;
} yd_shard_stats;
int yd_shard_last_stats(yd_sched* s, yd_shard_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* YDSHARD_H_ */
