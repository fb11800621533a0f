"""One logical scheduler over several GPUs (option 1).

Decisions couple only through per-servant `running_tasks`, so servants that share no
compiler digest never interact: the connected components of the digest<->servant graph
are independent FIFO sub-queues (the sharding the reference's authors propose at
yadcc/scheduler/task_dispatcher.h:286-288).  `ShardedDispatcher` gives every rank a
subset of the components -- their servants, their digests and the requests for them --
and one process per GPU runs an ordinary `TaskDispatcher` on its share.  No servant
state ever crosses ranks.

The only thing the shards share is the task-id space.  Two modes:

* "strided" (default): the library itself hands out id = local_id * world + rank
  (`yd_config.id_stride / id_offset`) and ignores ids that are not its own -- no
  communication at all.  Task-grant ids are opaque lease tokens on the wire
  (scheduler.proto:181-238), so uniqueness and routability are all the protocol needs.
* "fifo": exactly the numbering ONE reference scheduler would produce
  (`next_task_id++` in global FIFO order, task_dispatcher.cc:127): rank r must know how many
  requests EARLIER in the global queue were granted by other ranks, i.e. one all-reduce
  (sum) of the per-request grant flags per solve plus a local prefix sum.

Everything else (heartbeats, frees, keep-alives, ticks) is routed to the owning rank by
the caller-visible `owner_of_*` maps.

The class is backend-agnostic (any library speaking the ydsched C ABI) and
transport-agnostic (any torch.distributed backend), so the N>1 logic is tested on CPU
with gloo + the oracle and runs unchanged with NCCL on B200s.
"""
from __future__ import annotations

import zlib
from typing import Callable, Sequence

import numpy as np

from ._abi import GRANT_DTYPE, REQ_DTYPE, STATUS_ENVIRONMENT_NOT_FOUND, STATUS_GRANTED
from .dispatcher import Servant, TaskDispatcher


def default_digest_owner(digest: str, world: int) -> int:
    """Stable digest -> rank map.  Digests that co-occur on one servant must map to the
    same rank (they are one component); `keep_servant_alive` checks this."""
    return zlib.crc32(digest.encode()) % world


def component_digest_owner(servants: Sequence[Servant], world: int) -> Callable[[str, int], int]:
    """A digest -> rank map that keeps every component (digests joined through servants that hold several of them,
    the union-find of SyncTopology) on ONE rank: the owner is the default map applied to the component's smallest
    digest.  Digests no listed servant holds fall back to the default map.  Build it from the servant set the ranks
    agree on (all of them see every heartbeat) and pass it as `digest_owner`; with the plain default map a servant
    that advertises two compilers usually has digests on different ranks and `keep_servant_alive` refuses it."""
    parent: dict[str, str] = {}

    def find(x: str) -> str:
        parent.setdefault(x, x)
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for sv in servants:
        envs = list(sv.environments)
        for e in envs[1:]:
            a, b = find(envs[0]), find(e)
            if a != b:
                parent[max(a, b)] = min(a, b)
        if envs:
            find(envs[0])
    smallest: dict[str, str] = {}
    for d in list(parent):
        r = find(d)
        smallest[r] = min(smallest.get(r, d), d)

    def owner(digest: str, w: int) -> int:
        if digest in parent:
            return default_digest_owner(smallest[find(digest)], w)
        return default_digest_owner(digest, w)

    del world
    return owner


class ShardedDispatcher:
    def __init__(self, local: TaskDispatcher, rank: int, world: int, *, group=None, device=None,
                 digest_owner: Callable[[str, int], int] = default_digest_owner, id_mode: str = "strided"):
        """id_mode:
          "strided"  task id = local id * world + rank.  Needs `local` created with
                     id_stride=world, id_offset=rank (the library then hands out and accepts
                     such ids itself); no communication at all.  Ids are unique and routable
                     but not the single-scheduler numbering.
          "fifo"     exactly the ids one reference scheduler would hand out for the global
                     queue; costs one all-reduce of grant flags per solve."""
        assert id_mode in ("strided", "fifo")
        self.id_mode = id_mode
        self.local = local
        self.rank = rank
        self.world = world
        self.group = group
        self.device = device  # torch device for the collective (cuda:LOCAL_RANK with NCCL, cpu with gloo)
        self.digest_owner = digest_owner
        self.next_task_id = 0  # global id space
        # global task id -> local task id for the grants this rank owns: one chunk per solve
        # (global ids, local ids, alive flags), chunks ordered by first global id (ids only grow)
        self._chunks: list[list[np.ndarray]] = []
        self._owners_cache = None  # (owners array object, mine, mine on the collective's device)
        self.collective_bytes = 0

    # -- ownership -----------------------------------------------------------
    def owner_of_digest(self, digest: str) -> int:
        return self.digest_owner(digest, self.world)

    def owner_of_servant(self, servant: Servant) -> int:
        owners = {self.owner_of_digest(d) for d in servant.environments}
        if len(owners) > 1:
            raise ValueError(
                f"servant {servant.observed_location} holds digests owned by ranks {sorted(owners)}: "
                "the digest_owner map must keep a component on one rank"
            )
        return owners.pop() if owners else 0

    # -- servant maintenance ---------------------------------------------------
    def keep_servant_alive(self, servant: Servant, expires_in: float, *, now: float = 0.0) -> None:
        if self.owner_of_servant(servant) == self.rank:
            self.local.keep_servant_alive(servant, expires_in, now=now)

    def on_expiration_timer(self, *, now: float) -> None:
        self.local.on_expiration_timer(now=now)
        # "fifo" ids: leases that expired, were orphaned or swept never come back through free_tasks; once the library
        # holds no lease at all the whole id map is garbage
        if self._chunks and self.local.num_tasks() == 0:
            self._chunks = []

    # -- the hot path ------------------------------------------------------------
    def wait_for_starting_new_tasks(self, digests: Sequence[str], owners: np.ndarray, local_reqs: np.ndarray,
                                    now: float = 0.0) -> np.ndarray:
        """`owners[i]` = owning rank of global request i (or -1 if nobody holds its digest);
        `local_reqs` = REQ array of this rank's requests, in global order.  Returns a
        GRANT array for this rank's requests with GLOBAL task ids."""
        import torch
        import torch.distributed as dist

        del digests
        if self.id_mode == "strided":
            return self.local.wait_for_starting_new_tasks(local_reqs, now)
        if self._owners_cache is None or self._owners_cache[0] is not owners:
            mine = np.nonzero(owners == self.rank)[0]
            self._owners_cache = (owners, mine, torch.as_tensor(mine, device=self.device))
        _, mine, mine_t = self._owners_cache
        assert len(mine) == len(local_reqs)
        g = self.local.wait_for_starting_new_tasks(local_reqs, now).copy() if len(mine) else np.zeros(0, GRANT_DTYPE)
        ok = g["status"] == STATUS_GRANTED
        # the one exchange step: who was granted, over the whole global queue
        flags = torch.zeros(len(owners), dtype=torch.int32, device=self.device)
        if len(mine):
            flags[mine_t] = torch.as_tensor(ok.view(np.uint8)).to(self.device, non_blocking=True).to(torch.int32)
        if self.world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.SUM, group=self.group)
            self.collective_bytes += flags.numel() * 4
        csum = torch.cumsum(flags, 0)
        total = int(csum[-1].item()) if len(owners) else 0
        if len(mine):
            before = (csum - flags)[mine_t]  # grants strictly earlier in the global FIFO
            gids = np.uint64(self.next_task_id) + before.cpu().numpy().astype(np.uint64)
            if ok.any():
                self._chunks.append([gids[ok], g["task_id"][ok].copy(), np.ones(int(ok.sum()), dtype=bool)])
            g["task_id"][ok] = gids[ok]
        self.next_task_id += total
        return g

    def _lookup(self, global_ids):
        """Yields (chunk, positions in the chunk, positions in `global_ids`) for the ids this
        rank owns and still holds."""
        ids = np.asarray(global_ids, dtype=np.uint64)
        if not len(ids) or not self._chunks:
            return
        firsts = np.asarray([c[0][0] for c in self._chunks], dtype=np.uint64)
        which = np.searchsorted(firsts, ids, side="right").astype(np.int64) - 1
        for ci in np.unique(which[which >= 0]):
            gid, _, alive = self._chunks[ci]
            sel = np.nonzero(which == ci)[0]
            pos = np.searchsorted(gid, ids[sel])
            pos_c = np.minimum(pos, len(gid) - 1)
            hit = (pos < len(gid)) & (gid[pos_c] == ids[sel]) & alive[pos_c]
            if hit.any():
                yield self._chunks[ci], pos_c[hit], sel[hit]

    # -- lease maintenance, routed by global id -------------------------------------
    def free_tasks(self, global_ids) -> None:
        if self.id_mode == "strided":  # the library ignores ids that are not its own
            self.local.free_tasks(global_ids)
            return
        for chunk, pos, _ in self._lookup(global_ids):
            p = np.unique(pos)  # an id listed twice frees once (FreeTask of an unknown id is a no-op)
            self.local.free_tasks(chunk[1][p])
            chunk[2][p] = False
        self._chunks = [c for c in self._chunks if c[2].any()]

    def keep_tasks_alive(self, global_ids, new_expires_in: float, *, now: float = 0.0) -> np.ndarray:
        """Statuses for the ids this rank owns (False for ids owned elsewhere; the caller
        ORs the ranks' answers)."""
        if self.id_mode == "strided":
            return self.local.keep_tasks_alive(global_ids, new_expires_in, now=now)
        out = np.zeros(len(np.asarray(global_ids)), dtype=bool)
        for chunk, pos, where in self._lookup(global_ids):
            out[where] = self.local.keep_tasks_alive(chunk[1][pos], new_expires_in, now=now)
        return out


class RangeShardedDispatcher:
    """ONE scheduler whose pending queue is range-sharded over the GPUs of a node (include/ydshard.h;
    SURVEY.md 8(e) option 2, BASELINE.json north_star): rank g keeps the g-th contiguous FIFO range in
    its HBM, the servant table is replicated (every rank is fed the same heartbeats and ticks through
    the ordinary TaskDispatcher calls of `local`), and a solve makes exactly the decisions one
    TaskDispatcher would make on the concatenated queue.  The exchanges (class tables, per-class
    counts, reachable request records, per-servant claimed-slot counts) are NCCL collectives issued
    by the C++ library on its own stream; torch.distributed only carries the 128-byte ncclUniqueId
    here.  CUDA library only."""

    def __init__(self, local: TaskDispatcher, rank: int, world: int, *, device=None, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _abi

        self.local, self.rank, self.world = local, rank, world
        self.group = group
        lib = local._lib
        # Libraries without the NCCL path (the CPU oracles in the gloo tests): the same contract -- one queue cut into
        # per-rank ranges, replicated servant state, FIFO task ids, collective FreeTask -- with the exchange restated
        # over torch.distributed: the ranges are all-gathered and every rank decides the whole queue on its replica.
        self.native = hasattr(lib, "yd_shard_init")
        if not self.native:
            return
        buf = (C.c_uint8 * _abi.SHARD_UNIQUE_ID_BYTES)()
        if rank == 0 and lib.yd_shard_unique_id(buf) != 0:
            raise RuntimeError("yd_shard_unique_id failed (libnccl.so.2 not loadable?)")
        t = torch.tensor(list(buf), dtype=torch.uint8, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.broadcast(t, src=0, group=group)
        uid = (C.c_uint8 * _abi.SHARD_UNIQUE_ID_BYTES)(*t.cpu().tolist())
        rc = lib.yd_shard_init(local._h, rank, world, uid)
        if rc != 0:
            raise RuntimeError(f"yd_shard_init failed: {rc}")

    def wait_for_starting_new_tasks(self, reqs_local, now: float, out=None):
        """Collective.  reqs_local: this rank's FIFO range (REQ_DTYPE), or an int n = the first n staged
        requests (TaskDispatcher.stage_requests).  Returns this rank's grants, or None if the batch has to
        be solved on one rank (yd_shard_wait_for_starting_new_tasks returned 2: nothing was decided)."""
        from .dispatcher import _ns

        lib, h = self.local._lib, self.local._h
        if not self.native:
            import torch.distributed as dist

            parts: list = [None] * self.world
            dist.all_gather_object(parts, np.ascontiguousarray(reqs_local), group=self.group)
            lo = sum(len(p) for p in parts[: self.rank])
            g = self.local.wait_for_starting_new_tasks(np.concatenate(parts), now)
            return g[lo:lo + len(reqs_local)].copy()
        if isinstance(reqs_local, (int, np.integer)):
            n, ptr = int(reqs_local), None
        else:
            assert reqs_local.dtype == REQ_DTYPE and reqs_local.flags.c_contiguous
            n, ptr = reqs_local.shape[0], reqs_local.ctypes.data
        if out is None:
            out = np.zeros(max(n, 1), dtype=GRANT_DTYPE)
        rc = lib.yd_shard_wait_for_starting_new_tasks(h, _ns(now), ptr, n, out.ctypes.data)
        if rc == 2:
            return None
        if rc != 0:
            raise RuntimeError(f"yd_shard_wait_for_starting_new_tasks failed: {rc}")
        return out[:n]

    def free_tasks(self, ids) -> None:
        """Collective FreeTask: every rank passes the ids it wants released (its own grants, typically)."""
        ids = np.ascontiguousarray(np.asarray(ids, dtype=np.uint64))
        if not self.native:
            import torch.distributed as dist

            parts: list = [None] * self.world
            dist.all_gather_object(parts, ids, group=self.group)
            self.local.free_tasks(np.concatenate(parts))  # every replica holds every lease
            return
        rc = self.local._lib.yd_shard_free_tasks(self.local._h, ids.ctypes.data if len(ids) else None, len(ids))
        if rc != 0:
            raise RuntimeError(f"yd_shard_free_tasks failed: {rc}")

    def last_stats(self) -> dict | None:
        import ctypes as C

        from . import _abi

        if not self.native:
            return None
        st = _abi.yd_shard_stats()
        if not self.local._lib.yd_shard_last_stats(self.local._h, C.byref(st)):
            return None
        return {"total_ms": st.total_ms, "exchange_ms": list(st.exchange_ms), "exchange_bytes": list(st.exchange_bytes),
                "decisions_local": st.decisions_local, "granted_local": st.granted_local, "granted_total": st.granted_total,
                "merge_rounds": st.merge_rounds, "kernel_launches": st.kernel_launches}

    def close(self) -> None:
        if self.native:
            self.local._lib.yd_shard_finalize(self.local._h)
