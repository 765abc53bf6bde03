"""Episode statistics off the stepping stream: one fold kernel per reporting interval + the cross-GPU sum on a side stream.

The env instances never exchange data (grid_world.py:495-528, k_armed_bandit.py:175-200, cart_pole.py:107-180 and
pendulum_disk.py:108-172 touch only `self`); the one thing a multi-GPU job shares is the 64-byte vector of finished-episode
statistics.  `StatsReducer` keeps that exchange off the critical path:

    fold()            ONE tiny kernel on the stepping stream (`ef_stats_fold`): the replicated accumulators of all env sets
                      of this rank -> `local` (8 doubles).  Stream-ordered, so it is a consistent snapshot, and capturable in
                      a CUDA graph.
    exchange_async()  sum of `local` over the ranks -> `total`, on a SIDE stream behind an event, so the stepping stream goes
                      on with the next launches while the 64 bytes cross NVLink.  Transport "peer": `ef_stats_fold_allreduce`
                      (one kernel: P2P stores into every peer's mailbox, acquire-spin on the local one, rank-ordered sum);
                      transport "nccl": `torch.distributed.all_reduce`.
    wait()            the stepping stream waits for the side stream's event (no host sync); `total` is then valid.

Integer-valued entries (episodes, length sum, env-steps) are exact, and both transports add the ranks' vectors in rank
order, so every rank holds bit-identical totals.
"""
import ctypes as C
import os

from . import _lib


def fold_stats(env_sets, out=None):
    """`ef_stats_fold` over the `_stats_raw` buffers of `env_sets` (all on one device) -> float64 [EF_STATS_LEN] tensor."""
    env0 = env_sets[0]
    torch = env0._torch
    if len(env_sets) > _lib.EF_STATS_MAX_SETS:
        raise ValueError(f"at most {_lib.EF_STATS_MAX_SETS} env sets per fold")
    if out is None:
        out = torch.empty(_lib.EF_STATS_LEN, dtype=torch.float64, device=env0.device)
    ptrs = (C.c_void_p * len(env_sets))(*[e._stats_raw.data_ptr() for e in env_sets])
    with torch.cuda.device(env0.device):
        _lib.check(_lib.load().ef_stats_fold(ptrs, len(env_sets), out.data_ptr(),
                                             torch.cuda.current_stream(env0.device).cuda_stream))
    return out


class StatsReducer:
    """Folds the statistics of the env sets of THIS rank and sums them over the ranks of a process group.

    transport  "auto" (peer memory when it can be set up, else NCCL), "peer", "nccl".  With no process group (or one
               rank) the exchange is a device copy."""

    def __init__(self, env_sets, transport="auto", group=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.env_sets = list(env_sets)
        if not self.env_sets:
            raise ValueError("StatsReducer needs at least one env set")
        self.device = self.env_sets[0].device
        if any(e.device != self.device for e in self.env_sets):
            raise ValueError("all env sets of a StatsReducer must live on one device")
        self._lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.local = torch.zeros(_lib.EF_STATS_LEN, dtype=torch.float64, device=self.device)
        self.total = torch.zeros(_lib.EF_STATS_LEN, dtype=torch.float64, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._ptrs = (C.c_void_p * len(self.env_sets))(*[e._stats_raw.data_ptr() for e in self.env_sets])
        self.side = torch.cuda.Stream(device=self.device)
        self._ev_folded = torch.cuda.Event()
        self._ev_done = torch.cuda.Event()
        self._comm = None
        transport = os.environ.get("ENVFORGE_STATS_TRANSPORT", transport)
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be 'auto', 'peer' or 'nccl'")
        self.transport = "local" if self.world == 1 else transport
        if self.world > 1 and transport in ("auto", "peer"):
            ok = self._connect_peers()
            if not ok and transport == "peer":
                raise _lib.EnvForgeError("peer-memory statistics exchange could not be set up (cudaIpc / P2P unavailable)")
            self.transport = "peer" if ok else "nccl"

    # -- peer mailbox setup: every rank exports a 64-byte IPC handle, all_gather, open the others ------------------------
    def _connect_peers(self):
        torch, dist = self._torch, self._dist
        handle = (C.c_uint8 * _lib.EF_PEER_HANDLE_BYTES)()
        comm = C.c_void_p()
        with torch.cuda.device(self.device):
            status = self._lib.ef_peer_comm_create(self.rank, self.world, C.byref(comm), handle)
        mine = torch.tensor(list(handle) + [0 if status == 0 else 1], dtype=torch.uint8, device=self.device)
        everyone = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(everyone, mine, group=self.group)
        rows = [t.cpu().tolist() for t in everyone]
        ok = all(r[-1] == 0 for r in rows)
        if ok:
            blob = (C.c_uint8 * (_lib.EF_PEER_HANDLE_BYTES * self.world))(*[b for r in rows for b in r[:-1]])
            with torch.cuda.device(self.device):
                ok = self._lib.ef_peer_comm_connect(comm, blob) == 0
        # every rank must agree, or some would spin on peers that fell back to NCCL
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        ok = bool(flag.item())
        if ok:
            self._comm = comm
        elif status == 0:
            self._lib.ef_peer_comm_destroy(comm)
        return ok

    # -- the three steps ------------------------------------------------------------------------------------------------
    def fold(self):
        """One kernel on the current stream: statistics of all env sets -> `local`."""
        torch = self._torch
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ef_stats_fold(self._ptrs, len(self.env_sets), self.local.data_ptr(),
                                               torch.cuda.current_stream(self.device).cuda_stream))
        return self.local

    def exchange_async(self):
        """Sum `local` (as of the last fold() on the current stream) over the ranks into `total`, on the side stream."""
        torch = self._torch
        main = torch.cuda.current_stream(self.device)
        self._ev_folded.record(main)
        self.side.wait_event(self._ev_folded)
        with torch.cuda.stream(self.side):
            self._exchange(self.side)
        self._ev_done.record(self.side)

    def _exchange(self, stream):
        if self.transport == "peer":
            with self._torch.cuda.device(self.device):
                _lib.check(self._lib.ef_stats_fold_allreduce(self._comm, None, 0, self.local.data_ptr(), self.total.data_ptr(),
                                                             self._status.data_ptr(), stream.cuda_stream))
        else:
            self.total.copy_(self.local)
            if self.transport == "nccl":
                self._dist.all_reduce(self.total, op=self._dist.ReduceOp.SUM, group=self.group)

    def wait(self):
        """The current stream waits for the exchange started by exchange_async(); returns `total`."""
        self._torch.cuda.current_stream(self.device).wait_event(self._ev_done)
        return self.total

    def reduce(self):
        """fold + exchange + wait on the current stream's timeline (the exchange still runs on the side stream)."""
        self.fold()
        self.exchange_async()
        return self.wait()

    def reduce_inline(self):
        """Peer transport only: fold and exchange fused in ONE kernel on the current stream (`ef_stats_fold_allreduce` with
        the raw accumulators) -- the lowest-latency form when nothing is there to overlap with."""
        if self.transport != "peer":
            return self.reduce()
        torch = self._torch
        with torch.cuda.device(self.device):
            _lib.check(self._lib.ef_stats_fold_allreduce(self._comm, self._ptrs, len(self.env_sets), self.local.data_ptr(),
                                                         self.total.data_ptr(), self._status.data_ptr(),
                                                         torch.cuda.current_stream(self.device).cuda_stream))
        return self.total

    def check(self):
        """Raise if a peer never showed up in an exchange (the kernel gives up after ~10 s instead of hanging the GPU)."""
        if int(self._status.item()) != 0:
            raise _lib.EnvForgeError("statistics exchange timed out waiting for a peer rank")

    def close(self):
        if self._comm is not None:
            self._torch.cuda.synchronize(self.device)
            self._lib.ef_peer_comm_destroy(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
