"""Common host-side plumbing of the vectorised envs: device buffers, stream, statistics, error word, host staging.

PyTorch is used for device memory, streams and torch.distributed only; every transition runs in the sm_100a kernels
behind the C ABI (rl_envs_forge_b200/_lib.py).  No CPU / PyTorch fallback exists: constructing an env without a CUDA
device or without the built library raises.
"""
import os

import numpy as np

from . import _lib
from ._lib import EnvForgeError


def resolve_seed(seed):
    """None -> fresh 63-bit entropy (the reference seeds from OS entropy too); ints are taken mod 2^64."""
    if seed is None:
        return int.from_bytes(os.urandom(8), "little") >> 1
    if not isinstance(seed, (int, np.integer)):
        raise ValueError("seed must be an integer or None")
    return int(seed) & 0xFFFFFFFFFFFFFFFF


def shard_range(num_envs_global, rank, world_size):
    """Contiguous index range [lo, hi) of `rank`; ranges differ by at most one env and cover [0, N) exactly."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(num_envs_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


STAT_FIELDS = ("episodes", "length_sum", "return_sum", "return_sq_sum", "env_steps")


def stats_to_dict(vec):
    """EF_STATS_LEN doubles -> readable dict (means derived)."""
    v = [float(x) for x in vec]
    d = {k: v[i] for i, k in enumerate(STAT_FIELDS)}
    n = d["episodes"]
    d["mean_return"] = d["return_sum"] / n if n else float("nan")
    d["mean_length"] = d["length_sum"] / n if n else float("nan")
    return d


def all_reduce_stats(stats, group=None):
    """Sum the statistics vector over all ranks (NCCL over NVLink for CUDA tensors, gloo for CPU tensors).

    This is the only collective of the engine: env instances never exchange data.  No-op without an initialised
    process group.  Returns the (in-place reduced) tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


class _Residency:
    """Which env sets are alive on which device, and how many bytes a step of each touches: the basis of the
    EF_STEP_COLD_INPUTS hint (include/envforge_b200.h).  When the live env sets of a device together exceed its L2, a set's
    arrays have most likely been evicted by the others' traffic by the time it is stepped again."""

    def __init__(self):
        import weakref
        self._sets = weakref.WeakSet()
        self.generation = 0      # bumped whenever the population changes; envs cache their verdict against it
        self._l2 = {}

    def add(self, env):
        self._sets.add(env)
        self.generation += 1

    def bump(self, *_):
        self.generation += 1

    def l2_bytes(self, torch, device):
        idx = device.index
        if idx not in self._l2:
            self._l2[idx] = int(getattr(torch.cuda.get_device_properties(device), "L2_cache_size", 0)) or (126 << 20)
        return self._l2[idx]

    def cold(self, env):
        total = sum(e._step_bytes for e in self._sets if e.device == env.device)
        return total > self.l2_bytes(env._torch, env.device)


RESIDENCY = _Residency()


class VecEnv:
    """Base class: owns the SoA buffers shared by all env families (episode counters, stats, error word)."""

    _HOST_TRANSPORT = "zerocopy"   # default host-buffer transport of the family (see __init__)

    def __init__(self, num_envs, device=None, seed=None, env_offset=0, autoreset=True):
        import torch
        if not isinstance(num_envs, (int, np.integer)) or num_envs <= 0:
            raise ValueError("num_envs must be a positive integer")
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise EnvForgeError("rl_envs_forge_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._torch = torch
        dev = torch.device("cuda" if device is None else device)
        if dev.type != "cuda":
            raise EnvForgeError(f"device must be a CUDA device, got {dev}")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.num_envs = int(num_envs)
        self.env_offset = int(env_offset)
        if self.env_offset < 0 or self.env_offset + self.num_envs > 2 ** 32:
            raise ValueError("env_offset + num_envs must stay below 2**32")
        self.autoreset = bool(autoreset)
        self.seed = seed
        self._seed = resolve_seed(seed)
        self._t = 0            # step counter: the Philox time coordinate of the next step
        self._reset_epoch = 0  # counts explicit resets (Philox coordinate of STREAM_RESET_USER)
        n = self.num_envs
        self.ep_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ep_ret = torch.zeros(n, dtype=torch.float32, device=dev)
        # statistics accumulators: EF_STATS_REPLICAS copies (CTAs spread their atomics), summed on read -- see `stats`
        self._stats_raw = torch.zeros(_lib.EF_STATS_REPLICAS * _lib.EF_STATS_STRIDE, dtype=torch.float64, device=dev)
        self.err = torch.zeros(_lib.EF_ERR_LEN, dtype=torch.int32, device=dev)
        self._host = {}
        # how step() moves HOST buffers (numpy / CPU tensors in, numpy out):
        #   "staged"    copy-engine H2D -> one kernel -> D2H, serialised
        #   "zerocopy"  one launch that reads the actions from and writes the outputs to pinned host memory itself
        #   "pipelined" ef_*_step_host: the batch is cut into `host_chunks` ranges; the kernel of a chunk reads its actions
        #               straight from pinned host memory while the copy engine drains the previous chunk's outputs
        # The default is per env family, chosen from measurements on B200 / PCIe Gen5 (profiles/r01_host_transport.md):
        # the link moves ~50 GB/s in total whichever way the bytes are scheduled, so the fewest operations win.
        # ENVFORGE_HOST_TRANSPORT / ENVFORGE_HOST_CHUNKS override the defaults for A/B timing.
        self.host_transport = os.environ.get("ENVFORGE_HOST_TRANSPORT", self._HOST_TRANSPORT)
        if self.host_transport not in ("pipelined", "zerocopy", "staged"):
            raise ValueError("host_transport must be 'pipelined', 'zerocopy' or 'staged'")
        self.host_chunks = int(os.environ.get("ENVFORGE_HOST_CHUNKS", "4"))
        self.host_stage_actions = False   # pipelined: True = copy-engine H2D of the actions instead of zero-copy reads
        self._pipe_handle = None
        self._clock = None  # optional device-resident step clock (CUDA-graph capture), see enable_device_clock()
        # bytes one step of this env set touches (state, action, outputs: the canonical 42 B per env is close enough for
        # every family) -- see _Residency
        self._step_bytes = 42 * self.num_envs
        self._cold_cache = (-1, False)
        import weakref
        RESIDENCY.add(self)
        weakref.finalize(self, RESIDENCY.bump)

    def _cold_inputs(self):
        """True when the live env sets on this device together exceed its L2 (cached per population change)."""
        gen, verdict = self._cold_cache
        if gen != RESIDENCY.generation:
            verdict = RESIDENCY.cold(self)
            self._cold_cache = (RESIDENCY.generation, verdict)
        return verdict

    # -- device clock ---------------------------------------------------------------------------------------------
    def enable_device_clock(self):
        """Keep the Philox time coordinate in device memory so that step()/rollout() can be captured in a CUDA graph and
        replayed: the kernels read and advance the clock themselves instead of receiving `t` as a baked-in argument.
        The device word counts in units of 2**-CLOCK_SHIFT steps (every CTA adds its share when it retires)."""
        if self._clock is None:
            torch = self._torch
            self._clock = torch.tensor([self._t << _lib.EF_CLOCK_SHIFT, 0], dtype=torch.int64, device=self.device)
        return self._clock

    def disable_device_clock(self):
        if self._clock is not None:
            self._t = int(self._clock[0].item()) >> _lib.EF_CLOCK_SHIFT
            self._clock = None

    def _t_arg(self):
        return 0 if self._clock is not None else self._t

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _stream(self):
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def _call(self, fn, *args):
        with self._torch.cuda.device(self.device):
            _lib.check(fn(*args, self._stream()))

    def _call_lean(self, fn, *args):
        """_call without the device context manager when the env's device is already current (the usual case)."""
        torch = self._torch
        if torch.cuda.current_device() == self.device.index:
            status = fn(*args, torch.cuda.current_stream().cuda_stream)
            if status:
                _lib.check(status)
        else:
            self._call(fn, *args)

    def _pipe(self):
        """Lazily created ef_host_pipe (two copy streams + events) of the pipelined host-buffer step."""
        if self._pipe_handle is None:
            h = _lib.C.c_void_p()
            with self._torch.cuda.device(self.device):
                _lib.check(self._lib.ef_host_pipe_create(_lib.C.byref(h)))
            self._pipe_handle = h
        return self._pipe_handle

    def _device_scratch(self, name, shape, dtype):
        t = self._host.get("d_" + name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self._torch.empty(shape, dtype=dtype, device=self.device)
            self._host["d_" + name] = t
        return t

    def _pinned(self, name, shape, dtype):
        t = self._host.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self._torch.empty(shape, dtype=dtype, pin_memory=True)
            self._host[name] = t
        return t

    def _pinned_ring(self, name, shape, dtype, depth=4):
        """Pinned staging buffer for a host->device copy that nothing waits for on the host (a table re-uploaded every
        step).  `depth` buffers rotate, each guarded by a CUDA event recorded after ITS copy was enqueued (`mark_staged`):
        a buffer is rewritten only once the DMA that read it has completed, however far the host runs ahead of the stream.
        Returns (tensor, slot); call `_mark_staged(slot)` right after enqueuing the copy."""
        key = "ring_" + name
        ring = self._host.get(key)
        if ring is None or ring["shape"] != tuple(shape) or ring["dtype"] != dtype:
            ring = {"shape": tuple(shape), "dtype": dtype, "next": 0,
                    "slots": [[self._torch.empty(shape, dtype=dtype, pin_memory=True), None] for _ in range(depth)]}
            self._host[key] = ring
        i = ring["next"]
        ring["next"] = (i + 1) % len(ring["slots"])
        buf, ev = ring["slots"][i]
        if ev is not None:
            ev.synchronize()          # host blocks only if the copy issued `depth` uploads ago is still in flight
        return buf, (key, i)

    def _mark_staged(self, slot):
        key, i = slot
        ev = self._torch.cuda.Event()
        ev.record(self._torch.cuda.current_stream(self.device))
        self._host[key]["slots"][i][1] = ev

    def _host_in(self, name, array, dtype, shape):
        """Host array -> pinned CPU tensor of `dtype`/`shape`.  A pinned torch tensor of the right dtype is used as is (the
        device reads the caller's buffer); anything else is copied into a pinned staging buffer first."""
        torch = self._torch
        src = array if torch.is_tensor(array) else torch.as_tensor(np.ascontiguousarray(array))
        if tuple(src.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(src.shape)}")
        if src.is_pinned() and src.dtype == dtype and src.is_contiguous():
            return src
        stage = self._pinned("h_" + name, shape, dtype)
        stage.copy_(src)
        return stage

    def _to_device(self, name, array, dtype, shape):
        """Host array -> device tensor on the current stream (copy-engine DMA from pinned memory)."""
        torch = self._torch
        stage = self._host_in(name, array, dtype, shape)
        dev = self._host.get("d_" + name)
        if dev is None or tuple(dev.shape) != tuple(shape) or dev.dtype != dtype:
            dev = torch.empty(shape, dtype=dtype, device=self.device)
            self._host["d_" + name] = dev
        dev.copy_(stage, non_blocking=True)
        return dev

    def _to_host(self, named_tensors):
        """Device tensors -> numpy arrays through pinned buffers; one stream sync for the whole batch.

        The arrays are views of the pinned staging buffers (no extra host copy): they stay valid until the next
        host-buffer step() of this env, like the device tensors step() returns on the CUDA path."""
        torch = self._torch
        outs = []
        for name, t in named_tensors:
            stage = self._pinned("o_" + name, t.shape, t.dtype)
            stage.copy_(t, non_blocking=True)
            outs.append(stage)
        torch.cuda.current_stream(self.device).synchronize()
        return [o.numpy() for o in outs]

    # -- statistics / errors ---------------------------------------------------------------------------------------
    @property
    def stats(self):
        """float64 [EF_STATS_LEN] device tensor: the statistics vector (sum over the replicas the kernels add into),
        folded by one `ef_stats_fold` launch on the current stream (fixed order of additions: bit-reproducible)."""
        from .stats import fold_stats
        return fold_stats([self])

    def episode_stats(self, all_reduce=False, reset=False):
        """Finished-episode statistics accumulated on device since construction (or the last reset=True call).

        all_reduce=True sums them over the ranks of the default process group first (one NCCL all-reduce of 64 B)."""
        vec = self.stats
        if all_reduce:
            all_reduce_stats(vec)   # blocking convenience form; StatsReducer keeps the exchange off the stepping stream
        if reset:
            self._stats_raw.zero_()
        return stats_to_dict(vec.cpu().tolist())

    def running_episode_stats(self):
        """(count, mean length, mean return) over the episodes currently in flight (device reduction)."""
        out = self._torch.zeros(_lib.EF_STATS_REPLICAS * _lib.EF_STATS_STRIDE, dtype=self._torch.float64, device=self.device)
        ln = self.ep_len   # a computed temporary in the packed GridWorld layout: keep it alive across the launch
        self._call(self._lib.ef_episode_stat_reduce, _lib.ptr(ln), _lib.ptr(self.ep_ret), _lib.ptr(out), self.num_envs)
        del ln
        out = out.view(_lib.EF_STATS_REPLICAS, _lib.EF_STATS_STRIDE)[:, :_lib.EF_STATS_LEN].sum(dim=0)
        return stats_to_dict(out.cpu().tolist())

    def check_errors(self, clear=True):
        """Raise ValueError if any env-step since the last check hit a condition the reference raises on.

        Lazy by design: it costs a device->host read, so the batched path does not call it per step (strict mode does)."""
        e = [int(x) & 0xFFFFFFFF for x in self.err.cpu().tolist()]
        if clear and e[0]:
            self.err.zero_()
        if e[0]:
            packed = e[2] | (e[3] << 32)      # (1 + global env index) << 32 | kind << 24 | detail
            env, kind_id, detail = (packed >> 32) - 1, (packed >> 24) & 0xFF, packed & 0xFFFFFF
            kind = _lib.STEP_ERROR_NAMES.get(kind_id, "unknown")
            raise ValueError(self._describe_error(e[0], env, detail, kind_id, kind))

    def _describe_error(self, count, env, detail, kind_id, kind):
        return f"{count} env-step(s) rejected ({kind}); first offender env {env}, detail {detail}"

    @property
    def step_count(self):
        return self._t if self._clock is None else int(self._clock[0].item()) >> _lib.EF_CLOCK_SHIFT

    def close(self):
        if getattr(self, "_pipe_handle", None) is not None:
            self._torch.cuda.current_stream(self.device).synchronize()
            self._lib.ef_host_pipe_destroy(self._pipe_handle)
            self._pipe_handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
