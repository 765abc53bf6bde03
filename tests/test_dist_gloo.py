"""N > 1 host logic on CPU (gloo, world_size 2): contiguous sharding by global env index + the statistics all-reduce.

Each rank steps ITS shard of a GridWorld batch (with the oracle restatement standing in for the GPU kernels, keyed by
the same global env indices the kernels use), packs the engine's statistics vector and sums it with
`rl_envs_forge_b200.all_reduce_stats` -- the engine's only collective.  The reduced integer statistics must equal a
single-process run over the whole batch, whatever the number of ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import philox
from oracle.gen_golden import CONFIG2_LAYOUT
from oracle.gridworld import FlatTables, GridWorldPort, vec_step
from rl_envs_forge_b200 import all_reduce_stats, shard_range, stats_to_dict

N, STEPS, SEED = 6001, 60, 13


def _run_shard(lo, hi):
    port = GridWorldPort(**dict(CONFIG2_LAYOUT, p_success=0.8))
    tab = FlatTables(port)
    n = hi - lo
    g = np.arange(lo, hi)
    pos, ln, ret = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.float32)
    stats = np.zeros(8)
    for t in range(STEPS):
        wt, wp = philox.step_words(SEED, g, np.full(n, t, np.uint64))
        o = vec_step(tab, pos, ln, ret, (wp >> np.uint32(30)).astype(np.int32), wt, limit=200, start_pos=0, autoreset=True)
        fin = o["finished"]
        stats[0] += fin.sum()
        stats[1] += o["final_len"][fin].sum()
        stats[2] += o["final_ret"][fin].astype(np.float64).sum()
        stats[3] += (o["final_ret"][fin].astype(np.float64) ** 2).sum()
        stats[4] += (~o["error"]).sum()
        pos, ln, ret = o["pos"], o["ep_len"], o["ep_ret"]
    return stats, pos


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(N, rank, world)
    stats, pos = _run_shard(lo, hi)
    t = torch.from_numpy(stats.copy())
    all_reduce_stats(t)
    np.save(os.path.join(out_dir, f"stats_{rank}.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"pos_{rank}.npy"), pos)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 1000, 2 ** 20 + 3):
        for w in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_all_reduce_is_noop_without_process_group():
    t = torch.arange(8, dtype=torch.float64)
    assert torch.equal(all_reduce_stats(t.clone()), t)
    d = stats_to_dict([2, 10, 3.0, 5.0, 100, 0, 0, 0])
    assert d["episodes"] == 2 and d["mean_length"] == 5 and d["mean_return"] == 1.5 and d["env_steps"] == 100


@pytest.mark.timeout(300)
def test_two_rank_stats_allreduce_matches_single_process(tmp_path):
    whole_stats, whole_pos = _run_shard(0, N)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    s0 = np.load(tmp_path / "stats_0.npy")
    s1 = np.load(tmp_path / "stats_1.npy")
    assert np.array_equal(s0, s1)                                   # every rank holds the global statistics
    assert s0[0] == whole_stats[0] and s0[1] == whole_stats[1] and s0[4] == whole_stats[4]   # integers: exact
    assert s0[2] == pytest.approx(whole_stats[2], rel=1e-12) and s0[3] == pytest.approx(whole_stats[3], rel=1e-12)
    assert whole_stats[0] > 0
    pos = np.concatenate([np.load(tmp_path / f"pos_{r}.npy") for r in range(world)])
    assert np.array_equal(pos, whole_pos)                           # trajectories do not depend on the sharding
