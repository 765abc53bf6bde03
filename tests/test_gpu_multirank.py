"""Sharding invariance ON HARDWARE: the same seeds stepped by 1 rank and by 2 NCCL ranks (one process per
GPU, contiguous global env index ranges) give identical all-reduced INTEGER episode statistics -- instances are independent
(grid_world.py:495-528, cart_pole.py:107-180), and the Philox streams are keyed by the global env index.

Both transports of the statistics exchange are held to it: the fused fold + peer-memory kernel (`ef_stats_fold_allreduce`,
P2P stores over NVLink) and NCCL.  Needs >= 2 visible GPUs (skipped otherwise): run it with `gpurun --gpus 2`."""
import json
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

N_TOTAL, K_ROLL, N_STEP, SEED = 1 << 16, 96, 24, 77
CFG = dict(rows=8, cols=8, start_state=(0, 0),
           walls={(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)},
           terminal_states={(7, 7): 1.0, (0, 7): -1.0}, episode_length_limit=40, p_success=0.8)


def _run(rank, world, transport):
    """Steps this rank's shard (a fused rollout, then graph-free step() calls with the in-kernel policy's host twin left out:
    rollout only uses Philox) and returns the reduced statistics of GridWorld and CartPole as lists."""
    from rl_envs_forge_b200 import Action, StatsReducer, VecCartPole, VecGridWorld, shard_range
    lo, hi = shard_range(N_TOTAL, rank, world)
    spec = {((3, 3), Action.DOWN): ((6, 6), 0.5)}
    out = {}
    gw = VecGridWorld(hi - lo, seed=SEED, special_transitions=spec, env_offset=lo, **CFG)
    gw.rollout(K_ROLL)
    for _ in range(N_STEP):
        gw.rollout(1)
    cp = VecCartPole(hi - lo, seed=SEED, env_offset=lo, episode_length_limit=50)
    cp.rollout(K_ROLL)
    for name, env in (("gridworld", gw), ("cartpole", cp)):
        red = StatsReducer([env], transport=transport)
        tot = red.reduce()
        torch.cuda.synchronize()
        red.check()
        inline = red.reduce_inline().clone()      # fused fold + exchange in one kernel (peer) must agree bit for bit
        torch.cuda.synchronize()
        assert torch.equal(inline, tot)
        out[name] = tot.tolist()
        out[name + "_transport"] = red.transport
        red.close()
    return out


def _worker(rank, world, port, out_dir, transport):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    res = _run(rank, world, transport)
    with open(os.path.join(out_dir, f"{transport}_{rank}.json"), "w") as f:
        json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_two_nccl_ranks_reproduce_the_single_rank_statistics(tmp_path, transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    torch.cuda.set_device(0)
    whole = _run(0, 1, "auto")
    assert whole["gridworld_transport"] == "local" and whole["gridworld"][0] > 0 and whole["cartpole"][0] > 0
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), transport), nprocs=2, join=True)
    got = [json.load(open(tmp_path / f"{transport}_{r}.json")) for r in range(2)]
    for name in ("gridworld", "cartpole"):
        assert got[0][name + "_transport"] == transport
        assert got[0][name] == got[1][name]                       # every rank holds the same totals, bit for bit
        for i in (0, 1, 4):                                       # episodes, length sum, env-steps: integers, exact
            assert got[0][name][i] == whole[name][i], (name, i, got[0][name], whole[name])
        for i in (2, 3):                                          # float sums: same addends, different association
            assert got[0][name][i] == pytest.approx(whole[name][i], rel=1e-12)


def test_fold_kernel_matches_the_replica_sum():
    """`ef_stats_fold` over several env sets == the plain sum of their replicated accumulators (integers exact)."""
    from rl_envs_forge_b200 import VecGridWorld, _lib, fold_stats
    envs = [VecGridWorld(4096, seed=SEED + i, env_offset=4096 * i, **CFG) for i in range(3)]
    for e in envs:
        e.rollout(64)
    got = fold_stats(envs).cpu()
    want = sum(e._stats_raw.view(_lib.EF_STATS_REPLICAS, _lib.EF_STATS_STRIDE)[:, :_lib.EF_STATS_LEN].sum(dim=0) for e in envs).cpu()
    assert got[0] > 0 and got[4] == 3 * 4096 * 64 - sum(int(e.err[0].item()) for e in envs)
    for i in (0, 1, 4):
        assert got[i] == want[i]
    assert torch.allclose(got, want, rtol=1e-12, atol=0)
    assert torch.equal(envs[0].stats.cpu(), fold_stats([envs[0]]).cpu())
