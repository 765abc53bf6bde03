"""The DEVICE vs the unmodified reference driven by the engine's own Philox streams (tests/golden/engine_*.npz): no oracle
in between.  Env i of the batch is one reference instance fed the engine's words of (seed, global env index, step) and the
engine's random policy, so `rollout(K)` with the in-kernel policy -- and `step()` fed the policy's host twin -- must
reproduce what the reference returned.  Integer work bit-exact; bandit rewards (fp32 device vs the reference's float64)
within rtol 1e-5 + atol 1e-6 (BASELINE.json north_star)."""
import json

import numpy as np
import pytest

from oracle import philox
from oracle.gen_golden import BANDIT_ARMS
from tests._golden import GOLDEN, grid_kwargs, load

pytestmark = pytest.mark.gpu


def _g(name):
    return load(f"{GOLDEN}/engine_{name}.npz")


@pytest.mark.parametrize("name", ["gridworld_config2_p08", "gridworld_config2_p1", "gridworld_config2_p08_4096"])
def test_gridworld_device_reproduces_engine_driven_reference(name):
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorld
    g = _g(name)
    cfg = grid_kwargs(g)
    cfg["special_transitions"] = {(k[0], Action(int(k[1]))): v for k, v in cfg["special_transitions"].items()}
    seed, off = int(g["seed"]), int(g["env_offset"])
    T, n = g["reward"].shape
    gi = np.arange(n) + off
    env = VecGridWorld(n, seed=seed, env_offset=off, **cfg)
    raised = 0
    for t in range(T):
        wp = philox.group_word(seed, gi, t, philox.STREAM_POLICY)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy((wp >> np.uint32(30)).astype(np.int32)).cuda())
        f = g["flags"][t]
        assert np.array_equal(obs.cpu().numpy(), g["obs"][t].astype(np.int32)), t
        assert np.array_equal(rew.cpu().numpy(), g["reward"][t]), t
        assert np.array_equal(term.cpu().numpy(), (f & 1) > 0) and np.array_equal(trunc.cpu().numpy(), (f & 2) > 0), t
        raised += int(((f & 4) > 0).sum())
    assert int(env.err[0]) == raised                                  # env-steps the reference raised on
    # the fused rollout with the in-kernel policy: per-step rewards / flags and the final positions
    env2 = VecGridWorld(n, seed=seed, env_offset=off, **cfg)
    rew, flg = env2.rollout(T, return_rewards=True)
    assert np.array_equal(rew.cpu().numpy(), g["reward"]) and np.array_equal(flg.cpu().numpy(), g["flags"] & 3)
    assert np.array_equal(env2.state.cpu().numpy(), g["obs"][-1].astype(np.int32))


def test_gambler_device_reproduces_engine_driven_reference():
    from rl_envs_forge_b200 import VecGamblersProblem
    g = _g("gambler_default")
    kw = json.loads(str(g["config"]))
    T, n = g["reward"].shape
    env = VecGamblersProblem(n, seed=int(g["seed"]), env_offset=int(g["env_offset"]), **kw)
    rew, flg = env.rollout(T, return_rewards=True)                    # in-kernel policy: uniform legal bet
    assert np.array_equal(rew.cpu().numpy(), g["reward"]) and np.array_equal(flg.cpu().numpy() & 1, g["done"])
    assert np.array_equal(env.state.cpu().numpy(), g["state"][-1]) and int(env.err[0]) == 0
    assert env.episode_stats()["episodes"] == int(g["done"].sum())


def test_car_rental_device_reproduces_engine_driven_reference():
    import torch
    from rl_envs_forge_b200 import VecJacksCarRental
    g = _g("carrental_default")
    kw = json.loads(str(g["config"]))
    seed, off = int(g["seed"]), int(g["env_offset"])
    T, n = g["reward"].shape
    gi = np.arange(n) + off
    env = VecJacksCarRental(n, seed=seed, env_offset=off, init_state_option="equal", **kw)
    for t in range(T):
        wp = philox.group_word(seed, gi, t, philox.STREAM_POLICY).astype(np.uint64)
        act = ((wp * np.uint64(11)) >> np.uint64(32)).astype(np.int32)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
        assert np.array_equal(obs.cpu().numpy(), g["state"][t]) and np.array_equal(rew.cpu().numpy(), g["reward"][t]), t
    env2 = VecJacksCarRental(n, seed=seed, env_offset=off, init_state_option="equal", **kw)
    rew, _ = env2.rollout(T, return_rewards=True)
    assert np.array_equal(rew.cpu().numpy(), g["reward"]) and np.array_equal(env2.state.cpu().numpy(), g["state"][-1])


def test_bandit_device_reproduces_engine_driven_reference():
    from rl_envs_forge_b200 import VecKArmedBandit
    g = _g("bandit_config3_custom_arms")
    k = int(g["k"])
    T, n = g["reward"].shape
    env = VecKArmedBandit(n, k=k, arm_params={a: dict(BANDIT_ARMS[a]) for a in range(k)}, seed=int(g["seed"]),
                          env_offset=int(g["env_offset"]))
    obs, rew = env.rollout(T)                                         # in-kernel policy: uniform random arm
    assert np.array_equal(obs.cpu().numpy(), g["action"])
    dev, ref = rew.cpu().numpy().astype(np.float64), g["reward"]
    bad = np.abs(dev - ref) > 1e-5 * np.abs(ref) + 1e-6
    assert int(bad.sum()) == 0, f"{int(bad.sum())} of {bad.size} rewards outside rtol 1e-5 (accept/reject flips?)"
