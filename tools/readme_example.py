import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_envs_forge_b200 import (Action, VecGridWorld, VecGridWorldLayouts, VecKArmedBandit, VecCartPole, VecPendulumDisk,
                                VecGamblersProblem, VecJacksCarRental, VecLabyrinth)

env = VecGridWorld(1 << 20, rows=8, cols=8, walls={(1, 1), (2, 5)}, terminal_states={(7, 7): 1.0},
                   special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)}, p_success=0.8, episode_length_limit=200, seed=0)
obs, info = env.reset()
actions = torch.randint(0, 4, (env.num_envs,), dtype=torch.int32, device="cuda")
obs, reward, terminated, truncated, info = env.step(actions)
rewards, flags = env.rollout(64, return_rewards=True)
print(env.episode_stats())
env.mdp, env.P
cp = VecCartPole(1 << 22, seed=0)
cp.step(torch.empty(cp.num_envs, device="cuda").uniform_(-3, 3))
bandit = VecKArmedBandit(1 << 24, k=10, arm_params={0: dict(distribution="gamma", shape=9, scale=0.55)}, seed=0)
bandit.step(torch.randint(0, 10, (bandit.num_envs,), dtype=torch.int32, device="cuda"))
obs, reward, *_ = env.step(actions.cpu().numpy())
exact = VecGridWorld(4, seed=[7, 8, 9, 10], rng="pcg64")
print(type(obs), obs.shape, exact.step(torch.zeros(4, dtype=torch.int32, device="cuda"))[0].tolist())
