"""VecPendulumDisk -- N PendulumDisk (DC-motor disk pendulum) instances (fp32) stepped by one sm_100a kernel launch.

Drop-in surface for the reference class `PendulumDisk`
(rl_envs_forge/envs/inverted_pendulum/pendulum_disk/pendulum_disk.py): constructor kwargs and defaults (:10-20),
constants (:58-64), `reset` (:94-106), `step` (:108-172), spaces (:49-55), `state`, `current_step`, `total_reward`.
"""
import numpy as np

from . import _lib, spaces
from ._pendulum_base import VecPendulumBase


class VecPendulumDisk(VecPendulumBase):
    DIM = 2
    _FN = "ef_pendulum_disk"
    _ACC_DIM = 1
    metadata = {"render.modes": []}

    def __init__(self, num_envs=1, tau=0.005, continuous_reward=False, episode_length_limit=1000,
                 angle_termination=None, initial_state=None, nonlinear_reward=False, reward_decay_rate=3.0,
                 reward_angle_range=(-0.2, 0.2), *, device=None, seed=None, env_offset=0, autoreset=True, strict=False,
                 alias_obs=True):
        # pendulum_disk.py:58-64
        self.J = 1.91e-4
        self.m = 0.055
        self.g = 9.81
        self.l = 0.042
        self.b = 3e-6
        self.K = 0.0536
        self.R = 9.5
        self.single_observation_space = spaces.Box(low=np.array([-np.pi, -15 * np.pi]),
                                                   high=np.array([np.pi, 15 * np.pi]), shape=(2,), dtype=np.float64)
        super().__init__(num_envs, tau, continuous_reward, episode_length_limit, angle_termination, initial_state,
                         nonlinear_reward, reward_decay_rate, reward_angle_range, device, seed, env_offset, autoreset,
                         strict, alias_obs)

    def _make_params(self):
        p = _lib.PendulumParams()
        c = [self.m * self.g * self.l / self.J, (self.b + self.K ** 2 / self.R) / self.J, self.K / (self.R * self.J),
             15 * np.pi]
        for i, v in enumerate(c):
            p.c[i] = v
        self._fill_common(p, float(np.pi))
        return p

    def _info(self, acc, truncated, force):
        return {"truncated": truncated, "alpha_dot_dot": acc[:, 0], "force": force}
