"""VecCartPole -- N CartPole instances (fp32) stepped by one sm_100a kernel launch.

Drop-in surface for the reference class `CartPole`
(rl_envs_forge/envs/inverted_pendulum/cart_pole/cart_pole.py): constructor kwargs and defaults (:11-21), constants
(:56-64), `reset(seed, options, initial_state)` (:91-105), `step(action)` (:107-180), spaces (:50-53),
`state`, `current_step`, `total_reward`.  The reference computes in float64 per instance; here state is fp32 [N,4]
on the device (1e-5 relative per-step tolerance, see tests/test_gpu_pendulums.py).
"""
import numpy as np

from . import _lib, spaces
from ._pendulum_base import VecPendulumBase


class VecCartPole(VecPendulumBase):
    DIM = 4
    _FN = "ef_cartpole"
    _ACC_DIM = 2
    metadata = {"render.modes": []}

    def __init__(self, num_envs=1, tau=0.02, continuous_reward=False, episode_length_limit=1000,
                 angle_termination=None, initial_state=None, nonlinear_reward=False, reward_decay_rate=3.0,
                 reward_angle_range=(-0.1, 0.1), *, device=None, seed=None, env_offset=0, autoreset=True, strict=False,
                 alias_obs=True):
        # cart_pole.py:56-64
        self.gravity = 9.8
        self.mass_cart = 1.0
        self.mass_pole = 0.1
        self.total_mass = self.mass_cart + self.mass_pole
        self.length = 0.5
        self.pole_mass_length = self.mass_pole * self.length
        self.force_mag = 10.0
        self.theta_threshold_radians = 0.2
        self.x_threshold = 2.4
        self.single_observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(4,), dtype=np.float64)
        super().__init__(num_envs, tau, continuous_reward, episode_length_limit, angle_termination, initial_state,
                         nonlinear_reward, reward_decay_rate, reward_angle_range, device, seed, env_offset, autoreset,
                         strict, alias_obs)

    def _make_params(self):
        p = _lib.PendulumParams()
        c = [self.gravity, self.pole_mass_length, 1.0 / self.total_mass, self.length, 4.0 / 3.0,
             self.mass_pole / self.total_mass, self.pole_mass_length / self.total_mass, self.x_threshold]
        for i, v in enumerate(c):
            p.c[i] = v
        self._fill_common(p, self.theta_threshold_radians)
        return p

    def _info(self, acc, truncated, force):
        return {"truncated": truncated, "x_acc": acc[:, 0], "theta_acc": acc[:, 1], "force": force}
