"""TEST INFRASTRUCTURE ONLY (oracle) -- float64 CPU restatement of the reference CartPole / PendulumDisk hot path.

Follows mariusdgm/rl-envs-forge v5.22.0:
    rl_envs_forge/envs/inverted_pendulum/cart_pole/cart_pole.py      constants :56-64, normalize_angle :84-89,
                                                                     reset :91-105, step :107-180
    rl_envs_forge/envs/inverted_pendulum/pendulum_disk/pendulum_disk.py  constants :58-64, normalize_angle :87-92,
                                                                     reset :94-106, step :108-172

Functions are written over numpy arrays (N envs at once, float64); a scalar call is just N = 1.

numpy-version note: pendulum_disk.py:119 multiplies the python float K/R by the np.float32 action.
numpy >= 2 (NEP 50, what the build container has) rounds that product to float32; the numpy 1.26.4 the reference
pins (poetry.lock:786-787) keeps it in float64.  `nep50=True` reproduces the container's live reference bit-for-bit
in that term (used to pin this file against the golden vectors); `nep50=False` is the pinned-numpy arithmetic and is
what the device is compared against.  The two differ by <= 6e-8 relative in alpha_dot_dot.

Parity status: PINNED (tests/golden/cartpole_*.npz, pendulum_*.npz from oracle/gen_golden.py; live reference in
tests/test_oracle_vs_reference.py).  Nothing in the product package may import this module.
"""
import numpy as np

PI = np.pi


def normalize_angle(a):
    """cart_pole.py:84-89 / pendulum_disk.py:87-92: (a + pi) floored-mod 2pi - pi, with -pi mapped to +pi."""
    a = np.remainder(np.asarray(a, np.float64) + PI, 2 * PI) - PI
    return np.where(a == -PI, PI, a)


def _reward(angle, *, continuous, nonlinear, decay, span, lo, hi):
    """cart_pole.py:145-156 / pendulum_disk.py:138-149 (span = 0.2 for CartPole, pi for PendulumDisk)."""
    if continuous:
        if nonlinear:
            return 1.0 - (1.0 - np.exp(-decay * np.abs(angle))) / (1.0 - np.exp(-decay * span))
        return 1.0 - np.abs(angle) / span
    return np.where((lo <= angle) & (angle <= hi), 1.0, 0.0)


class CartPoleParams:
    def __init__(self, tau=0.02, continuous_reward=False, episode_length_limit=1000, angle_termination=None,
                 nonlinear_reward=False, reward_decay_rate=3.0, reward_angle_range=(-0.1, 0.1)):
        self.tau = tau
        self.continuous_reward = continuous_reward
        self.episode_length_limit = episode_length_limit
        self.angle_termination = angle_termination
        self.nonlinear_reward = nonlinear_reward
        self.reward_decay_rate = reward_decay_rate
        self.reward_angle_range = reward_angle_range
        # cart_pole.py:56-64
        self.gravity = 9.8
        self.mass_cart = 1.0
        self.mass_pole = 0.1
        self.total_mass = self.mass_cart + self.mass_pole
        self.length = 0.5
        self.pole_mass_length = self.mass_pole * self.length
        self.theta_threshold_radians = 0.2
        self.x_threshold = 2.4


def cartpole_step(p: CartPoleParams, state, force, current_step):
    """One reference CartPole.step for N envs.  state float64[N,4]; force = action[0] values (float32 in the
    reference, promoted to float64 by the first addition); current_step int[N] = steps taken so far.
    Returns dict(state, reward, terminated, truncated, x_acc, theta_acc, current_step)."""
    state = np.asarray(state, np.float64)
    x, x_dot, theta, theta_dot = (state[..., i] for i in range(4))
    force = np.asarray(force, np.float64)
    cos_t, sin_t = np.cos(theta), np.sin(theta)
    temp = (force + p.pole_mass_length * theta_dot ** 2 * sin_t) / p.total_mass
    theta_acc = (p.gravity * sin_t - cos_t * temp) / (
        p.length * (4.0 / 3.0 - p.mass_pole * cos_t ** 2 / p.total_mass))
    x_acc = temp - p.pole_mass_length * theta_acc * cos_t / p.total_mass
    x = x + p.tau * x_dot
    x_dot = x_dot + p.tau * x_acc
    theta = normalize_angle(theta + p.tau * theta_dot)
    theta_dot = theta_dot + p.tau * theta_acc
    done = ((x < -p.x_threshold) | (x > p.x_threshold)
            | (theta < -p.theta_threshold_radians) | (theta > p.theta_threshold_radians))
    if p.angle_termination is not None:
        done = done | (np.abs(theta) > p.angle_termination)
    reward = _reward(theta, continuous=p.continuous_reward, nonlinear=p.nonlinear_reward,
                     decay=p.reward_decay_rate, span=p.theta_threshold_radians,
                     lo=p.reward_angle_range[0], hi=p.reward_angle_range[1])
    current_step = np.asarray(current_step) + 1
    truncated = current_step >= p.episode_length_limit
    done = done & ~truncated  # cart_pole.py:163-164
    return dict(state=np.stack([x, x_dot, theta, theta_dot], axis=-1), reward=np.asarray(reward, np.float64),
                terminated=done, truncated=truncated, x_acc=x_acc, theta_acc=theta_acc, current_step=current_step)


class PendulumDiskParams:
    def __init__(self, tau=0.005, continuous_reward=False, episode_length_limit=1000, angle_termination=None,
                 nonlinear_reward=False, reward_decay_rate=3.0, reward_angle_range=(-0.2, 0.2)):
        self.tau = tau
        self.continuous_reward = continuous_reward
        self.episode_length_limit = episode_length_limit
        self.angle_termination = angle_termination
        self.nonlinear_reward = nonlinear_reward
        self.reward_decay_rate = reward_decay_rate
        self.reward_angle_range = reward_angle_range
        # pendulum_disk.py:58-64
        self.J = 1.91e-4
        self.m = 0.055
        self.g = 9.81
        self.l = 0.042
        self.b = 3e-6
        self.K = 0.0536
        self.R = 9.5


def pendulum_disk_step(p: PendulumDiskParams, state, u, current_step, nep50=False):
    """One reference PendulumDisk.step for N envs.  state float64[N,2]; u = action[0] (float32 values)."""
    state = np.asarray(state, np.float64)
    alpha, alpha_dot = state[..., 0], state[..., 1]
    if nep50:
        drive = (np.float32(p.K / p.R) * np.asarray(u, np.float32)).astype(np.float64)
    else:
        drive = p.K / p.R * np.asarray(u, np.float64)
    acc = (p.m * p.g * p.l * np.sin(alpha) - p.b * alpha_dot - (p.K ** 2 / p.R) * alpha_dot + drive) / p.J
    alpha = normalize_angle(alpha + p.tau * alpha_dot)
    alpha_dot = alpha_dot + p.tau * acc
    done = (alpha < -PI) | (alpha > PI) | (alpha_dot < -15 * PI) | (alpha_dot > 15 * PI)
    if p.angle_termination is not None:
        done = done | (np.abs(alpha) > p.angle_termination)
    reward = _reward(alpha, continuous=p.continuous_reward, nonlinear=p.nonlinear_reward,
                     decay=p.reward_decay_rate, span=PI,
                     lo=p.reward_angle_range[0], hi=p.reward_angle_range[1])
    current_step = np.asarray(current_step) + 1
    truncated = current_step >= p.episode_length_limit
    done = done & ~truncated  # pendulum_disk.py:156-157
    return dict(state=np.stack([alpha, alpha_dot], axis=-1), reward=np.asarray(reward, np.float64),
                terminated=done, truncated=truncated, alpha_dot_dot=acc, current_step=current_step)
