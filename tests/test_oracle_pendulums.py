"""Pins oracle/pendulums.py against golden vectors from the unmodified reference (float64, numpy 2.x => nep50)."""
import json

import numpy as np
import pytest

from oracle import pendulums as P
from tests._golden import golden_files, load

RTOL = 1e-13  # float64 restatement vs float64 reference: differences only from pow(x,2) vs x*x style last-bit effects


def _close(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=1e-15)


@pytest.mark.parametrize("path", golden_files("cartpole"), ids=lambda p: p.split("cartpole_")[-1][:-4])
def test_cartpole_matches_reference(path):
    g = load(path)
    prm = P.CartPoleParams(**json.loads(str(g["config"])))
    out = P.cartpole_step(prm, g["pre"], g["action"], g["step"] - 1)
    assert _close(out["state"], g["post"])
    assert _close(out["x_acc"], g["acc"][:, 0]) and _close(out["theta_acc"], g["acc"][:, 1])
    assert np.array_equal(out["reward"], g["reward"]) or _close(out["reward"], g["reward"])
    assert np.array_equal(out["terminated"], g["done"])
    assert np.array_equal(out["truncated"], g["trunc"])


@pytest.mark.parametrize("path", golden_files("pendulumdisk"), ids=lambda p: p.split("pendulumdisk_")[-1][:-4])
def test_pendulum_disk_matches_reference(path):
    g = load(path)
    prm = P.PendulumDiskParams(**json.loads(str(g["config"])))
    nep50 = int(str(g["numpy_version"]).split(".")[0]) >= 2
    out = P.pendulum_disk_step(prm, g["pre"], g["action"], g["step"] - 1, nep50=nep50)
    assert _close(out["state"], g["post"])
    assert _close(out["alpha_dot_dot"], g["acc"][:, 0])
    assert _close(out["reward"], g["reward"])
    assert np.array_equal(out["terminated"], g["done"])
    assert np.array_equal(out["truncated"], g["trunc"])
    # the pinned-numpy (float64) arithmetic differs from the container's NEP-50 value only at fp32 rounding level
    out64 = P.pendulum_disk_step(prm, g["pre"], g["action"], g["step"] - 1, nep50=False)
    assert np.allclose(out64["alpha_dot_dot"], g["acc"][:, 0], rtol=1e-6, atol=1e-4)


def test_survey_known_answers():
    """SURVEY.md 8(c) golden one-step values."""
    o = P.cartpole_step(P.CartPoleParams(), np.full((1, 4), 0.1), np.float32([1.0]), np.zeros(1, int))
    assert np.allclose(o["state"][0], [0.102, 0.11807538, 0.102, 0.1023734], rtol=0, atol=5e-9)
    assert o["reward"][0] == 0.0
    assert abs(o["x_acc"][0] - 0.9037691387312428) < 1e-15 and abs(o["theta_acc"][0] - 0.11867013847739322) < 1e-15
    o = P.pendulum_disk_step(P.PendulumDiskParams(), np.full((1, 2), 0.1), np.float32([1.0]), np.zeros(1, int), nep50=True)
    assert np.allclose(o["state"][0], [0.1005, 0.306123], rtol=0, atol=5e-7) and o["reward"][0] == 1.0
    assert abs(o["alpha_dot_dot"][0] - 41.22460059345029) < 1e-12


def test_normalize_angle_reference_values():
    """tests/envs/inverted_pendulum/cart_pole/test_cart_pole.py:151-156."""
    pi = np.pi
    got = P.normalize_angle(np.array([0, pi, -pi, 2 * pi, -2 * pi, 3 * pi, -3 * pi]))
    assert np.allclose(got, [0, pi, pi, 0, 0, pi, pi])


def test_reference_reward_and_flag_pins():
    """test_cart_pole.py:74-140 and test_pendulum_disk.py:78-133 re-expressed on the restatement."""
    z = np.zeros(1, int)
    a0 = np.float32([0.0])
    o = P.cartpole_step(P.CartPoleParams(continuous_reward=True), [[0, 0, 0.1, 0]], a0, z)
    assert o["reward"][0] == pytest.approx(1 - abs(o["state"][0, 2]) / 0.2, rel=1e-12)
    o = P.cartpole_step(P.CartPoleParams(continuous_reward=True, nonlinear_reward=True, reward_decay_rate=30.0),
                        [[0, 0, 0.1, 0]], a0, z)
    th = abs(o["state"][0, 2])
    assert o["reward"][0] == pytest.approx(1 - (1 - np.exp(-30 * th)) / (1 - np.exp(-30 * 0.2)), rel=1e-12)
    assert P.cartpole_step(P.CartPoleParams(), [[0, 0, 0.0, 0]], a0, z)["reward"][0] == 1.0
    assert P.cartpole_step(P.CartPoleParams(), [[0, 0, 0.2, 0]], a0, z)["reward"][0] == 0.0
    o = P.cartpole_step(P.CartPoleParams(angle_termination=0.1), [[0, 0, 0.11, 0]], a0, z)
    assert o["terminated"][0] and not o["truncated"][0]
    o = P.cartpole_step(P.CartPoleParams(episode_length_limit=10), [[0, 0, 0.3, 0]], a0, np.array([9]))
    assert o["truncated"][0] and not o["terminated"][0]          # truncation overrides done
    o = P.pendulum_disk_step(P.PendulumDiskParams(continuous_reward=True), [[0.1, 0]], a0, z)
    assert o["reward"][0] == pytest.approx(1 - abs(o["state"][0, 0]) / np.pi, rel=1e-12)
    assert P.pendulum_disk_step(P.PendulumDiskParams(), [[0.0, 0]], a0, z)["reward"][0] == 1.0
    assert P.pendulum_disk_step(P.PendulumDiskParams(), [[0.25, 0]], a0, z)["reward"][0] == 0.0
