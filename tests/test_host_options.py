"""Host-side argument validation of the optional modes (runs without a GPU: these checks happen before any CUDA call)."""
import pytest

from rl_envs_forge_b200 import VecGridWorld, VecKArmedBandit


def test_gridworld_wire_option_is_validated_before_cuda():
    with pytest.raises(ValueError, match="wire must be"):
        VecGridWorld(4, wire="tiny")
    with pytest.raises(ValueError, match="rows <= 256"):
        VecGridWorld(4, rows=300, cols=4, wire="compact")
    with pytest.raises(ValueError, match="philox"):
        VecGridWorld(4, wire="compact", rng="pcg64", seed=1)


def test_bandit_rng_option_is_validated_before_cuda():
    with pytest.raises(ValueError, match="rng must be"):
        VecKArmedBandit(4, rng="xorshift")
    with pytest.raises(ValueError, match="2.5 KB"):
        VecKArmedBandit(VecKArmedBandit.MT19937_MAX_ENVS + 1, rng="mt19937")
    with pytest.raises(ValueError, match="seed sequence"):
        VecKArmedBandit(4, seed=[1, 2, 3], rng="mt19937")
