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


def test_cold_inputs_option_is_validated_before_cuda():
    with pytest.raises(ValueError, match="cold_inputs"):
        VecGridWorld(4, cold_inputs="yes")


def test_residency_rule_counts_live_env_sets_per_device():
    """vec_env._Residency (the EF_STEP_COLD_INPUTS rule): cold once the env sets alive on ONE device exceed its L2 together;
    sets on other devices do not count, dead sets drop out, and every change of the population bumps the generation the
    env sets cache their verdict against."""
    import gc
    from rl_envs_forge_b200.vec_env import _Residency

    class Dev:
        def __init__(self, index):
            self.index = index

        def __eq__(self, other):
            return self.index == other.index

        def __hash__(self):
            return hash(self.index)

    class FakeSet:
        def __init__(self, nbytes, dev):
            self._step_bytes, self.device, self._torch = nbytes, dev, None

    reg = _Residency()
    reg._l2 = {0: 100, 1: 100}                       # bytes of L2 per device (normally read from the device properties)
    a, b = FakeSet(60, Dev(0)), FakeSet(60, Dev(0))
    g0 = reg.generation
    reg.add(a)
    assert reg.generation == g0 + 1 and not reg.cold(a)
    reg.add(b)
    assert reg.cold(a) and reg.cold(b)                # 120 > 100
    other = FakeSet(90, Dev(1))
    reg.add(other)
    assert not reg.cold(other)                        # device 1 holds 90 <= 100
    del b
    gc.collect()
    assert not reg.cold(a)                            # the dead set no longer counts
    reg.bump()
    assert reg.generation == g0 + 4
