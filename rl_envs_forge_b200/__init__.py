"""rl_envs_forge_b200 -- B200-native batched stepping for the hot path of mariusdgm/rl-envs-forge.

    from rl_envs_forge_b200 import VecGridWorld, VecKArmedBandit, VecCartPole, VecPendulumDisk

Each class mirrors the constructor kwargs / reset / step surface of the reference class of the same name (without the
`Vec` prefix) for `num_envs` independent instances whose state lives in device memory; transitions run in hand-written
sm_100a CUDA kernels behind the C ABI declared in include/envforge_b200.h.  There is no CPU fallback.
"""
from ._lib import EnvForgeError, LIB_PATH  # noqa: F401
from .grid_world import Action, GridSpec, VecGridWorld, VecGridWorldLayouts  # noqa: F401
from .k_armed_bandit import Arm, VecKArmedBandit  # noqa: F401
from .cart_pole import VecCartPole  # noqa: F401
from .pendulum_disk import VecPendulumDisk  # noqa: F401
from .acml import VecGamblersProblem, VecJacksCarRental  # noqa: F401
from .labyrinth import VecLabyrinth  # noqa: F401
from .vec_env import all_reduce_stats, shard_range, stats_to_dict  # noqa: F401
from .stats import StatsReducer, fold_stats  # noqa: F401

__version__ = "0.1.0"
