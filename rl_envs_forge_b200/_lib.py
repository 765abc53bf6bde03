"""ctypes binding of the C-ABI library (include/envforge_b200.h).

There is deliberately NO fallback: if libenvforge_b200.so is missing or fails to load, importing an env class raises.
The library is built in-tree by `python -m rl_envs_forge_b200._build` (also run by __graft_entry__.build()).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ENVFORGE_LIB selects an alternative build of the same library (kernel tuning experiments); default: the in-tree build
LIB_PATH = os.environ.get("ENVFORGE_LIB") or os.path.join(_HERE, "csrc", "libenvforge_b200.so")

EF_STATS_LEN = 8
EF_STATS_REPLICAS = 32
EF_STATS_STRIDE = 16
EF_ERR_LEN = 4
EF_STATS_MAX_SETS = 32
EF_PEER_HANDLE_BYTES = 64
EF_ABI_VERSION = 2
EF_WIRE_U8 = 0x1
EF_STEP_COLD_INPUTS = 0x2
EF_CLOCK_SHIFT = 20

EF_GRID_DONE = 0x1
EF_GRID_NO_OUTCOMES = 0x2
EF_GRID_BAD_PROBS = 0x4

STEP_ERROR_NAMES = {1: "bad action", 2: "no outcomes defined", 3: "probabilities do not sum to 1"}

ARM_FAMILIES = {"normal": 0, "uniform": 1, "exponential": 2, "gamma": 3, "beta": 4, "logistic": 5, "weibull": 6,
                "lognormal": 7, "default_normal": 8}


class GridDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("n_slots", C.c_int32), ("idx_cap", C.c_int32),
                ("thr", C.c_uint32 * 3), ("start_cell", C.c_int32), ("episode_limit", C.c_int32),
                ("table_cells", C.c_int32), ("table", C.c_void_p)]


class GridStepArgs(C.Structure):
    _fields_ = [("grid", C.POINTER(GridDesc)), ("pos", C.c_void_p), ("ep_len", C.c_void_p), ("ep_ret", C.c_void_p),
                ("action", C.c_void_p), ("draw", C.c_void_p), ("seed", C.c_uint64), ("t", C.c_uint64),
                ("env_offset", C.c_uint64), ("obs", C.c_void_p), ("reward", C.c_void_p), ("terminated", C.c_void_p),
                ("truncated", C.c_void_p), ("final_obs", C.c_void_p), ("stats", C.c_void_p), ("err", C.c_void_p),
                ("clock", C.c_void_p), ("n", C.c_int64), ("autoreset", C.c_int32), ("wire", C.c_int32),
                ("stream", C.c_void_p)]


class PendulumParams(C.Structure):
    _fields_ = [("tau", C.c_float), ("c", C.c_float * 8), ("angle_threshold", C.c_float),
                ("angle_termination", C.c_float), ("reward_mode", C.c_int32), ("reward_lo", C.c_float),
                ("reward_hi", C.c_float), ("reward_decay", C.c_float), ("reward_inv_span", C.c_float),
                ("reward_inv_norm", C.c_float), ("episode_limit", C.c_int32), ("fixed_init", C.c_int32),
                ("init_lo", C.c_float), ("init_hi", C.c_float), ("init_state", C.c_float * 4)]


class GamblerParams(C.Structure):
    _fields_ = [("goal_amount", C.c_int32), ("start_capital", C.c_int32), ("episode_limit", C.c_int32),
                ("reserved", C.c_int32), ("win_threshold", C.c_uint64)]


class CarRentalParams(C.Structure):
    _fields_ = [("max_cars", C.c_int32), ("max_move_cars", C.c_int32), ("rental_credit", C.c_int32),
                ("move_car_cost", C.c_int32), ("episode_limit", C.c_int32), ("init_equal", C.c_int32),
                ("lam", C.c_double * 4), ("exp_neg_lam", C.c_double * 4), ("log_lam", C.c_double * 4),
                ("ptrs_a", C.c_double * 4), ("ptrs_b", C.c_double * 4), ("ptrs_log_invalpha", C.c_double * 4),
                ("ptrs_vr", C.c_double * 4)]


class LabyrinthDesc(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("n_mazes", C.c_int32), ("vision_range", C.c_int32),
                ("episode_limit", C.c_int32), ("neutral_reward", C.c_float), ("wall_collision_reward", C.c_float),
                ("target_reached_reward", C.c_float), ("grids", C.c_void_p), ("maze_index", C.c_void_p),
                ("start_cell", C.c_void_p), ("target_cell", C.c_void_p)]


_P, _I32, _I64, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64

# name -> (restype, argtypes); mirrors include/envforge_b200.h one to one (tests check the symbol list against the header)
SIGNATURES = {
    "ef_abi_version": (C.c_int, []),
    "ef_status_string": (C.c_char_p, [C.c_int]),
    "ef_last_cuda_error": (C.c_int, []),
    "ef_last_cuda_error_string": (C.c_char_p, []),
    "ef_device_info": (C.c_int, [_P, _P, _P]),
    "ef_gridworld_step": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P, _P,
                                    _P, _P, _I64, _I32, _P]),
    "ef_gridworld_step_v": (C.c_int, [C.POINTER(GridStepArgs)]),
    "ef_gridworld_step_u8": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P, _P,
                                       _P, _P, _I64, _I32, _P]),
    "ef_gridworld_step_layouts": (C.c_int, [C.POINTER(GridDesc), _I32, _P, _P, _P, _P, _P, _P, _P, _U64, _U64, _U64, _P, _P,
                                            _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "ef_gridworld_rollout_layouts": (C.c_int, [C.POINTER(GridDesc), _I32, _P, _P, _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P,
                                               _P, _P, _P, _P, _I64, _P]),
    "ef_gridworld_reset_layouts": (C.c_int, [_I32, _P, _P, _I32, _P, _P, _P, _P, _P, _U64, _I64, _P]),
    "ef_gridworld_step_pcg64": (C.c_int, [C.POINTER(GridDesc), C.POINTER(C.c_uint64), _I32, _P, _P, _P, _P, _P, _P, _U64, _P,
                                          _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
    "ef_gridworld_rollout": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P,
                                       _P, _I64, _P]),
    "ef_gridworld_rollout_traj": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P, _P,
                                            _P, _I64, _P]),
    "ef_gridworld_reset": (C.c_int, [_I32, _I32, _P, _P, _P, _P, _P, _I64, _P]),
    "ef_gridworld_serve_chunk_envs": (C.c_int, [_I64]),
    "ef_gridworld_serve": (C.c_int, [C.POINTER(GridDesc), _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                     _P, _I64, _P]),
    "ef_policy_demo_serve": (C.c_int, [_P, _P, _P, _P, _I32, _P, _I64, _P]),
    "ef_policy_demo_step": (C.c_int, [_P, _P, _I32, _I64, _P]),
    "ef_cartpole_step": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P, _P,
                                   _P, _P, _P, _I64, _I32, _P]),
    "ef_cartpole_rollout": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P,
                                      _P, _P, _I64, _P]),
    "ef_cartpole_rollout_traj": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P,
                                           _P, _P, _P, _I64, _P]),
    "ef_pendulum_disk_rollout_traj": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P,
                                                _P, _P, _P, _P, _I64, _P]),
    "ef_cartpole_reset": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _U64, _U64, _U64, _P, _I64, _P]),
    "ef_pendulum_disk_step": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P,
                                        _P, _P, _P, _P, _I64, _I32, _P]),
    "ef_pendulum_disk_rollout": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P,
                                           _P, _P, _P, _I64, _P]),
    "ef_pendulum_disk_reset": (C.c_int, [C.POINTER(PendulumParams), _P, _P, _P, _U64, _U64, _U64, _P, _I64, _P]),
    "ef_bandit_step": (C.c_int, [_P, _I32, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P, _I64, _P]),
    "ef_bandit_step_means": (C.c_int, [_P, _I32, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P, _I64, _P]),
    "ef_bandit_step_mt19937": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P, _U64, _P, _P, _P, _P, _P, _I64, _P]),
    "ef_bandit_default_means": (C.c_int, [_I32, _U64, _U64, _P, _I64, _P]),
    "ef_episode_stat_reduce": (C.c_int, [_P, _P, _P, _I64, _P]),
    "ef_stats_fold": (C.c_int, [C.POINTER(C.c_void_p), _I32, _P, _P]),
    "ef_peer_comm_create": (C.c_int, [_I32, _I32, C.POINTER(C.c_void_p), _P]),
    "ef_peer_comm_connect": (C.c_int, [_P, _P]),
    "ef_peer_comm_destroy": (C.c_int, [_P]),
    "ef_stats_fold_allreduce": (C.c_int, [_P, C.POINTER(C.c_void_p), _I32, _P, _P, _P, _P]),
    "ef_gambler_step": (C.c_int, [C.POINTER(GamblerParams), _P, _P, _P, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P, _P,
                                  _P, _P, _I64, _I32, _P]),
    "ef_gambler_rollout": (C.c_int, [C.POINTER(GamblerParams), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P, _P,
                                     _I64, _P]),
    "ef_car_rental_rollout": (C.c_int, [C.POINTER(CarRentalParams), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P,
                                        _I64, _P]),
    "ef_labyrinth_rollout": (C.c_int, [C.POINTER(LabyrinthDesc), _P, _P, _P, _P, _U64, _U64, _U64, _I32, _P, _P, _P, _P, _P,
                                       _I64, _P]),
    "ef_gambler_reset": (C.c_int, [_I32, _P, _P, _P, _P, _P, _I64, _P]),
    "ef_car_rental_step": (C.c_int, [C.POINTER(CarRentalParams), _P, _P, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P,
                                     _P, _P, _I64, _I32, _P]),
    "ef_car_rental_reset": (C.c_int, [C.POINTER(CarRentalParams), _P, _P, _P, _U64, _U64, _U64, _P, _P, _I64, _P]),
    "ef_labyrinth_step": (C.c_int, [C.POINTER(LabyrinthDesc), _P, _P, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P, _P, _P,
                                    _P, _I64, _I32, _P]),
    "ef_labyrinth_observe": (C.c_int, [C.POINTER(LabyrinthDesc), _P, _U64, _P, _I64, _P]),
    "ef_labyrinth_reset": (C.c_int, [C.POINTER(LabyrinthDesc), _P, _P, _P, _U64, _P, _I64, _P]),
    "ef_host_pipe_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "ef_host_pipe_destroy": (C.c_int, [_P]),
    "ef_gridworld_step_host": (C.c_int, [_P, C.POINTER(GridDesc), _P, _P, _P, _P, _P, _U64, _U64, _U64] + [_P] * 10
                               + [_P, _P, _I64, _I32, _I32, _P]),
    "ef_gridworld_step_u8_host": (C.c_int, [_P, C.POINTER(GridDesc), _P, _P, _P, _P, _P, _U64, _U64, _U64] + [_P] * 10
                                  + [_P, _P, _I64, _I32, _I32, _P]),
    "ef_cartpole_step_host": (C.c_int, [_P, C.POINTER(PendulumParams), _P, _P, _P, _P, _P, _U64, _U64, _U64] + [_P] * 12
                              + [_P, _P, _I64, _I32, _I32, _P]),
    "ef_pendulum_disk_step_host": (C.c_int, [_P, C.POINTER(PendulumParams), _P, _P, _P, _P, _P, _U64, _U64, _U64]
                                   + [_P] * 12 + [_P, _P, _I64, _I32, _I32, _P]),
    "ef_bandit_step_host": (C.c_int, [_P, _P, _I32, _P, _P, _U64, _U64, _U64, _P, _P, _P, _P, _P, _P, _I64, _I32, _P]),
}

_lib = None


class EnvForgeError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library.  Raises EnvForgeError when it is absent -- no fallback path exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EnvForgeError(
            f"{LIB_PATH} not found. Build it with `python -m rl_envs_forge_b200._build` (needs nvcc). "
            "rl_envs_forge_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = lib.ef_abi_version()
    if got != EF_ABI_VERSION:
        raise EnvForgeError(f"libenvforge_b200.so ABI version {got} != expected {EF_ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(status):
    if status == 0:
        return
    lib = load()
    name = lib.ef_status_string(status).decode()
    if status == 2:
        raise EnvForgeError(f"{name}: {lib.ef_last_cuda_error_string().decode()} ({lib.ef_last_cuda_error()})")
    raise EnvForgeError(name)


def ptr(t):
    """Device (or None) pointer of a torch tensor as an int for ctypes c_void_p."""
    return None if t is None else t.data_ptr()
