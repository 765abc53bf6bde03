"""VecLabyrinth -- N players walking M host-generated mazes, stepped and rendered by one sm_100a kernel launch.

Drop-in surface for the STEPPING path of the reference class `Labyrinth` (rl_envs_forge/envs/labyrinth/labyrinth.py):
`step(action)` :245-288, `state` :195-243 (full matrix or `state_vision_range` window), `set_state(player_position)`
:311-331, `reward_schema` :103-108, spaces :88-101, `Action` (constants.py:4-8).

What stays on the host: maze GENERATION (maze/maze.py, room.py, corridor.py).  A maze enters as the arrays the reference's
`Labyrinth.maze` exposes once built -- `grid`, `start_position`, `target_position` -- e.g.
    VecLabyrinth.from_mazes([Labyrinth(30, 30, seed=s).maze for s in seeds], num_envs=1 << 20)
or as plain arrays (`grids` uint8 [M, rows, cols], `start_positions` / `target_positions` [M, 2]).  Consequently `reset()`
puts every player back on the start cell of ITS maze (the reference's `reset(same_seed=True)`); a fresh maze is installed
with `set_mazes(...)`.
"""
import numpy as np

from . import _lib, spaces
from .grid_world import Action  # same encoding: UP=0, RIGHT=1, DOWN=2, LEFT=3  # noqa: F401
from .vec_env import VecEnv, resolve_seed

WALL, PATH, TARGET, START, PLAYER = 0, 1, 2, 3, 4  # labyrinth/constants.py:12-16

DEFAULT_REWARD_SCHEMA = {"neutral_reward": -0.01, "wall_collision_reward": -1, "target_reached_reward": 10}


class VecLabyrinth(VecEnv):
    _HOST_TRANSPORT = "staged"

    def __init__(self, num_envs=1, grids=None, start_positions=None, target_positions=None, maze_index=None,
                 state_vision_range=None, reward_schema=None, *, episode_length_limit=None, device=None, seed=None,
                 env_offset=0, autoreset=True, strict=False):
        if state_vision_range is not None:
            assert isinstance(state_vision_range, (int, np.integer)) and state_vision_range > 0, \
                "state_vision_range must be a positive integer or None"      # labyrinth.py:95-97
        super().__init__(num_envs, device, seed, env_offset, autoreset)
        self.state_vision_range = None if state_vision_range is None else int(state_vision_range)
        self.reward_schema = dict(reward_schema) if reward_schema else dict(DEFAULT_REWARD_SCHEMA)
        self.episode_length_limit = episode_length_limit
        self.strict = bool(strict)
        self.single_action_space = spaces.Discrete(4)
        self.action_space = spaces.Batched(self.single_action_space, num_envs)
        self.pos = self._torch.zeros(self.num_envs, dtype=self._torch.int32, device=self.device)
        self._final_obs = None
        self.set_mazes(grids, start_positions, target_positions, maze_index)

    @classmethod
    def from_mazes(cls, mazes, num_envs=None, **kw):
        """Build from objects exposing `.grid`, `.start_position`, `.target_position` (the reference's `Maze`)."""
        grids = np.stack([np.asarray(m.grid, np.uint8) for m in mazes])
        starts = np.array([tuple(m.start_position) for m in mazes])
        targets = np.array([tuple(m.target_position) for m in mazes])
        return cls(len(mazes) if num_envs is None else num_envs, grids, starts, targets, **kw)

    # -- mazes ----------------------------------------------------------------------------------------------------
    def set_mazes(self, grids, start_positions, target_positions, maze_index=None):
        """Install M mazes (and reset every player to its start cell).  grids uint8 [M, rows, cols] (or one [rows, cols]);
        positions [M, 2] (row, col).  maze_index int [N] assigns a maze to each player; default: (env_offset + i) % M."""
        torch = self._torch
        if grids is None or start_positions is None or target_positions is None:
            raise ValueError("grids, start_positions and target_positions are required (maze generation is host work: "
                             "pass the arrays of the reference's Labyrinth.maze)")
        g = np.asarray(grids)
        if g.ndim == 2:
            g = g[None]
        if g.ndim != 3:
            raise ValueError("grids must have shape [M, rows, cols]")
        g = np.ascontiguousarray(g.astype(np.uint8))
        m, rows, cols = g.shape
        st = np.asarray(start_positions, np.int64).reshape(-1, 2)
        tg = np.asarray(target_positions, np.int64).reshape(-1, 2)
        if st.shape[0] != m or tg.shape[0] != m:
            raise ValueError(f"start_positions / target_positions must hold {m} (row, col) pairs")
        for name, p in (("start", st), ("target", tg)):
            if (p < 0).any() or (p[:, 0] >= rows).any() or (p[:, 1] >= cols).any():
                raise ValueError(f"{name} position outside the maze")
        self.rows, self.cols, self.n_mazes = rows, cols, m
        self.grids = g
        self.start_positions, self.target_positions = st, tg
        dev = self.device
        self._grids = torch.from_numpy(g.reshape(m, rows * cols)).to(dev)
        self._start = torch.from_numpy((st[:, 0] * cols + st[:, 1]).astype(np.int32)).to(dev)
        self._target = torch.from_numpy((tg[:, 0] * cols + tg[:, 1]).astype(np.int32)).to(dev)
        self._maze_index = None
        if maze_index is not None:
            mi = np.asarray(maze_index, np.int64).reshape(-1)
            if mi.shape[0] != self.num_envs or (mi < 0).any() or (mi >= m).any():
                raise ValueError(f"maze_index must hold {self.num_envs} values in 0..{m - 1}")
            self._maze_index = torch.from_numpy(mi.astype(np.int32)).to(dev)
        v = self.state_vision_range
        self._obs_shape = (rows, cols) if v is None else (2 * v + 1, 2 * v + 1)
        self.single_observation_space = spaces.Box(low=0, high=4, shape=self._obs_shape, dtype=np.uint8)
        self.observation_space = spaces.Batched(self.single_observation_space, self.num_envs)
        n = self.num_envs
        self._obs = torch.zeros((n,) + self._obs_shape, dtype=torch.uint8, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._term = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._final_obs = None
        d = _lib.LabyrinthDesc()
        d.rows, d.cols, d.n_mazes = rows, cols, m
        d.vision_range = 0 if v is None else v
        d.episode_limit = int(self.episode_length_limit) if self.episode_length_limit is not None else 0
        d.neutral_reward = float(self.reward_schema["neutral_reward"])
        d.wall_collision_reward = float(self.reward_schema["wall_collision_reward"])
        d.target_reached_reward = float(self.reward_schema["target_reached_reward"])
        d.grids, d.start_cell, d.target_cell = self._grids.data_ptr(), self._start.data_ptr(), self._target.data_ptr()
        d.maze_index = None if self._maze_index is None else self._maze_index.data_ptr()
        self._desc = d
        self.reset()

    def maze_of_env(self):
        """int64 [N] maze index of every player (host array)."""
        if self._maze_index is not None:
            return self._maze_index.cpu().numpy().astype(np.int64)
        return (self.env_offset + np.arange(self.num_envs, dtype=np.int64)) % self.n_mazes

    # -- reference-style attributes -----------------------------------------------------------------------------------
    @property
    def player_position(self):
        """int32 [N,2] (row, col) of every player."""
        p = self.pos.to(self._torch.int64)
        return self._torch.stack([p // self.cols, p % self.cols], dim=1).to(self._torch.int32)

    @property
    def state(self):
        """Labyrinth.state (:195-208): uint8 [N, rows, cols] or [N, 2v+1, 2v+1], rendered from the current positions."""
        self._call(self._lib.ef_labyrinth_observe, self._desc, _lib.ptr(self.pos), self.env_offset, _lib.ptr(self._obs),
                   self.num_envs)
        return self._obs

    def set_state(self, player_position):
        """Labyrinth.set_state (:311-331): place the player(s); one (row, col) for all or [N,2].  Raises ValueError for a
        wall cell or the target cell, like the reference.  A dict from get_state() restores a snapshot instead."""
        torch = self._torch
        if isinstance(player_position, dict):
            snap = torch.as_tensor(player_position["pos"], device=self.device)
            if snap.numel() and (bool((snap < 0).any()) or bool((snap >= self.rows * self.cols).any())):
                raise ValueError(f"player cells must lie in 0..{self.rows * self.cols - 1}")   # unchecked on the device
            self.pos.copy_(snap)
            self.ep_len.copy_(player_position["ep_len"])
            self.ep_ret.copy_(player_position["ep_ret"])
            self._t = int(player_position.get("t", self._t))
            return
        p = np.asarray(player_position.cpu() if torch.is_tensor(player_position) else player_position, np.int64).reshape(-1, 2)
        if p.shape[0] == 1:
            p = np.repeat(p, self.num_envs, axis=0)
        if p.shape[0] != self.num_envs:
            raise ValueError(f"player_position must be one (row, col) or [{self.num_envs}, 2]")
        if (p < 0).any() or (p[:, 0] >= self.rows).any() or (p[:, 1] >= self.cols).any():
            raise IndexError("player position outside the maze")
        mi = self.maze_of_env()
        if (self.grids[mi, p[:, 0], p[:, 1]] == WALL).any():
            raise ValueError("Invalid position, the player can't be on a wall.")
        if (p == self.target_positions[mi]).all(axis=1).any():
            raise ValueError("Invalid position, can't place the player on the target position.")
        self.pos.copy_(torch.from_numpy((p[:, 0] * self.cols + p[:, 1]).astype(np.int32)))

    def get_state(self):
        return {"pos": self.pos.clone(), "ep_len": self.ep_len.clone(), "ep_ret": self.ep_ret.clone(), "t": self._t}

    # -- hot path ---------------------------------------------------------------------------------------------------
    def reset(self, seed=None, same_seed=True, *, mask=None):
        """Players back to the start cell of their maze (all, or those selected by `mask`); returns (state, {}).  The mazes
        are kept (see the module docstring); `seed` re-keys the random-policy stream only."""
        if seed is not None:
            self.seed, self._seed = seed, resolve_seed(seed)
        m = None if mask is None else self._torch.as_tensor(mask, device=self.device).to(self._torch.uint8).contiguous()
        self._call(self._lib.ef_labyrinth_reset, self._desc, _lib.ptr(self.pos), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret),
                   self.env_offset, _lib.ptr(m), self.num_envs)
        self._reset_epoch += 1
        return self.state, {}

    def step(self, actions, *, return_final_obs=False, return_obs=True):
        """Labyrinth.step (:245-288) for all players.  actions: int [N] in {0,1,2,3} (CUDA tensor, or a host array for the
        numpy-in / numpy-out path).  `return_obs=False` skips rendering the observation (rows*cols bytes per player).
        Returns (state uint8 [N, ...], reward fp32 [N], terminated bool [N], truncated bool [N], info)."""
        torch, n = self._torch, self.num_envs
        host_io = not (torch.is_tensor(actions) and actions.is_cuda)
        if host_io:
            a = actions.reshape(-1).to(torch.int32) if torch.is_tensor(actions) else \
                np.asarray(actions).reshape(-1).astype(np.int32, copy=False)
            act = self._to_device("action", self._host_in("action", a, torch.int32, (n,)), torch.int32, (n,))
        else:
            if actions.numel() != n:
                raise ValueError(f"actions must hold {n} values, got {tuple(actions.shape)}")
            act = actions.reshape(-1)
            act = (act if act.dtype == torch.int32 else act.to(torch.int32)).contiguous()
        fobs = None
        if return_final_obs:
            if self._final_obs is None:
                self._final_obs = torch.zeros((n,) + self._obs_shape, dtype=torch.uint8, device=self.device)
            fobs = self._final_obs
        obs = self._obs if return_obs else None
        self._call(self._lib.ef_labyrinth_step, self._desc, _lib.ptr(self.pos), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret),
                   _lib.ptr(act), self._seed, self._t_arg(), self.env_offset, _lib.ptr(obs), _lib.ptr(self._reward),
                   _lib.ptr(self._term), _lib.ptr(self._trunc), _lib.ptr(fobs), _lib.ptr(self._stats_raw), _lib.ptr(self.err),
                   _lib.ptr(self._clock), n, int(self.autoreset))
        self._t += 1
        if self.strict:
            self.check_errors()
        info = {}
        if host_io:
            named = ([("obs", obs)] if return_obs else []) + [("reward", self._reward), ("term", self._term),
                                                              ("trunc", self._trunc)]
            out = self._to_host(named + ([("final_obs", fobs)] if return_final_obs else []))
            if not return_obs:
                out = [None] + out
            if return_final_obs:
                info["final_observation"] = out[4]
            return out[0], out[1], out[2].view(np.bool_), out[3].view(np.bool_), info
        if return_final_obs:
            info["final_observation"] = fobs
        return obs, self._reward, self._term.view(torch.bool), self._trunc.view(torch.bool), info

    def rollout(self, K, actions=None, *, return_rewards=False):
        """K fused steps in one kernel (positions in registers, no observation rendering -- read `state` afterwards).
        actions: CUDA int32 [K,N] or None (uniform random direction).  Returns (rewards fp32 [K,N], flags uint8 [K,N]:
        bit0 terminated, bit1 truncated) if return_rewards else None; auto-reset to the maze's start cell is always on."""
        torch, n = self._torch, self.num_envs
        if actions is not None:
            if not (torch.is_tensor(actions) and actions.is_cuda and actions.dtype == torch.int32
                    and tuple(actions.shape) == (K, n)):
                raise ValueError(f"actions must be a CUDA int32 tensor of shape ({K}, {n})")
            actions = actions.contiguous()
        rew = flg = None
        if return_rewards:
            rew = torch.empty((K, n), dtype=torch.float32, device=self.device)
            flg = torch.empty((K, n), dtype=torch.uint8, device=self.device)
        self._call(self._lib.ef_labyrinth_rollout, self._desc, _lib.ptr(self.pos), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret),
                   _lib.ptr(actions), self._seed, self._t_arg(), self.env_offset, int(K), _lib.ptr(rew), _lib.ptr(flg),
                   _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), n)
        self._t += int(K)
        return (rew, flg) if return_rewards else None

    def _describe_error(self, count, env, detail, kind_id, kind):
        return f"{detail} is not a valid Action (env {env}; {count} rejected env-step(s))"
