"""VecGridWorld -- N GridWorld instances stepped by one sm_100a kernel launch.

Drop-in surface for the reference class `GridWorld` (rl_envs_forge/envs/grid_world/grid_world.py): same constructor
kwargs and defaults (:19-34), same validation errors (validate_args :101-212), `reset` / `step` with batched
arguments, `action_space` / `observation_space`, `state`, `start_state`, `episode_length_counter`, `mdp`, `P`,
`add_special_transition`.  What changes: state lives in SoA device buffers (cell index int32, episode length int32,
episode return fp32), and the dict MDP is compiled into a flat outcome table that the kernels keep in shared memory.

Host-only pieces (`Action`, `validate_args`, `GridSpec`) import without CUDA; `VecGridWorld` needs a B200.
"""
import math
import os
from enum import IntEnum

import numpy as np

from . import _lib, spaces
from ._lib import EnvForgeError
from .vec_env import RESIDENCY, VecEnv, resolve_seed


class Action(IntEnum):  # grid_world.py:11-15
    UP = 0
    RIGHT = 1
    DOWN = 2
    LEFT = 3


_SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)  # numpy Generator.choice tolerance on sum(p)
_DEFAULT_REWARDS = {"valid_move": -0.1, "wall_collision": -1, "out_of_bounds": -1, "default": 0.0}
_DEFAULT_TERMINALS = {(3, 4): 1.0}  # grid_world.py:24
# outcome slot j >= 1 of intended action a is the j-th action of the enum order with a removed (grid_world.py:348-357)
_SLIP_ORDER = np.array([[a] + [b for b in range(4) if b != a] for a in range(4)])


def _is_cell(x):
    return isinstance(x, tuple) and len(x) == 2 and all(isinstance(v, int) for v in x)


def validate_args(rows, cols, start_state, terminal_states, walls, special_transitions, rewards, seed,
                  slip_distribution, p_success, episode_length_limit):
    """Same acceptance rules and messages as the reference's GridWorld.validate_args (grid_world.py:101-212)."""
    if not isinstance(rows, int) or rows <= 0:
        raise ValueError("rows must be a positive integer")
    if not isinstance(cols, int) or cols <= 0:
        raise ValueError("cols must be a positive integer")
    if not _is_cell(start_state):
        raise ValueError("start_state must be a tuple of two integers")
    if not (isinstance(terminal_states, dict)
            and all(_is_cell(k) and isinstance(v, (int, float)) for k, v in terminal_states.items())):
        raise ValueError("terminal_states must be a dictionary with keys as tuples of two integers and values as "
                         "integers or floats")
    if walls is not None and not (isinstance(walls, set) and all(_is_cell(w) for w in walls)):
        raise ValueError("walls must be a set of tuples of two integers or an empty set")
    if special_transitions is not None and not (
            isinstance(special_transitions, dict)
            and all(isinstance(k, tuple) and len(k) == 2 and _is_cell(k[0]) and isinstance(k[1], Action)
                    and isinstance(v, tuple) and len(v) == 2 and _is_cell(v[0]) and isinstance(v[1], (int, float))
                    for k, v in special_transitions.items())):
        raise ValueError(
            "special_transitions must be a dictionary with keys as tuples where the first element is a tuple of two "
            "integers and the second element is an Action, and values as tuples where the first element is a tuple of "
            "two integers and the second element is an integer or float, or an empty dictionary")
    if rewards is not None and not (
            isinstance(rewards, dict)
            and all(isinstance(k, str) and isinstance(v, (int, float)) for k, v in rewards.items())):
        raise ValueError("rewards must be a dictionary with keys as strings and values as integers or floats, or an "
                         "empty dictionary")
    if seed is not None and not isinstance(seed, int):
        raise ValueError("seed must be an integer or None")
    if slip_distribution is not None and not (
            isinstance(slip_distribution, dict)
            and all(isinstance(k, Action) and isinstance(v, (int, float)) for k, v in slip_distribution.items())):
        raise ValueError("slip_distribution must be a dictionary with keys as Actions and values as integers or "
                         "floats, or an empty dictionary")
    if not isinstance(p_success, (int, float)) or not (0 <= p_success <= 1):
        raise ValueError("p_success must be a float between 0 and 1")
    if episode_length_limit is not None and not isinstance(episode_length_limit, int):
        raise ValueError("episode_length_limit must be an integer or None")


class GridSpec:
    """Host-side compiled form of one GridWorld configuration (shared by all N envs).

    Arrays are indexed by row = cell*4 + action and outcome slot (4 slots; single-outcome rows replicate slot 0):
        nxt int32 [R,4], reward float64 [R,4], done bool [R,4], count int32 [R] (0 = no MDP entry), bad bool [R]
    `thr`, `idx_cap`, `n_slots` implement numpy's Generator.choice rule with integer thresholds (see the header).
    """

    def __init__(self, rows=5, cols=5, start_state=(0, 0), terminal_states=_DEFAULT_TERMINALS, walls=None,
                 special_transitions=None, rewards=None, p_success=1.0, episode_length_limit=None):
        self.rows, self.cols = rows, cols
        self.start_state = start_state
        self.terminal_states = terminal_states
        self.walls = walls or set()
        self.special_transitions = special_transitions or {}
        self.rewards = rewards or dict(_DEFAULT_REWARDS)
        self.p_success = p_success
        self.episode_length_limit = episode_length_limit
        S = rows * cols
        if S > 65535:
            raise ValueError("rl_envs_forge_b200 supports at most 65535 cells per grid (16-bit cell index in the table)")
        self.n_cells = S
        self._construction_specials = dict(self.special_transitions)
        self._compile()
        self._construction_tables = (self.nxt.copy(), self.count.copy(), self.prob_row.copy())

    # -- compiler (vectorised restatement of build_mdp / calculate_outcomes / calculate_transition / calculate_reward) ----
    def _in_grid(self, cell):
        return 0 <= cell[0] < self.rows and 0 <= cell[1] < self.cols

    def _cell_index(self, cell):
        return cell[0] * self.cols + cell[1]

    def _compile(self):
        rows, cols, S = self.rows, self.cols, self.n_cells
        r, c = np.divmod(np.arange(S), cols)
        is_wall = np.zeros(S, bool)
        for w in self.walls:
            if self._in_grid(w):
                is_wall[self._cell_index(w)] = True
        term_reward = np.zeros(S, np.float64)
        is_term = np.zeros(S, bool)
        for t, rew in self.terminal_states.items():
            if self._in_grid(t):
                is_term[self._cell_index(t)] = True
                term_reward[self._cell_index(t)] = rew
        # default_transition (:379-399), wall => stay (:373-374)
        tr = np.stack([np.maximum(r - 1, 0), r, np.minimum(r + 1, rows - 1), r], axis=1)
        tc = np.stack([c, np.minimum(c + 1, cols - 1), c, np.maximum(c - 1, 0)], axis=1)
        t1 = tr * cols + tc                                   # [S,4] destination of (cell, action)
        here = np.arange(S)[:, None]
        t1 = np.where(is_wall[t1], here, t1)
        # calculate_reward (:467-493): terminal reward, else stay => out_of_bounds, else valid_move
        rew1 = np.where(is_term[t1], term_reward[t1],
                        np.where(t1 == here, float(self.rewards.get("out_of_bounds", -1)),
                                 float(self.rewards.get("valid_move", -0.1))))
        done1 = is_term[t1]
        # special transitions replace calculate_transition's answer (:365-369), hence also appear as slip outcomes
        special_rows = {}
        for (frm, act), (to, rew) in self.special_transitions.items():
            if not self._in_grid(frm):
                continue
            if not self._in_grid(to):
                raise ValueError(f"special transition target {to} is outside the {rows}x{cols} grid; the reference "
                                 "accepts it but can never step from there -- not representable on device")
            i = self._cell_index(frm)
            t1[i, int(act)] = self._cell_index(to)
            rew1[i, int(act)] = rew
            done1[i, int(act)] = to in self.terminal_states
            special_rows[i * 4 + int(act)] = True
        p = float(self.p_success)
        self.n_slots = 4 if p < 1.0 else 1
        R = S * 4
        # outcome slots: slot 0 intended, slots 1..3 the other actions in enum order (:327-359)
        order = _SLIP_ORDER[None, :, :]                                     # [1,4(action),4(slot)]
        nxt = np.take_along_axis(t1[:, None, :].repeat(4, 1), order, 2).reshape(R, 4)
        reward = np.take_along_axis(rew1[:, None, :].repeat(4, 1), order, 2).reshape(R, 4)
        done = np.take_along_axis(done1[:, None, :].repeat(4, 1), order, 2).reshape(R, 4)
        count = np.full(R, 4 if p < 1.0 else 1, np.int32)
        prob_row = np.tile(np.array([p] + [(1.0 - p) / 3.0] * 3), (R, 1))
        if p >= 1.0:
            prob_row[:, 1:] = 0.0
        bad = np.zeros(R, bool)
        cell_of_row = np.arange(R) // 4
        # walls have no entry (:247-252); terminals self-loop with reward 0 (:263-265)
        count[is_wall[cell_of_row]] = 0
        term_rows = is_term[cell_of_row]  # also for a cell listed as wall AND terminal: the loop at :263 ignores walls
        count[term_rows] = 1
        nxt[term_rows] = cell_of_row[term_rows][:, None]
        reward[term_rows] = 0.0
        done[term_rows] = True
        prob_row[term_rows] = [1.0, 0.0, 0.0, 0.0]
        # specials overwrite their row with one outcome of probability p_success (:268-274)
        for row in special_rows:
            a = row % 4
            i = row // 4
            count[row] = 1
            nxt[row] = t1[i, a]
            reward[row] = rew1[i, a]
            done[row] = done1[i, a]
            prob_row[row] = [p, 0.0, 0.0, 0.0]
            bad[row] = abs(p - 1.0) > _SQRT_EPS
        single = count == 1
        nxt[single] = nxt[single][:, :1]
        reward[single] = reward[single][:, :1]
        done[single] = done[single][:, :1]
        self.nxt, self.reward, self.done, self.count, self.bad, self.prob_row = nxt, reward, done, count, bad, prob_row
        self._set_thresholds()

    def _set_thresholds(self):
        p = float(self.p_success)
        if self.n_slots == 1:
            self.thr, self.idx_cap = (0xFFFFFFFF,) * 3, 0
            return
        cdf = np.cumsum(np.array([p] + [(1.0 - p) / 3.0] * 3, np.float64))
        cdf /= cdf[-1]
        t = [int(math.ceil(float(cdf[k]) * 2.0 ** 32)) for k in range(3)]   # exact: scaling by 2^32 is exact in binary64
        self.idx_cap = sum(1 for v in t if v < 2 ** 32)
        self.thr = tuple(min(v, 0xFFFFFFFF) for v in t)

    # -- runtime patch (:214-243) -----------------------------------------------------------------------------------
    def add_special_transition(self, from_state, action, to_state=None, reward=None):
        a = int(action)
        if to_state is None:  # calculate_next_state(check_special=False) (:432-465)
            i = self._cell_index(from_state)
            rr, cc = from_state
            cand = {0: (max(rr - 1, 0), cc), 1: (rr, min(cc + 1, self.cols - 1)),
                    2: (min(rr + 1, self.rows - 1), cc), 3: (rr, max(cc - 1, 0))}[a]
            to_state = from_state if (cand in self.walls or cand == from_state) else cand
        if reward is None:
            reward = self.rewards["default"]
        if not self._in_grid(from_state):
            return  # the reference stores an unreachable dict entry; nothing to compile
        if not self._in_grid(to_state):
            raise ValueError(f"special transition target {to_state} is outside the grid -- not representable on device")
        row = self._cell_index(from_state) * 4 + a
        self.count[row] = 1
        self.nxt[row] = self._cell_index(to_state)
        self.reward[row] = reward
        self.done[row] = to_state in self.terminal_states
        self.prob_row[row] = [self.p_success, 0.0, 0.0, 0.0]
        self.bad[row] = abs(float(self.p_success) - 1.0) > _SQRT_EPS

    # -- exports ------------------------------------------------------------------------------------------------
    def mdp_dict(self):
        """The reference's `mdp` attribute: {((r, c), Action): [(next_state, reward, done, prob), ...]}."""
        out = {}
        cols = self.cols
        for row in np.nonzero(self.count)[0]:
            cell, a = divmod(int(row), 4)
            key = ((cell // cols, cell % cols), Action(a))
            out[key] = [((int(self.nxt[row, k]) // cols, int(self.nxt[row, k]) % cols), float(self.reward[row, k]),
                         bool(self.done[row, k]), float(self.prob_row[row, k])) for k in range(int(self.count[row]))]
        return out

    def transition_probs(self, slip_distribution):
        """The reference's `transition_probs` attribute (default_transition_probs :401-430 over calculate_next_state
        :432-465): {((r, c), Action): {next_state: probability}} for every cell that is neither terminal nor wall.  A
        second, differently normalised model than `mdp` -- the intended move gets p_success, then EVERY action a' (the
        intended one included) adds (1 - p_success) * slip_distribution[a'] -- that `step` never reads; exported because
        tabular code written against the reference may.  Uses the construction-time special transitions only."""
        rows, cols = self.rows, self.cols
        walls = {w for w in self.walls}
        special = {(tuple(k[0]), int(k[1])): tuple(v[0]) for k, v in self._construction_specials.items()}
        move = {0: (-1, 0), 1: (0, 1), 2: (1, 0), 3: (0, -1)}

        def destination(cell, a):
            if (cell, a) in special:
                return special[(cell, a)]
            to = (min(max(cell[0] + move[a][0], 0), rows - 1), min(max(cell[1] + move[a][1], 0), cols - 1))
            return cell if (to in walls or to == cell) else to

        slip = [(int(a), float(w)) for a, w in slip_distribution.items()]   # the reference iterates the dict in its order
        out = {}
        for r in range(rows):
            for c in range(cols):
                cell = (r, c)
                if cell in self.terminal_states or cell in walls:
                    continue
                for a in Action:
                    acc = {destination(cell, int(a)): 0 + self.p_success}
                    for sa, w in slip:
                        to = destination(cell, sa)
                        acc[to] = acc.get(to, 0) + (1 - self.p_success) * w
                    out[(cell, a)] = acc
        return out

    def probability_matrix(self):
        """create_probability_matrix (:276-300) from the construction-time tables."""
        nxt, count, prob = self._construction_tables
        S = self.n_cells
        P = np.zeros((S, S, 4))
        rows_idx = np.arange(S * 4)
        cell, act = rows_idx // 4, rows_idx % 4
        absorbing = np.zeros(S, bool)
        for t in self.terminal_states:
            if self._in_grid(t):
                absorbing[self._cell_index(t)] = True
        for w in self.walls:
            if self._in_grid(w):
                absorbing[self._cell_index(w)] = True
        for k in range(4):
            live = (count > k) & ~absorbing[cell]
            np.add.at(P, (cell[live], nxt[live, k], act[live]), prob[live, k])
        idx = np.nonzero(absorbing)[0]
        P[idx, idx, :] = 1.0
        return P

    def device_table(self, extra_cells=0):
        """Packed ef_grid_outcome array (uint32 [entries, 2]): word0 = next | flags << 16, word1 = fp32 reward bits."""
        R, slots = self.n_cells * 4, self.n_slots
        rejected = (self.count == 0) | self.bad
        # rejected rows are compiled INERT (next = own cell, reward 0, not done): the branch-free kernel path applies
        # them and only masks the step counter; the flags let it report them (the reference raises on these rows)
        own = (np.arange(R) // 4)[:, None]
        nxt = np.where(rejected[:, None], own, self.nxt[:, :slots])
        reward = np.where(rejected[:, None], 0.0, self.reward[:, :slots])
        flags = (self.done[:, :slots] & ~rejected[:, None]).astype(np.uint32) * _lib.EF_GRID_DONE
        flags |= (self.count == 0)[:, None].astype(np.uint32) * _lib.EF_GRID_NO_OUTCOMES
        flags |= (self.bad & (self.count != 0))[:, None].astype(np.uint32) * _lib.EF_GRID_BAD_PROBS
        word0 = (nxt.astype(np.uint32) & 0xFFFF) | (flags << 16)
        word1 = reward.astype(np.float32).view(np.uint32)
        tab = np.stack([word0, word1], axis=-1).reshape(R * slots, 2)
        if extra_cells:
            pad = np.zeros((extra_cells * 4 * slots, 2), np.uint32)
            cells = self.n_cells + np.arange(extra_cells * 4 * slots) // (4 * slots)
            pad[:, 0] = (cells.astype(np.uint32) & 0xFFFF) | (_lib.EF_GRID_NO_OUTCOMES << 16)
            tab = np.concatenate([tab, pad])
        return np.ascontiguousarray(tab)


class _GridStateLayout:
    """State storage shared by VecGridWorld and VecGridWorldLayouts: `pos` / `ep_len` as two int32 arrays (split) or, when
    the episode counter provably stays below 2^16 (auto-reset on, 0 < episode_length_limit <= 65535), packed into one
    int32 word per env -- cell in bits 0-15, counter in bits 16-31: 8 fewer bytes through HBM per env-step.  Set
    `self._packed` before VecEnv.__init__ runs, then call `_alloc_grid_state()`."""

    @staticmethod
    def _can_pack(autoreset, episode_length_limit):
        return (bool(autoreset) and isinstance(episode_length_limit, (int, np.integer)) and 0 < episode_length_limit <= 65535)

    def _alloc_grid_state(self):
        torch = self._torch
        if self._packed:
            self._pl = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        else:
            self._pos_split = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)

    def _state_ptrs(self):
        """(pos, ep_len) arguments of the C ABI: the packed word and NULL, or the two split arrays."""
        if self._packed:
            return self._pl.data_ptr(), None
        return self._pos_split.data_ptr(), self._ep_len_split.data_ptr()

    @property
    def pos(self):
        """int32 [N] cell index (row * cols + col).  Split layout: the live tensor; packed layout: a computed copy --
        assign (`env.pos = t`) to write."""
        return (self._pl & 0xFFFF) if self._packed else self._pos_split

    @pos.setter
    def pos(self, value):
        v = self._torch.as_tensor(value, device=self.device).to(self._torch.int32)
        # the kernels index the outcome table with cell * 4 + action unchecked: a cell from another grid (a snapshot of a
        # different layout, a bad tensor) must be refused here, on this cold path, not fault on the device
        top = getattr(self, "_max_cell", None)
        if top is not None and v.numel() and (bool((v < 0).any()) or bool((v > top).any())):
            raise ValueError(f"cell indices must lie in 0..{top} (rows * cols cells plus the off-grid pseudo-cell)")
        if self._packed:
            self._pl.copy_((self._pl & -65536) | (v & 0xFFFF))
        else:
            self._pos_split.copy_(v)

    @property
    def ep_len(self):
        """int32 [N] episode_length_counter (see `pos` for the layout caveat)."""
        return ((self._pl >> 16) & 0xFFFF) if self.__dict__.get("_packed") else self._ep_len_split

    @ep_len.setter
    def ep_len(self, value):
        if "_ep_len_given" not in self.__dict__:   # first assignment: VecEnv.__init__ hands over the tensor it allocated
            self._ep_len_given = True
            if not self._packed:
                self._ep_len_split = value
            return
        v = self._torch.as_tensor(value, device=self.device).to(self._torch.int32)
        if self._packed:
            if bool((v < 0).any()) or bool((v > 65535).any()):
                raise ValueError("packed state layout holds episode_length_counter in 16 bits")
            self._pl.copy_((self._pl & 0xFFFF) | (v << 16))
        else:
            self._ep_len_split.copy_(v)

    @property
    def autoreset(self):
        return self._autoreset

    @autoreset.setter
    def autoreset(self, value):
        self._autoreset = bool(value)
        if not self._autoreset:
            self._leave_packed_layout()

    def _leave_packed_layout(self):
        """Needed once auto-reset is switched off: the episode counter is then unbounded."""
        if self._packed and "_pl" in self.__dict__:
            pos, ln = self.pos.clone(), self.ep_len.clone()
            self._packed = False
            self._pos_split, self._ep_len_split = pos, ln
            del self._pl
            self._fast_args = None


class VecGridWorld(_GridStateLayout, VecEnv):
    """N GridWorld instances sharing one configuration.  See module docstring; kwargs as in the reference."""
    _HOST_TRANSPORT = "staged"   # the one-pass persistent step kernel serialises its PCIe reads and writes: DMA wins

    def __init__(self, num_envs=1, rows=5, cols=5, start_state=(0, 0), terminal_states=_DEFAULT_TERMINALS, walls=None,
                 special_transitions=None, rewards=None, seed=None, slip_distribution=None, p_success=1.0,
                 episode_length_limit=None, *, device=None, env_offset=0, autoreset=True, strict=False, rng="philox",
                 state_layout="auto", wire="canonical", cold_inputs=None):
        """`wire="compact"`: step() takes uint8 actions and returns uint8 (row, col) observations (rows, cols <= 256)
        through `ef_gridworld_step_u8` -- 9 instead of 18 bytes per env-step between host and device, same trajectories.

        `cold_inputs`: the EF_STEP_COLD_INPUTS hint of the step kernel (a latency hint, results do not depend on it):
        True / False, or None = decided per call from the env sets alive on the device (True once they exceed its L2
        together, i.e. when a set's arrays have been evicted by the others by the time it is stepped again).

        `rng="philox"` (default): counter-based draws keyed by (seed, global env index, step).
        `rng="pcg64"`: every env owns the generator the reference would build for it,
        Generator(PCG64(SeedSequence(seed_i))) with seed_i = seed + global env index (or `seed` given as a sequence of
        N ints): `VecGridWorld(1, seed=s, rng="pcg64")` reproduces `GridWorld(seed=s)` draw for draw."""
        if state_layout not in ("auto", "packed", "split"):
            raise ValueError("state_layout must be 'auto', 'packed' or 'split'")
        if wire not in ("canonical", "compact"):
            raise ValueError("wire must be 'canonical' or 'compact'")
        self._w8 = wire == "compact"
        self.wire = wire
        if cold_inputs not in (None, True, False):
            raise ValueError("cold_inputs must be None, True or False")
        self.cold_inputs = cold_inputs
        can_pack = self._can_pack(autoreset, episode_length_limit)
        if state_layout == "packed" and not can_pack:
            raise ValueError("state_layout='packed' needs autoreset=True and 0 < episode_length_limit <= 65535")
        # packed: cell (bits 0-15) and episode_length_counter (bits 16-31) share one int32 per env -- 8 fewer bytes per
        # env-step through HBM; `pos` / `ep_len` stay available as (computed) properties
        self._packed = can_pack and state_layout != "split"
        seed_list = None
        if rng not in ("philox", "pcg64"):
            raise ValueError("rng must be 'philox' or 'pcg64'")
        if rng == "pcg64" and seed is not None and not isinstance(seed, int):
            seed_list = [int(x) for x in seed]
            seed = seed_list[0] if seed_list else None
        validate_args(rows, cols, start_state, terminal_states, walls, special_transitions, rewards, seed,
                      slip_distribution, p_success, episode_length_limit)
        if self._w8 and (rows > 256 or cols > 256):
            raise ValueError("wire='compact' needs rows <= 256 and cols <= 256 (uint8 observations)")
        if self._w8 and rng == "pcg64":
            raise ValueError("wire='compact' is available with rng='philox' only")
        self.spec = GridSpec(rows, cols, start_state, terminal_states, walls, special_transitions, rewards, p_success,
                             episode_length_limit)
        super().__init__(num_envs, device, seed, env_offset, autoreset)
        if self._w8 and "ENVFORGE_HOST_TRANSPORT" not in os.environ:
            # 9 B/env: the kernel's own PCIe accesses (219 us per 2^20 envs) beat copy-engine staging (279 us) -- the
            # opposite of the 18 B/env canonical format (profiles/r01_host_transport.md)
            self.host_transport = "zerocopy"
        torch = self._torch
        self.rows, self.cols = rows, cols
        self.start_state = start_state
        self.terminal_states = terminal_states
        self.walls = self.spec.walls
        self.special_transitions = self.spec.special_transitions
        self.rewards = self.spec.rewards
        self.p_success = p_success
        self.slip_distribution = slip_distribution or {a: (1.0 - p_success) / (len(Action) - 1) for a in Action}
        self.episode_length_limit = episode_length_limit
        self.strict = bool(strict)
        self.single_action_space = spaces.Discrete(4)
        self.single_observation_space = spaces.Tuple((spaces.Discrete(rows), spaces.Discrete(cols)))
        self.action_space = spaces.Batched(self.single_action_space, num_envs)
        self.observation_space = spaces.Batched(self.single_observation_space, num_envs)
        n, dev = self.num_envs, self.device
        self._max_cell = self.spec.n_cells      # `pos` / set_state accept 0 .. n_cells (the last one = off-grid pseudo-cell)
        self._alloc_grid_state()
        # the four per-step outputs live back to back in ONE allocation (16-byte aligned sections), so the host-buffer
        # path brings them back with a single device->host copy
        pad = lambda b: (b + 15) // 16 * 16  # noqa: E731
        ob = 2 * n if self._w8 else 8 * n    # bytes of the observation section: uint8 or int32 (row, col) pairs
        o_rew, o_term, o_trunc = pad(ob), pad(ob) + pad(4 * n), pad(ob) + pad(4 * n) + pad(n)
        self._out_bytes = o_trunc + pad(n)
        self._out = torch.zeros(self._out_bytes, dtype=torch.uint8, device=dev)
        self._out_sections = (o_rew, o_term, o_trunc)
        self._obs_bytes = ob
        self._obs = self._out[:ob].view(n, 2) if self._w8 else self._out[:ob].view(torch.int32).view(n, 2)
        self._reward = self._out[o_rew:o_rew + 4 * n].view(torch.float32)
        self._term = self._out[o_term:o_term + n]
        self._trunc = self._out[o_trunc:o_trunc + n]
        self._final_obs = None
        self._fast_args = None   # cached pointer arguments of the lean step() path
        self._P = None
        self._transition_probs = None
        self.rng = rng
        self._rng_state = self._rng_inc = None
        if rng == "pcg64":
            from . import pcg64
            if seed_list is None:
                base = self._seed if seed is None else int(seed)
                seed_list = [base + self.env_offset + i for i in range(n)]
            if len(seed_list) != n:
                raise ValueError(f"seed sequence must hold {n} integers")
            st, inc = pcg64.pcg64_initial_state(seed_list)
            self._rng_state = torch.from_numpy(st.view(np.int64)).to(dev)
            self._rng_inc = torch.from_numpy(inc.view(np.int64)).to(dev)
            thr, cap = pcg64.thresholds53(p_success)
            self._thr53 = (_lib.C.c_uint64 * 3)(*thr)
            self._cap53 = cap
        self._upload_table()
        self.reset()

    # -- table ----------------------------------------------------------------------------------------------------
    def _start_cell(self):
        s = self.start_state
        if self.spec._in_grid(s):
            return self.spec._cell_index(s)
        if self._w8:   # an off-grid start is reported verbatim by the reference; it need not fit the uint8 observation
            raise ValueError("wire='compact' needs a start_state inside the grid")
        return self.spec.n_cells  # pseudo-cell without outcomes: the reference accepts the start and raises on step

    def _upload_table(self):
        torch = self._torch
        tab = self.spec.device_table(extra_cells=1)
        self._table = torch.from_numpy(tab.view(np.int32)).to(self.device)
        d = _lib.GridDesc()
        d.rows, d.cols, d.n_slots, d.idx_cap = self.rows, self.cols, self.spec.n_slots, self.spec.idx_cap
        d.thr[0], d.thr[1], d.thr[2] = self.spec.thr
        d.start_cell = self._start_cell()
        d.episode_limit = self.episode_length_limit if self.episode_length_limit is not None else 0
        d.table_cells = self.spec.n_cells + 1
        d.table = self._table.data_ptr()
        self._desc = d
        self._fast_args = None
        # an episode_length_limit <= 0 truncates on every step in the reference (counter >= limit); the kernel's
        # "<= 0 means None" encoding cannot express that, so refuse it loudly rather than diverge silently.
        if self.episode_length_limit is not None and self.episode_length_limit <= 0:
            raise ValueError("episode_length_limit <= 0 is not supported by the device path")

    def add_special_transition(self, from_state, action, to_state=None, reward=None):
        """Runtime MDP patch (grid_world.py:214-243); recompiles and re-uploads the table."""
        self.spec.add_special_transition(from_state, action, to_state, reward)
        self._upload_table()

    @property
    def mdp(self):
        return self.spec.mdp_dict()

    @property
    def P(self):
        if self._P is None:
            self._P = self.spec.probability_matrix()
        return self._P

    @property
    def transition_probs(self):
        """The reference's `transition_probs` dict (grid_world.py:82, :401-430), built at construction like there."""
        if self._transition_probs is None:
            self._transition_probs = self.spec.transition_probs(self.slip_distribution)
        return self._transition_probs

    # -- reference-style attributes -----------------------------------------------------------------------------------
    @property
    def state(self):
        """int32 [N,2] (row, col) tensor; assignable with a (row, col) tuple, an [N,2] array or tensor."""
        p = self.pos.to(self._torch.int64)
        return self._torch.stack([p // self.cols, p % self.cols], dim=1).to(self._torch.int32)

    @state.setter
    def state(self, value):
        torch = self._torch
        v = torch.as_tensor(np.asarray(value) if not torch.is_tensor(value) else value).to(self.device)
        v = v.reshape(-1, 2).to(torch.int64)
        if v.shape[0] == 1:
            v = v.expand(self.num_envs, 2)
        inside = (v[:, 0] >= 0) & (v[:, 0] < self.rows) & (v[:, 1] >= 0) & (v[:, 1] < self.cols)
        cell = torch.where(inside, v[:, 0] * self.cols + v[:, 1], torch.full_like(v[:, 0], self.spec.n_cells))
        self.pos = cell.to(torch.int32)

    @property
    def episode_length_counter(self):
        return self.ep_len

    def get_state(self):
        return {"pos": self.pos.clone(), "ep_len": self.ep_len.clone(), "ep_ret": self.ep_ret.clone(), "t": self._t}

    def set_state(self, state):
        self.pos = state["pos"]
        self.ep_len = state["ep_len"]
        self.ep_ret.copy_(state["ep_ret"])
        self._t = int(state.get("t", self._t))

    # -- hot path ---------------------------------------------------------------------------------------------------
    def reset(self, new_start_state=None, *, seed=None, options=None, mask=None):
        """GridWorld.reset (:530-551) for all envs, or those selected by `mask` ([N] bool/uint8 tensor).

        `seed` re-keys the Philox streams (the reference's reset never reseeds; passing nothing keeps that)."""
        if new_start_state is not None:
            if not _is_cell(new_start_state):
                raise ValueError("new_start_state must be a tuple of two integers")
            if self._w8 and not self.spec._in_grid(new_start_state):
                raise ValueError("wire='compact' needs a start_state inside the grid")
            self.start_state = new_start_state
            self.spec.start_state = new_start_state
            self._desc.start_cell = self._start_cell()
        if seed is not None:
            from .vec_env import resolve_seed
            self.seed, self._seed = seed, resolve_seed(seed)
        m = None
        if mask is not None:
            m = self._torch.as_tensor(mask, device=self.device).to(self._torch.uint8).contiguous()
        self._call(self._lib.ef_gridworld_reset, self._desc.start_cell, self.cols, *self._state_ptrs(), _lib.ptr(self.ep_ret),
                   None if self._w8 else _lib.ptr(self._obs), _lib.ptr(m), self.num_envs)
        self._reset_epoch += 1
        if self._w8:   # the reset kernel writes int32 pairs; the compact observation is derived from the state (cold path)
            self._obs.copy_(self.state.clamp(0, 255).to(self._torch.uint8))
        obs = self._obs
        if not self.spec._in_grid(self.start_state):  # off-grid start: report it verbatim like the reference
            obs = self._torch.tensor(self.start_state, dtype=self._torch.int32, device=self.device).expand(self.num_envs, 2)
        return obs, {}

    def step(self, actions, draws=None, *, return_final_obs=False):
        """GridWorld.step (:495-528) for all envs.

        actions  int tensor/array [N] with values in {0,1,2,3} (Action members work at num_envs == 1).  CUDA tensors are
                 used in place; host arrays take the pinned-staging path and numpy arrays come back.
        draws    optional uint32-valued [N] injected outcome draws r (u = r * 2**-32); default: Philox.
        Returns (obs int32 [N,2], reward fp32 [N], terminated bool [N], truncated bool [N], info)."""
        torch = self._torch
        if (draws is None and not return_final_obs and not self.strict and self.rng == "philox"
                and torch.is_tensor(actions) and actions.is_cuda
                and actions.dtype == (torch.uint8 if self._w8 else torch.int32)
                and actions.dim() == 1 and actions.shape[0] == self.num_envs and actions.is_contiguous()):
            # lean path of the usual call (CUDA int32 actions in, CUDA tensors out): the fixed pointer arguments are
            # cached, so one step() costs one ctypes call -- a Python-driven loop then keeps pace with the ~10 us kernel
            fixed = self._fast_args
            state_t = self._pl if self._packed else self._pos_split
            if fixed is not None and (fixed[5] is not self.ep_ret or fixed[6] is not state_t):
                fixed = None     # a state tensor was replaced by assignment: rebuild the cached pointers
            if fixed is None:
                x = _lib.GridStepArgs()
                x.grid = _lib.C.pointer(self._desc)
                x.pos, x.ep_len = self._state_ptrs()
                x.ep_ret = self.ep_ret.data_ptr()
                x.seed, x.env_offset = self._seed, self.env_offset
                x.obs, x.reward = self._obs.data_ptr(), self._reward.data_ptr()
                x.terminated, x.truncated = self._term.data_ptr(), self._trunc.data_ptr()
                x.stats, x.err = self._stats_raw.data_ptr(), self.err.data_ptr()
                x.n = self.num_envs
                self._wire_cold = None      # x.wire not set yet
                fixed = self._fast_args = (x, _lib.C.byref(x), self._lib.ef_gridworld_step_v, self.device.index,
                                           (self._obs, self._reward, self._term.view(torch.bool),
                                            self._trunc.view(torch.bool)), self.ep_ret, state_t)
            x, ref, fn, dev_index, outs = fixed[:5]
            clock = self._clock
            x.action = actions.data_ptr()
            x.t = 0 if clock is not None else self._t
            x.clock = None if clock is None else clock.data_ptr()
            x.seed = self._seed
            x.autoreset = int(self._autoreset)
            cold = self.cold_inputs
            if cold is None:         # the automatic rule, re-evaluated only when the population of env sets changed
                cc = self._cold_cache
                cold = cc[1] if cc[0] == RESIDENCY.generation else self._cold_inputs()
            if cold is not self._wire_cold:
                x.wire = (_lib.EF_WIRE_U8 if self._w8 else 0) | (_lib.EF_STEP_COLD_INPUTS if cold else 0)
                self._wire_cold = cold
            if torch.cuda.current_device() == dev_index:
                x.stream = torch.cuda.current_stream().cuda_stream
                status = fn(ref)
            else:
                with torch.cuda.device(self.device):
                    x.stream = torch.cuda.current_stream().cuda_stream
                    status = fn(ref)
            if status:
                _lib.check(status)
            self._t += 1
            return outs[0], outs[1], outs[2], outs[3], {}
        host_io = not (torch.is_tensor(actions) and actions.is_cuda)
        mode = self.host_transport if host_io else None
        if mode == "pipelined" and (self.rng == "pcg64" or draws is not None or self._clock is not None):
            mode = "staged"     # the pipelined entry point covers the plain Philox step only
        zero_copy = mode == "zerocopy"
        n = self.num_envs
        w8 = self._w8
        act_dtype = torch.uint8 if w8 else torch.int32
        obs_dtype, obs_np = (torch.uint8, np.uint8) if w8 else (torch.int32, np.int32)
        if host_io:
            if torch.is_tensor(actions):      # CPU torch tensor (pinned, of the wire dtype => used in place)
                a = actions.reshape(-1)
                if a.dtype != act_dtype:      # narrowing keeps invalid actions invalid: everything outside 0..255 -> 255
                    a = a.clamp(-1, 255).to(torch.uint8) if w8 else a.to(torch.int32)
            else:
                a = np.asarray(actions)
                a = a.reshape(1) if a.shape == () else a
                if w8 and a.dtype != np.uint8:
                    a = np.where((a < 0) | (a > 255), 255, a).astype(np.uint8)
                elif not w8:
                    a = a.astype(np.int32, copy=False)
            h_act = self._host_in("action", a, act_dtype, (n,))
            if mode == "staged":
                p_act = _lib.ptr(self._to_device("action", h_act, act_dtype, (n,)))
            else:
                p_act = h_act.data_ptr()
        else:
            if tuple(actions.shape) != (n,):
                raise ValueError(f"actions must have shape ({n},), got {tuple(actions.shape)}")
            if actions.dtype == act_dtype:
                act = actions
            else:
                act = actions.clamp(-1, 255).to(torch.uint8) if w8 else actions.to(torch.int32)
            act = act.contiguous()
            p_act = _lib.ptr(act)
        drw = None
        if draws is not None:
            if torch.is_tensor(draws) and draws.is_cuda:
                drw = draws.to(torch.int64).to(torch.int32).contiguous() if draws.dtype != torch.int32 else draws.contiguous()
            else:
                d = np.asarray(draws).astype(np.uint32, copy=False).view(np.int32)
                drw = self._to_device("draw", d, torch.int32, (n,))
        fobs = None
        if return_final_obs:
            if zero_copy:
                fobs = self._pinned("o_final_obs", (n, 2), obs_dtype)
            else:
                if self._final_obs is None:
                    self._final_obs = torch.zeros((n, 2), dtype=obs_dtype, device=self.device)
                fobs = self._final_obs
        o_rew, o_term, o_trunc = self._out_sections
        if zero_copy:
            # the kernel stores obs/reward/flags straight into pinned host memory (UVA pointer): no staging copy
            h_out = self._pinned("o_out", (self._out_bytes,), torch.uint8)
            base = h_out.data_ptr()
        else:
            base = self._out.data_ptr()
        p_obs, p_rew, p_term, p_trunc = base, base + o_rew, base + o_term, base + o_trunc
        if mode == "pipelined":
            h_out = self._pinned("o_out", (self._out_bytes,), torch.uint8)
            hb = h_out.data_ptr()
            d_act = _lib.ptr(self._device_scratch("action", (n,), act_dtype)) if self.host_stage_actions else None
            h_fobs = self._pinned("o_final_obs", (n, 2), obs_dtype) if return_final_obs else None
            with torch.cuda.device(self.device):
                _lib.check((self._lib.ef_gridworld_step_u8_host if w8 else self._lib.ef_gridworld_step_host)(
                    self._pipe(), self._desc, *self._state_ptrs(), _lib.ptr(self.ep_ret), p_act, d_act,
                    self._seed, self._t, self.env_offset, _lib.ptr(self._obs), _lib.ptr(self._reward), _lib.ptr(self._term),
                    _lib.ptr(self._trunc), _lib.ptr(fobs), hb, hb + o_rew, hb + o_term, hb + o_trunc, _lib.ptr(h_fobs),
                    _lib.ptr(self._stats_raw), _lib.ptr(self.err), n, int(self.autoreset), self.host_chunks, self._stream()))
            if return_final_obs:
                fobs = h_fobs
        elif self.rng == "pcg64":
            if drw is not None:
                raise ValueError("injected draws are not available with rng='pcg64' (each env owns its generator)")
            self._call(self._lib.ef_gridworld_step_pcg64, self._desc, self._thr53, self._cap53, *self._state_ptrs(), _lib.ptr(self.ep_ret), _lib.ptr(self._rng_state), _lib.ptr(self._rng_inc),
                       p_act, self.env_offset, p_obs, p_rew, p_term, p_trunc, _lib.ptr(fobs), _lib.ptr(self._stats_raw),
                       _lib.ptr(self.err), n, int(self.autoreset))
        else:
            self._call(self._lib.ef_gridworld_step_u8 if w8 else self._lib.ef_gridworld_step, self._desc, *self._state_ptrs(),
                       _lib.ptr(self.ep_ret), p_act, _lib.ptr(drw), self._seed, self._t_arg(), self.env_offset,
                       p_obs, p_rew, p_term, p_trunc, _lib.ptr(fobs), _lib.ptr(self._stats_raw), _lib.ptr(self.err),
                       _lib.ptr(self._clock), n, int(self.autoreset))
        self._t += 1
        if self.strict:
            self.check_errors()
        info = {}
        if host_io:
            if mode != "staged":
                torch.cuda.current_stream(self.device).synchronize()
                buf = h_out.numpy()
                if return_final_obs:
                    info["final_observation"] = fobs.numpy()
            else:
                named = [("out", self._out)]
                if return_final_obs:
                    named.append(("final_obs", fobs))
                out = self._to_host(named)
                if return_final_obs:
                    info["final_observation"] = out[1]
                buf = out[0]
            return (buf[:self._obs_bytes].view(obs_np).reshape(n, 2), buf[o_rew:o_rew + 4 * n].view(np.float32),
                    buf[o_term:o_term + n].view(np.bool_), buf[o_trunc:o_trunc + n].view(np.bool_), info)
        if return_final_obs:
            info["final_observation"] = fobs
        return self._obs, self._reward, self._term.view(torch.bool), self._trunc.view(torch.bool), info

    # -- persistent step server: K steps without a launch per step, actions from a policy kernel of its own ---------------
    def serve_begin(self, K):
        """Start the step server for K steps on a private stream and return what the action producer needs:

            {"action": int32 [N] CUDA tensor to write, "obs" / "reward" / "terminated" / "truncated": the output tensors to
             read, "act_seq" / "obs_seq": uint32 [chunks] sequence words, "chunk_envs": envs per chunk, "stream": the
             server's stream, "status": int32 [1] (1 after a give-up)}

        Protocol per chunk c and step k (see include/envforge_b200.h, ef_gridworld_serve): the producer waits for
        obs_seq[c] >= k (k > 0), writes action[chunk c] and releases act_seq[c] = k + 1; the server steps the chunk and
        releases obs_seq[c] = k + 1.  The producer must run CONCURRENTLY (its own stream, resident next to the server).
        `serve_end()` waits for the server and returns the last outputs.  Same results as K step() calls."""
        torch = self._torch
        if self.rng != "philox" or self.strict or self._w8 or not self._packed or self._clock is not None:
            raise ValueError("the step server covers the plain Philox step in the packed state layout (canonical wire)")
        if self.__dict__.get("_serve") is not None:
            raise RuntimeError("serve_begin() called again before serve_end()")
        n = self.num_envs
        chunk = int(self._lib.ef_gridworld_serve_chunk_envs(n))
        if chunk <= 0 or n % 4:
            raise ValueError("the step server needs num_envs to be a multiple of 4 and at most 8192 envs per SM")
        chunks = -(-n // chunk)
        st = {"K": int(K), "stream": torch.cuda.Stream(device=self.device),
              "action": torch.zeros(n, dtype=torch.int32, device=self.device),
              "act_seq": torch.zeros(chunks, dtype=torch.int32, device=self.device),
              "obs_seq": torch.zeros(chunks, dtype=torch.int32, device=self.device),
              "status": torch.zeros(1, dtype=torch.int32, device=self.device), "chunk_envs": chunk,
              "obs": self._obs, "reward": self._reward, "terminated": self._term, "truncated": self._trunc}
        cur = torch.cuda.current_stream(self.device)
        st["stream"].wait_stream(cur)          # the zeroed sequence words, resets, earlier steps
        with torch.cuda.stream(st["stream"]):
            self._call(self._lib.ef_gridworld_serve, self._desc, self._pl.data_ptr(), _lib.ptr(self.ep_ret),
                       _lib.ptr(st["action"]), self._seed, self._t, self.env_offset, int(K), _lib.ptr(self._obs),
                       _lib.ptr(self._reward), _lib.ptr(self._term), _lib.ptr(self._trunc), _lib.ptr(self._stats_raw),
                       _lib.ptr(self.err), None, _lib.ptr(st["act_seq"]), _lib.ptr(st["obs_seq"]), _lib.ptr(st["status"]), n)
        self._serve = st
        return st

    def serve_end(self):
        """Wait for the step server started by serve_begin(); returns (obs, reward, terminated, truncated) of the last step."""
        torch = self._torch
        st = self.__dict__.get("_serve")
        if st is None:
            raise RuntimeError("serve_end() without serve_begin()")
        st["stream"].synchronize()
        self._serve = None
        if int(st["status"].item()) != 0:
            taken = int(st["obs_seq"].min().item())   # complete steps of the whole set; chunks that went further are mid-way
            self._t += taken
            raise EnvForgeError(f"the step server gave up waiting for actions after {taken} step(s): is the producer running "
                                "concurrently on another stream?")
        self._t += st["K"]
        torch.cuda.current_stream(self.device).wait_stream(st["stream"])
        return self._obs, self._reward, self._term.view(torch.bool), self._trunc.view(torch.bool)

    def serve_with_demo_policy(self, K):
        """K served steps driven by the built-in stand-in policy kernel (`ef_policy_demo_serve`: action = (row * 3 + col +
        k) & 3 of the previous observation) on a second stream -- the smallest complete example of the protocol."""
        torch = self._torch
        st = self.serve_begin(K)
        producer = self.__dict__.setdefault("_serve_producer_stream", torch.cuda.Stream(device=self.device))
        producer.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(producer), torch.cuda.device(self.device):
            _lib.check(self._lib.ef_policy_demo_serve(_lib.ptr(self._obs), _lib.ptr(st["action"]), _lib.ptr(st["obs_seq"]),
                                                      _lib.ptr(st["act_seq"]), int(K), _lib.ptr(st["status"]), self.num_envs,
                                                      producer.cuda_stream))
        out = self.serve_end()
        producer.synchronize()
        return out

    # -- host-buffer step without the host wait: two env sets in flight keep PCIe busy in both directions ----------------
    def step_async(self, actions):
        """Enqueue one host-buffer step and return at once; `wait()` hands out the results.

        The call a host-side actor loop makes when it drives TWO (or more) env sets: while the host consumes the results of
        set A and picks A's next actions, the step of set B is crossing PCIe -- its actions host->device, its results
        device->host -- so the link never idles between steps (a plain `step(host buffers)` leaves it idle for the host's
        share of every step).  Each env set owns a private stream; transport "staged" (copy engine in, kernel, copy
        engine out -- the default here, see below) or "zerocopy" (one kernel whose action loads and output stores go
        straight to pinned host memory), chosen with `env.async_transport`.  Same kernels, same results as step().  Plain Philox stepping only (no injected draws, no
        final_obs, not strict); actions: pinned CPU tensor of the wire dtype (used in place) or any host array (copied)."""
        torch = self._torch
        if self.rng != "philox" or self.strict or self._clock is not None:
            raise ValueError("step_async() covers the plain Philox step (rng='philox', strict=False, no device clock)")
        st = self.__dict__.get("_async_state")
        if st is None:
            st = self._async_state = {"stream": torch.cuda.Stream(device=self.device), "event": torch.cuda.Event(),
                                      "pending": False}
        if st["pending"]:
            raise RuntimeError("step_async() called again before wait()")
        n, w8 = self.num_envs, self._w8
        act_dtype = torch.uint8 if w8 else torch.int32
        if torch.is_tensor(actions):
            if actions.is_cuda:
                raise ValueError("step_async() takes HOST actions; CUDA tensors go through step()")
            a = actions.reshape(-1)
            if a.dtype != act_dtype:
                a = a.clamp(-1, 255).to(torch.uint8) if w8 else a.to(torch.int32)
        else:
            a = np.asarray(actions).reshape(-1)
            a = (np.where((a < 0) | (a > 255), 255, a).astype(np.uint8) if w8 and a.dtype != np.uint8
                 else a.astype(np.uint8 if w8 else np.int32, copy=False))
        h_act = self._host_in("action", a, act_dtype, (n,))
        h_out = self._pinned("o_out", (self._out_bytes,), torch.uint8)
        o_rew, o_term, o_trunc = self._out_sections
        # With several env sets in flight the copy engines win for BOTH wire formats: the H2D of one set, the kernel of
        # another and the D2H of a third overlap, and the DMA engines keep PCIe at ~54 GB/s device->host, which the kernel's
        # own zero-copy stores do not (measured at 2^20 envs, 2 sets in flight: compact 6.7e9 vs 5.5e9 env-steps/s, int32
        # 3.9e9 vs 1.8e9; profiles/r02l_e2e_probe.txt).  ENVFORGE_ASYNC_TRANSPORT=zerocopy switches for A/B timing.
        mode = os.environ.get("ENVFORGE_ASYNC_TRANSPORT", self.__dict__.get("async_transport", "staged"))
        if mode not in ("staged", "zerocopy"):
            raise ValueError("async transport must be 'staged' or 'zerocopy'")
        stream = st["stream"]
        stream.wait_stream(torch.cuda.current_stream(self.device))   # whatever was enqueued on this env before (reset, ...)
        fn = self._lib.ef_gridworld_step_u8 if w8 else self._lib.ef_gridworld_step
        with torch.cuda.stream(stream):
            if mode == "zerocopy":
                p_act, base = h_act.data_ptr(), h_out.data_ptr()
            else:
                p_act = _lib.ptr(self._to_device("action", h_act, act_dtype, (n,)))
                base = self._out.data_ptr()
            self._call(fn, self._desc, *self._state_ptrs(), _lib.ptr(self.ep_ret), p_act, None, self._seed, self._t,
                       self.env_offset, base, base + o_rew, base + o_term, base + o_trunc, None, _lib.ptr(self._stats_raw),
                       _lib.ptr(self.err), None, n, int(self.autoreset))
            if mode != "zerocopy":
                h_out.copy_(self._out, non_blocking=True)
            st["event"].record(stream)
        st["pending"] = True
        self._t += 1

    def wait(self):
        """Results of the step enqueued by step_async(): (obs, reward, terminated, truncated, info) as numpy views of the
        pinned result block (valid until this env's next host-buffer step)."""
        st = self.__dict__.get("_async_state")
        if st is None or not st["pending"]:
            raise RuntimeError("wait() without a step_async() in flight")
        st["event"].synchronize()
        st["pending"] = False
        n = self.num_envs
        o_rew, o_term, o_trunc = self._out_sections
        buf = self._host["o_out"].numpy()
        obs_np = np.uint8 if self._w8 else np.int32
        return (buf[:self._obs_bytes].view(obs_np).reshape(n, 2), buf[o_rew:o_rew + 4 * n].view(np.float32),
                buf[o_term:o_term + n].view(np.bool_), buf[o_trunc:o_trunc + n].view(np.bool_), {})

    def rollout(self, K, actions=None, *, return_rewards=False, return_obs=False):
        """K fused steps in one kernel (state in registers, auto-reset on).  actions: CUDA int32 [K,N] or None for the
        in-kernel uniform random policy.  Returns None, (rewards fp32 [K,N], flags uint8 [K,N]) if return_rewards, or
        (obs int32 [K,N,2], rewards, flags) if return_obs -- the observations K step() calls would have returned;
        episode statistics accumulate in `self.stats` either way."""
        torch = self._torch
        n = self.num_envs
        if self.rng == "pcg64":
            raise NotImplementedError("rollout() uses the Philox streams; rng='pcg64' supports step() only")
        if actions is not None:
            if not (torch.is_tensor(actions) and actions.is_cuda and actions.dtype == torch.int32
                    and tuple(actions.shape) == (K, n)):
                raise ValueError(f"actions must be a CUDA int32 tensor of shape ({K}, {n})")
            actions = actions.contiguous()
        rew = flg = obs = None
        if return_rewards or return_obs:
            rew = torch.empty((K, n), dtype=torch.float32, device=self.device)
            flg = torch.empty((K, n), dtype=torch.uint8, device=self.device)
        if return_obs:
            obs = torch.empty((K, n, 2), dtype=torch.int32, device=self.device)
            self._call(self._lib.ef_gridworld_rollout_traj, self._desc, *self._state_ptrs(), _lib.ptr(self.ep_ret),
                       _lib.ptr(actions), self._seed, self._t_arg(), self.env_offset, int(K), _lib.ptr(obs), _lib.ptr(rew),
                       _lib.ptr(flg), _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), n)
        else:
            self._call(self._lib.ef_gridworld_rollout, self._desc, *self._state_ptrs(), _lib.ptr(self.ep_ret),
                       _lib.ptr(actions), self._seed, self._t_arg(), self.env_offset, int(K), _lib.ptr(rew), _lib.ptr(flg),
                       _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), n)
        self._t += int(K)
        if return_obs:
            return obs, rew, flg
        return (rew, flg) if return_rewards else None

    def _describe_error(self, count, env, detail, kind_id, kind):
        if kind_id == 1:
            return f"{detail} is not a valid Action (env {env}; {count} rejected env-step(s))"
        cell, a = divmod(detail, 4)
        st = (cell // self.cols, cell % self.cols) if cell < self.spec.n_cells else self.start_state
        if kind_id == 2:
            return f"No outcomes defined for state {st} and action {Action(a)!r} (env {env}; {count} rejected env-step(s))"
        return (f"probabilities do not sum to 1 for state {st} and action {Action(a)!r} "
                f"(env {env}; {count} rejected env-step(s))")


class VecGridWorldLayouts(_GridStateLayout, VecEnv):
    """N GridWorld instances spread over several LAYOUTS in one batch (per-env heterogeneous grids).

    `layouts` is a list of dicts with the per-layout reference kwargs `start_state`, `terminal_states`, `walls`,
    `special_transitions` (each optional, reference defaults); `rows`, `cols`, `rewards`, `p_success`,
    `episode_length_limit` are shared, so one set of outcome thresholds serves every table.  Env i walks layout
    `layout_index[i]`, default `(env_offset + i) % len(layouts)`.  Each layout is compiled by the same table compiler as
    `VecGridWorld` (`GridSpec`, i.e. grid_world.py:245-274 and friends); `mdp_of(l)` / `P_of(l)` export them.
    One launch steps the whole mixed batch (`ef_gridworld_step_layouts`), one launch rolls it out for K fused steps
    (`ef_gridworld_rollout_layouts`); the PCG64 stream mode is a per-layout feature of `VecGridWorld`."""
    _HOST_TRANSPORT = "staged"

    def __init__(self, num_envs, layouts, rows=5, cols=5, rewards=None, p_success=1.0, episode_length_limit=None, *,
                 layout_index=None, device=None, seed=None, env_offset=0, autoreset=True, strict=False, state_layout="auto"):
        if state_layout not in ("auto", "packed", "split"):
            raise ValueError("state_layout must be 'auto', 'packed' or 'split'")
        can_pack = self._can_pack(autoreset, episode_length_limit)
        if state_layout == "packed" and not can_pack:
            raise ValueError("state_layout='packed' needs autoreset=True and 0 < episode_length_limit <= 65535")
        self._packed = can_pack and state_layout != "split"
        self._fast_args = None
        if not layouts:
            raise ValueError("layouts must hold at least one layout dict")
        self.specs = []
        for lay in layouts:
            unknown = set(lay) - {"start_state", "terminal_states", "walls", "special_transitions"}
            if unknown:
                raise ValueError(f"unknown per-layout keys {sorted(unknown)}")
            kw = dict(start_state=lay.get("start_state", (0, 0)), terminal_states=lay.get("terminal_states", _DEFAULT_TERMINALS),
                      walls=lay.get("walls"), special_transitions=lay.get("special_transitions"))
            validate_args(rows, cols, kw["start_state"], kw["terminal_states"], kw["walls"], kw["special_transitions"],
                          rewards, seed, None, p_success, episode_length_limit)
            self.specs.append(GridSpec(rows, cols, kw["start_state"], kw["terminal_states"], kw["walls"],
                                       kw["special_transitions"], rewards, p_success, episode_length_limit))
        if episode_length_limit is not None and episode_length_limit <= 0:
            raise ValueError("episode_length_limit <= 0 is not supported by the device path")
        super().__init__(num_envs, device, seed, env_offset, autoreset)
        torch = self._torch
        self.rows, self.cols, self.p_success, self.episode_length_limit = rows, cols, p_success, episode_length_limit
        self.n_layouts = len(self.specs)
        self.strict = bool(strict)
        self.single_action_space = spaces.Discrete(4)
        self.single_observation_space = spaces.Tuple((spaces.Discrete(rows), spaces.Discrete(cols)))
        self.action_space = spaces.Batched(self.single_action_space, num_envs)
        self.observation_space = spaces.Batched(self.single_observation_space, num_envs)
        n, dev = self.num_envs, self.device
        self._layout_index = None
        if layout_index is not None:
            li = np.asarray(layout_index, np.int64).reshape(-1)
            if li.shape[0] != n or (li < 0).any() or (li >= self.n_layouts).any():
                raise ValueError(f"layout_index must hold {n} values in 0..{self.n_layouts - 1}")
            self._layout_index = torch.from_numpy(li.astype(np.int32)).to(dev)
        self._max_cell = rows * cols            # `pos` / set_state accept 0 .. rows * cols (the last one = pseudo-cell)
        self._alloc_grid_state()
        self._obs = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._term = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._final_obs = None
        self._upload_tables()
        self.reset()

    def _upload_tables(self):
        torch = self._torch
        S = self.rows * self.cols
        starts = []
        for sp in self.specs:
            s = sp.start_state
            starts.append(sp._cell_index(s) if sp._in_grid(s) else S)   # pseudo-cell S: the reference raises on step
        tabs = np.concatenate([sp.device_table(extra_cells=1) for sp in self.specs])
        self._table = torch.from_numpy(tabs.view(np.int32)).to(self.device)
        self._start_cells = torch.from_numpy(np.array(starts, np.int32)).to(self.device)
        sp0 = self.specs[0]
        d = _lib.GridDesc()
        d.rows, d.cols, d.n_slots, d.idx_cap = self.rows, self.cols, sp0.n_slots, sp0.idx_cap
        d.thr[0], d.thr[1], d.thr[2] = sp0.thr
        d.start_cell = 0
        d.episode_limit = self.episode_length_limit if self.episode_length_limit is not None else 0
        d.table_cells = S + 1
        d.table = self._table.data_ptr()
        self._desc = d

    def layout_of_env(self):
        """int64 [N] layout index of every env (host array)."""
        if self._layout_index is not None:
            return self._layout_index.cpu().numpy().astype(np.int64)
        return (self.env_offset + np.arange(self.num_envs, dtype=np.int64)) % self.n_layouts

    def mdp_of(self, layout):
        return self.specs[layout].mdp_dict()

    def P_of(self, layout):
        return self.specs[layout].probability_matrix()

    def add_special_transition(self, layout, from_state, action, to_state=None, reward=None):
        self.specs[layout].add_special_transition(from_state, action, to_state, reward)
        self._upload_tables()

    @property
    def state(self):
        p = self.pos.to(self._torch.int64)
        return self._torch.stack([p // self.cols, p % self.cols], dim=1).to(self._torch.int32)

    @property
    def episode_length_counter(self):
        return self.ep_len

    def get_state(self):
        return {"pos": self.pos.clone(), "ep_len": self.ep_len.clone(), "ep_ret": self.ep_ret.clone(), "t": self._t}

    def set_state(self, state):
        self.pos = state["pos"]
        self.ep_len = state["ep_len"]
        self.ep_ret.copy_(state["ep_ret"])
        self._t = int(state.get("t", self._t))

    def reset(self, *, seed=None, mask=None):
        """Every env (or those selected by `mask`) back to the start state of ITS layout (grid_world.py:530-551)."""
        torch = self._torch
        if seed is not None:
            self.seed, self._seed = seed, resolve_seed(seed)
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        self._call(self._lib.ef_gridworld_reset_layouts, self.n_layouts, _lib.ptr(self._layout_index),
                   _lib.ptr(self._start_cells), self.cols, *self._state_ptrs(), _lib.ptr(self.ep_ret), _lib.ptr(self._obs),
                   _lib.ptr(m), self.env_offset, self.num_envs)
        self._reset_epoch += 1
        return self.state, {}

    def step(self, actions, draws=None, *, return_final_obs=False):
        """GridWorld.step (:495-528) for the whole mixed batch; arguments and returns as `VecGridWorld.step`."""
        torch, n = self._torch, self.num_envs
        host_io = not (torch.is_tensor(actions) and actions.is_cuda)
        if host_io:
            a = actions.reshape(-1).to(torch.int32) if torch.is_tensor(actions) else \
                np.asarray(actions).reshape(-1).astype(np.int32, copy=False)
            act = self._to_device("action", self._host_in("action", a, torch.int32, (n,)), torch.int32, (n,))
        else:
            if tuple(actions.shape) != (n,):
                raise ValueError(f"actions must have shape ({n},), got {tuple(actions.shape)}")
            act = (actions if actions.dtype == torch.int32 else actions.to(torch.int32)).contiguous()
        drw = None
        if draws is not None:
            if torch.is_tensor(draws) and draws.is_cuda:
                drw = (draws if draws.dtype == torch.int32 else draws.to(torch.int64).to(torch.int32)).contiguous()
            else:
                drw = self._to_device("draw", np.asarray(draws).astype(np.uint32, copy=False).view(np.int32), torch.int32, (n,))
        fobs = None
        if return_final_obs:
            if self._final_obs is None:
                self._final_obs = torch.zeros((n, 2), dtype=torch.int32, device=self.device)
            fobs = self._final_obs
        self._call(self._lib.ef_gridworld_step_layouts, self._desc, self.n_layouts, _lib.ptr(self._layout_index),
                   _lib.ptr(self._start_cells), *self._state_ptrs(), _lib.ptr(self.ep_ret), _lib.ptr(act),
                   _lib.ptr(drw), self._seed, self._t_arg(), self.env_offset, _lib.ptr(self._obs), _lib.ptr(self._reward),
                   _lib.ptr(self._term), _lib.ptr(self._trunc), _lib.ptr(fobs), _lib.ptr(self._stats_raw), _lib.ptr(self.err),
                   _lib.ptr(self._clock), n, int(self.autoreset))
        self._t += 1
        if self.strict:
            self.check_errors()
        info = {}
        if host_io:
            named = [("obs", self._obs), ("reward", self._reward), ("term", self._term), ("trunc", self._trunc)]
            out = self._to_host(named + ([("final_obs", fobs)] if return_final_obs else []))
            if return_final_obs:
                info["final_observation"] = out[4]
            return out[0], out[1], out[2].view(np.bool_), out[3].view(np.bool_), info
        if return_final_obs:
            info["final_observation"] = fobs
        return self._obs, self._reward, self._term.view(torch.bool), self._trunc.view(torch.bool), info

    def rollout(self, K, actions=None, *, return_rewards=False, return_obs=False):
        """K fused steps of the mixed batch in one kernel (state in registers, auto-reset to each env's own start cell);
        arguments and returns as `VecGridWorld.rollout`."""
        torch, n = self._torch, self.num_envs
        if actions is not None:
            if not (torch.is_tensor(actions) and actions.is_cuda and actions.dtype == torch.int32
                    and tuple(actions.shape) == (K, n)):
                raise ValueError(f"actions must be a CUDA int32 tensor of shape ({K}, {n})")
            actions = actions.contiguous()
        rew = flg = obs = None
        if return_rewards or return_obs:
            rew = torch.empty((K, n), dtype=torch.float32, device=self.device)
            flg = torch.empty((K, n), dtype=torch.uint8, device=self.device)
        if return_obs:
            obs = torch.empty((K, n, 2), dtype=torch.int32, device=self.device)
        self._call(self._lib.ef_gridworld_rollout_layouts, self._desc, self.n_layouts, _lib.ptr(self._layout_index),
                   _lib.ptr(self._start_cells), *self._state_ptrs(), _lib.ptr(self.ep_ret), _lib.ptr(actions), self._seed,
                   self._t_arg(), self.env_offset, int(K), _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(flg),
                   _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), n)
        self._t += int(K)
        if return_obs:
            return obs, rew, flg
        return (rew, flg) if return_rewards else None

    def _describe_error(self, count, env, detail, kind_id, kind):
        if kind_id == 1:
            return f"{detail} is not a valid Action (env {env}; {count} rejected env-step(s))"
        cell, a = divmod(detail, 4)
        return (f"{'No outcomes defined' if kind_id == 2 else 'probabilities do not sum to 1'} for state "
                f"{(cell // self.cols, cell % self.cols)} and action {Action(a)!r} (env {env}; {count} rejected env-step(s))")
