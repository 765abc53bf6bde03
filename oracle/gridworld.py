"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference GridWorld hot path.

Follows rl_envs_forge/envs/grid_world/grid_world.py of mariusdgm/rl-envs-forge v5.22.0:
    MDP compiler  build_mdp :245-274, calculate_outcomes :327-359, calculate_transition :361-377,
                  default_transition :379-399, calculate_reward :467-493, add_special_transition :214-243
    P tensor      create_probability_matrix :276-300
    step / reset  :495-528 / :530-551
    outcome draw  numpy Generator.choice(n, p=...) (third-party, numpy 1.26.4 pinned by poetry.lock:786-787):
                  |sum(p)-1| <= sqrt(eps) check, cdf = cumsum(p)/cumsum(p)[-1], idx = searchsorted(cdf, u, 'right')

Two layers:
  * `GridWorldPort`  -- one-instance, dict-based, scalar restatement (what `cpu_baseline` times and what the
                       live-reference / golden tests pin).
  * `FlatTables` + `vec_step` -- numpy-vectorised restatement over N envs used for the large-N parity checks;
                       itself checked against `GridWorldPort`.

Parity status: PINNED -- against golden vectors generated from the reference in the build container
(tests/golden/gridworld_*.npz and engine_*.npz, made by oracle/gen_golden.py from the unmodified reference;
tests/test_oracle_gridworld.py, tests/test_oracle_engine_golden.py).

Nothing in the product package may import this module.
"""
import math

import numpy as np

UP, RIGHT, DOWN, LEFT = 0, 1, 2, 3
ACTIONS = (UP, RIGHT, DOWN, LEFT)
_SQRT_EPS = math.sqrt(np.finfo(np.float64).eps)
_DEFAULT_REWARDS = {"valid_move": -0.1, "wall_collision": -1, "out_of_bounds": -1, "default": 0.0}


class ChoiceError(ValueError):
    """numpy's `probabilities do not sum to 1` raised by Generator.choice."""


def numpy_choice_index(probs, u):
    """Index numpy's Generator.choice(len(probs), p=probs) returns when its one uniform draw equals u."""
    p = np.asarray(probs, dtype=np.float64)
    if abs(math.fsum(p) - 1.0) > _SQRT_EPS:
        raise ChoiceError("probabilities do not sum to 1")
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


class GridWorldPort:
    """Scalar restatement of reference GridWorld (construction, MDP, step, reset).  `draw` supplies uniforms."""

    def __init__(self, rows=5, cols=5, start_state=(0, 0), terminal_states=None, walls=None,
                 special_transitions=None, rewards=None, p_success=1.0, episode_length_limit=None):
        self.rows, self.cols = rows, cols
        self.start_state = start_state
        self.terminal_states = {(3, 4): 1.0} if terminal_states is None else terminal_states
        self.walls = walls or set()
        self.special_transitions = special_transitions or {}
        self.rewards = rewards or dict(_DEFAULT_REWARDS)
        self.p_success = p_success
        self.episode_length_limit = episode_length_limit
        self.state = start_state
        self.episode_length_counter = 0
        self.mdp = self._build_mdp()

    # -- table compiler ------------------------------------------------------------------------------------
    def _move(self, state, action):  # grid_world.py:379-399
        r, c = state
        if action == UP:
            return (max(r - 1, 0), c)
        if action == DOWN:
            return (min(r + 1, self.rows - 1), c)
        if action == LEFT:
            return (r, max(c - 1, 0))
        return (r, min(c + 1, self.cols - 1))

    def _reward(self, cur, nxt):  # grid_world.py:467-493 (the wall_collision branch is unreachable there too)
        if nxt in self.terminal_states:
            return self.terminal_states[nxt]
        if nxt == cur:
            return self.rewards.get("out_of_bounds", -1)
        if nxt in self.walls:
            return self.rewards.get("wall_collision", -1)
        return self.rewards.get("valid_move", -0.1)

    def _transition(self, state, action):  # grid_world.py:361-377
        if (state, action) in self.special_transitions:
            to, rew = self.special_transitions[(state, action)]
            return to, rew, to in self.terminal_states
        to = self._move(state, action)
        if to in self.walls:
            to = state
        return to, self._reward(state, to), to in self.terminal_states

    def _outcomes(self, state, action):  # grid_world.py:327-359
        out = [(*self._transition(state, action), self.p_success)]
        if self.p_success < 1.0:
            q = (1.0 - self.p_success) / (len(ACTIONS) - 1)
            out += [(*self._transition(state, a), q) for a in ACTIONS if a != action]
        return out

    def _build_mdp(self):  # grid_world.py:245-274
        mdp = {}
        for r in range(self.rows):
            for c in range(self.cols):
                s = (r, c)
                if s in self.walls or s in self.terminal_states:
                    continue
                for a in ACTIONS:
                    mdp[(s, a)] = self._outcomes(s, a)
        for t in self.terminal_states:
            for a in ACTIONS:
                mdp[(t, a)] = [(t, 0, True, 1.0)]
        for (s, a), (to, rew) in self.special_transitions.items():
            mdp[(s, int(a))] = [(to, rew, to in self.terminal_states, self.p_success)]
        return mdp

    def add_special_transition(self, from_state, action, to_state=None, reward=None):  # grid_world.py:214-243
        if to_state is None:
            to_state = self._next_state_no_special(from_state, action)
        if reward is None:
            reward = self.rewards["default"]
        self.mdp[(from_state, int(action))] = [(to_state, reward, to_state in self.terminal_states, self.p_success)]

    def _next_state_no_special(self, state, action):  # grid_world.py:432-465 with check_special=False
        nxt = self._move(state, action)
        return state if (nxt in self.walls or nxt == state) else nxt

    def probability_matrix(self):  # grid_world.py:276-300
        S = self.rows * self.cols
        P = np.zeros((S, S, 4))
        for r in range(self.rows):
            for c in range(self.cols):
                s, i = (r, c), r * self.cols + c
                if s in self.terminal_states or s in self.walls:
                    P[i, i, :] = 1.0
                    continue
                for a in ACTIONS:
                    for nxt, _, _, p in self.mdp.get((s, a), []):
                        P[i, nxt[0] * self.cols + nxt[1], a] += p
        return P

    # -- hot path ------------------------------------------------------------------------------------------
    def step(self, action, u):  # grid_world.py:495-528; u = the one uniform Generator.choice consumes
        if action not in ACTIONS:
            raise ValueError(f"{action} is not a valid Action")
        outcomes = self.mdp.get((self.state, int(action)), [])
        if not outcomes:
            raise ValueError(f"No outcomes defined for state {self.state} and action {action}")
        idx = numpy_choice_index([o[3] for o in outcomes], u)
        nxt, reward, done, _ = outcomes[idx]
        self.state = nxt
        self.episode_length_counter += 1
        truncated = (self.episode_length_limit is not None
                     and self.episode_length_counter >= self.episode_length_limit)
        return self.state, reward, done, truncated, {}

    def reset(self, new_start_state=None):  # grid_world.py:530-551
        if new_start_state is not None:
            if (not isinstance(new_start_state, tuple) or len(new_start_state) != 2
                    or not all(isinstance(x, int) for x in new_start_state)):
                raise ValueError("new_start_state must be a tuple of two integers")
            self.start_state = new_start_state
        self.state = self.start_state
        self.episode_length_counter = 0
        return self.state, {}


# -- vectorised restatement --------------------------------------------------------------------------------
class FlatTables:
    """The dict MDP flattened to arrays indexed by row = cell*4 + action, slot = outcome index (padded to 4).

    `count[row]` = number of outcomes (0 => the reference raises `No outcomes defined`), `raises[row]` = the
    reference's choice() would raise ChoiceError, `cdf[row, k]` = numpy's normalised cumulative probabilities.
    """

    def __init__(self, port: GridWorldPort):
        self.rows, self.cols = port.rows, port.cols
        S = port.rows * port.cols
        self.count = np.zeros(S * 4, np.int32)
        self.raises = np.zeros(S * 4, bool)
        self.nxt = np.zeros((S * 4, 4), np.int64)
        self.reward = np.zeros((S * 4, 4), np.float64)
        self.done = np.zeros((S * 4, 4), bool)
        self.cdf = np.full((S * 4, 4), np.inf)
        for (s, a), outs in port.mdp.items():
            if not (0 <= s[0] < port.rows and 0 <= s[1] < port.cols):
                continue  # an off-grid key can never be the agent's cell index
            row = (s[0] * port.cols + s[1]) * 4 + a
            self.count[row] = len(outs)
            p = np.array([o[3] for o in outs], np.float64)
            if abs(math.fsum(p) - 1.0) > _SQRT_EPS:
                self.raises[row] = True
            c = np.cumsum(p)
            with np.errstate(invalid="ignore", divide="ignore"):   # a raising row (e.g. p_success = 0 special) may sum to 0
                c /= c[-1]
            self.cdf[row, :len(outs)] = c
            for k, (to, rew, dn, _) in enumerate(outs):
                self.nxt[row, k] = to[0] * port.cols + to[1]
                self.reward[row, k] = rew
                self.done[row, k] = dn


def vec_step(tab: FlatTables, pos, ep_len, ep_ret, action, r_u32, *, limit, start_pos, autoreset):
    """Advance N envs one transition; mirrors the device kernel's contract (`ef_gridworld_step`).

    u = r_u32 * 2**-32 is the uniform handed to numpy's choice rule.  Envs whose (cell, action) row is missing
    or raising are reported in `error` and left untouched (the reference raises before mutating anything).
    ep_ret accumulates in float32 exactly like the device.  With autoreset, a finished env (done or truncated)
    has its episode folded into the `finished_*` outputs and is put back on start_pos with zeroed counters.
    """
    pos = np.asarray(pos, np.int64)
    action = np.asarray(action, np.int64)
    bad_action = (action < 0) | (action > 3)
    row = pos * 4 + np.where(bad_action, 0, action)
    error = bad_action | (tab.count[row] == 0) | tab.raises[row]
    u = np.asarray(r_u32, np.uint32).astype(np.float64) * 2.0 ** -32
    idx = (tab.cdf[row] <= u[:, None]).sum(axis=1)  # searchsorted(cdf, u, 'right') per row
    idx = np.where(error, 0, idx)
    nxt = tab.nxt[row, idx]
    rew = tab.reward[row, idx].astype(np.float32)
    done = tab.done[row, idx]
    new_pos = np.where(error, pos, nxt)
    new_len = np.asarray(ep_len, np.int32) + np.where(error, 0, 1).astype(np.int32)
    new_ret = (np.asarray(ep_ret, np.float32) + np.where(error, np.float32(0), rew)).astype(np.float32)
    rew = np.where(error, np.float32(0), rew).astype(np.float32)
    term = done & ~error
    trunc = ((new_len >= limit) & ~error) if limit is not None else np.zeros_like(term)
    out = dict(reward=rew, terminated=term, truncated=trunc, error=error,
               final_pos=new_pos.copy(), final_len=new_len.copy(), final_ret=new_ret.copy())
    fin = (term | trunc) if autoreset else np.zeros_like(term)
    out["finished"] = fin
    out["pos"] = np.where(fin, start_pos, new_pos)
    out["ep_len"] = np.where(fin, 0, new_len).astype(np.int32)
    out["ep_ret"] = np.where(fin, np.float32(0), new_ret).astype(np.float32)
    out["obs"] = np.stack([out["pos"] // tab.cols, out["pos"] % tab.cols], axis=1).astype(np.int32)
    return out
