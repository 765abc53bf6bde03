"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference Labyrinth stepping path.

Follows mariusdgm/rl-envs-forge v5.22.0, rl_envs_forge/envs/labyrinth/:
    labyrinth.py  state :195-208, build_state_matrix :210-215, build_state_around_the_player :217-243, step :245-288,
                  is_valid_move :290-309, set_state :311-331, agent_move :333-345
    entities/player.py  potential_next_position :61-82
    constants.py  Action :4-8, WALL/PATH/TARGET/START/PLAYER :12-16
Maze GENERATION (maze/maze.py, room.py, corridor.py: ~2500 lines of host-side procedural code) is out of scope: a maze
enters as its finished `grid` plus start / target positions, exactly what `Labyrinth.maze` exposes after construction.

Parity status: PINNED by tests/golden/labyrinth_*.npz (mazes generated and stepped by the unmodified reference, with the
observation of every step recorded) -- tests/test_oracle_labyrinth.py.  Nothing in the product package may import this.
"""
import numpy as np

WALL, PATH, TARGET, START, PLAYER = 0, 1, 2, 3, 4
DELTAS = ((-1, 0), (0, 1), (1, 0), (0, -1))  # UP, RIGHT, DOWN, LEFT (player.py:72-80)

DEFAULT_REWARDS = {"neutral_reward": -0.01, "wall_collision_reward": -1, "target_reached_reward": 10}


class LabyrinthPort:
    """Scalar port: one maze, one player."""

    def __init__(self, grid, start, target, state_vision_range=None, reward_schema=None):
        self.grid = np.asarray(grid)
        self.rows, self.cols = self.grid.shape
        self.start, self.target = tuple(start), tuple(target)
        self.pos = tuple(start)
        self.vision = state_vision_range
        self.rewards = dict(reward_schema or DEFAULT_REWARDS)

    def full_state(self):
        s = self.grid.astype(np.uint8).copy()          # build_state_matrix :210-215: target first, player on top
        s[self.target] = TARGET
        s[self.pos] = PLAYER
        return s

    @property
    def state(self):
        full = self.full_state()
        if self.vision is None:
            return full
        v = self.vision                                 # build_state_around_the_player :217-243
        r0, r1 = max(0, self.pos[0] - v), min(self.rows, self.pos[0] + v + 1)
        c0, c1 = max(0, self.pos[1] - v), min(self.cols, self.pos[1] + v + 1)
        local = full[r0:r1, c0:c1]
        out = np.full((2 * v + 1, 2 * v + 1), WALL, np.uint8)
        ro, co = v - (self.pos[0] - r0), v - (self.pos[1] - c0)
        out[ro:ro + local.shape[0], co:co + local.shape[1]] = local
        return out

    def step(self, action):
        if action not in (0, 1, 2, 3):
            raise ValueError(f"{action} is not a valid Action")
        d = DELTAS[int(action)]
        nxt = (self.pos[0] + d[0], self.pos[1] + d[1])
        inside = 0 <= nxt[0] < self.rows and 0 <= nxt[1] < self.cols
        if not inside or self.grid[nxt] == WALL:        # is_valid_move :290-309
            return self.state, self.rewards["wall_collision_reward"], False, False, {"info": "Invalid move!"}
        self.pos = nxt
        if self.pos == self.target:
            return self.state, self.rewards["target_reached_reward"], True, False, {"info": "Reached the target!"}
        return self.state, self.rewards["neutral_reward"], False, False, {}


def vec_step(grids, maze_index, target_cell, pos_cell, action, cols, rewards=None):
    """Vectorised transition.  grids uint8 [M, rows*cols]; maze_index [N]; target_cell [M]; pos_cell / action [N].
    -> (pos_cell, reward f32, done, invalid_action).  Invalid actions (outside 0..3) leave the env untouched."""
    rw = dict(rewards or DEFAULT_REWARDS)
    grids = np.asarray(grids)
    cells = grids.shape[1]
    rows = cells // cols
    pos = np.asarray(pos_cell, np.int64)
    act = np.asarray(action, np.int64)
    invalid = (act < 0) | (act > 3)
    a = np.where(invalid, 0, act)
    dr = np.array([-1, 0, 1, 0])[a]
    dc = np.array([0, 1, 0, -1])[a]
    r, c = pos // cols + dr, pos % cols + dc
    inside = (r >= 0) & (r < rows) & (c >= 0) & (c < cols)
    nxt = np.where(inside, r * cols + c, 0)
    blocked = ~inside | (grids[maze_index, nxt] == WALL)
    new_pos = np.where(blocked | invalid, pos, nxt)
    done = ~blocked & ~invalid & (new_pos == np.asarray(target_cell)[maze_index])
    reward = np.where(blocked, np.float32(rw["wall_collision_reward"]),
                      np.where(done, np.float32(rw["target_reached_reward"]), np.float32(rw["neutral_reward"])))
    reward = np.where(invalid, np.float32(0), reward).astype(np.float32)
    return new_pos, reward, done, invalid


def vec_obs(grids, maze_index, target_cell, pos_cell, cols, vision=None):
    """Vectorised observation: uint8 [N, rows, cols] or [N, 2v+1, 2v+1]."""
    grids = np.asarray(grids)
    n = len(pos_cell)
    cells = grids.shape[1]
    rows = cells // cols
    full = grids[maze_index].astype(np.uint8).copy()
    idx = np.arange(n)
    full[idx, np.asarray(target_cell)[maze_index]] = TARGET
    full[idx, pos_cell] = PLAYER
    full = full.reshape(n, rows, cols)
    if vision is None:
        return full
    v = vision
    pad = np.full((n, rows + 2 * v, cols + 2 * v), WALL, np.uint8)
    pad[:, v:v + rows, v:v + cols] = full
    pr, pc = np.asarray(pos_cell) // cols, np.asarray(pos_cell) % cols
    d = 2 * v + 1
    ii = pr[:, None, None] + np.arange(d)[None, :, None]
    jj = pc[:, None, None] + np.arange(d)[None, None, :]
    return pad[idx[:, None, None], ii, jj]
