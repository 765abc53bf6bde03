"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference; see oracle/ref_shim):
    python -m oracle.gen_golden
The vectors are committed; the GPU box (which has no reference tree) replays them.  Every file stores the
constructor kwargs as a JSON string under `config`, the injected inputs (actions, and for GridWorld the uniform
`u` that Generator.choice consumed on that step, read from a clone of the env's PCG64 state) and the reference's
outputs.  numpy version used is recorded in `numpy_version` (matters for PendulumDisk, see oracle/pendulums.py).
"""
import json
import math
import os

import numpy as np

from oracle.ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CONFIG2_LAYOUT = dict(  # SURVEY.md 8(d) config 2 -- canonical 8x8 layout
    rows=8, cols=8, start_state=(0, 0),
    walls={(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)},
    terminal_states={(7, 7): 1.0, (0, 7): -1.0},
    special_transitions={((3, 3), 0): ((6, 6), 0.5)},
    episode_length_limit=200,
)


def _jsonable_grid_cfg(cfg):
    c = dict(cfg)
    if c.get("walls") is not None:
        c["walls"] = sorted(list(w) for w in c["walls"])
    if c.get("terminal_states") is not None:
        c["terminal_states"] = [[list(k), v] for k, v in c["terminal_states"].items()]
    if c.get("special_transitions") is not None:
        c["special_transitions"] = [[list(k[0]), int(k[1]), list(v[0]), v[1]]
                                    for k, v in c["special_transitions"].items()]
    if "start_state" in c:
        c["start_state"] = list(c["start_state"])
    return json.dumps(c)


def _peek_uniform(gen):
    clone = np.random.Generator(np.random.PCG64())
    clone.bit_generator.state = gen.bit_generator.state
    return clone.random()


def _dump_mdp(env):
    S = env.rows * env.cols
    count = np.zeros(S * 4, np.int32)
    nxt = np.full((S * 4, 4, 2), -1, np.int32)
    rew = np.zeros((S * 4, 4), np.float64)
    done = np.zeros((S * 4, 4), bool)
    prob = np.zeros((S * 4, 4), np.float64)
    for (s, a), outs in env.mdp.items():
        if not (0 <= s[0] < env.rows and 0 <= s[1] < env.cols):
            continue
        row = (s[0] * env.cols + s[1]) * 4 + int(a)
        count[row] = len(outs)
        for k, (to, r, d, p) in enumerate(outs):
            nxt[row, k] = to
            rew[row, k] = r
            done[row, k] = d
            prob[row, k] = p
    return dict(mdp_count=count, mdp_next=nxt, mdp_reward=rew, mdp_done=done, mdp_prob=prob, P=env.P)


def gridworld_case(R, name, cfg, n_steps, seed, action_seed, runtime_specials=()):
    kw = dict(cfg)
    if kw.get("special_transitions"):
        kw["special_transitions"] = {(k[0], R.Action(k[1])): v for k, v in kw["special_transitions"].items()}
    env = R.GridWorld(seed=seed, **kw)
    for (fs, a, ts, rw) in runtime_specials:
        env.add_special_transition(fs, R.Action(a), ts, rw)
    arng = np.random.default_rng(action_seed)
    actions = arng.integers(0, 4, n_steps)
    rec = dict(action=[], u=[], pre=[], post=[], reward=[], done=[], trunc=[], raised=[], ep_len=[])
    env.reset()
    for a in actions:
        pre = env.state
        u = _peek_uniform(env.np_random)
        try:
            s, r, d, t, _ = env.step(int(a))
            raised = 0
        except ValueError as e:
            # reference raises before mutating state; numpy's choice() raises before drawing
            s, r, d, t = env.state, 0.0, False, False
            raised = 1 if "sum to 1" in str(e) else 2
        rec["action"].append(a); rec["u"].append(u); rec["pre"].append(pre); rec["post"].append(s)
        rec["reward"].append(r); rec["done"].append(d); rec["trunc"].append(t); rec["raised"].append(raised)
        rec["ep_len"].append(env.episode_length_counter)
        if d or t:
            env.reset()
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(_dump_mdp(env))
    out["config"] = _jsonable_grid_cfg(cfg)
    out["runtime_specials"] = json.dumps([[list(fs), int(a), None if ts is None else list(ts), rw]
                                          for fs, a, ts, rw in runtime_specials])
    out["seed"] = seed
    out["numpy_version"] = np.__version__
    np.savez_compressed(os.path.join(OUT, f"gridworld_{name}.npz"), **out)
    print(name, "steps", n_steps, "raised", int((out["raised"] > 0).sum()), "episodes", int((out["done"] | out["trunc"]).sum()))


def random_grid_cfg(rng):
    rows, cols = int(rng.integers(2, 10)), int(rng.integers(2, 10))
    cells = [(r, c) for r in range(rows) for c in range(cols)]
    rng.shuffle(cells)
    n_w = int(rng.integers(0, max(1, len(cells) // 4)))
    n_t = int(rng.integers(1, 3))
    walls = set(map(tuple, cells[:n_w]))
    terms = {tuple(c): float(np.round(rng.uniform(-2, 2), 3)) for c in cells[n_w:n_w + n_t]}
    free = [tuple(c) for c in cells[n_w + n_t:]]
    start = free[0] if free else tuple(cells[-1])
    specials = {}
    for _ in range(int(rng.integers(0, 4))):
        if not free:
            break
        fs = free[int(rng.integers(len(free)))]
        to = tuple(cells[int(rng.integers(len(cells)))])
        specials[(fs, int(rng.integers(4)))] = (to, float(np.round(rng.uniform(-1, 1), 3)))
    p = float(rng.choice([1.0, 1.0, 0.9, 0.7, 0.5, 0.25, 0.0]))
    rewards = None if rng.random() < 0.5 else {"valid_move": float(np.round(rng.uniform(-0.5, 0), 3)),
                                                 "out_of_bounds": float(np.round(rng.uniform(-2, 0), 3)),
                                                 "wall_collision": -3.0, "default": 0.25}
    lim = None if rng.random() < 0.3 else int(rng.integers(3, 60))
    return dict(rows=rows, cols=cols, start_state=tuple(int(x) for x in start), walls=walls,
                terminal_states=terms, special_transitions=specials, rewards=rewards, p_success=p,
                episode_length_limit=lim)


BANDIT_ARMS = {  # SURVEY.md 8(d) config 3 (families/params of tests/envs/k_armed_bandit/test_k_armed_bandit.py:65-72)
    0: dict(distribution="normal", mean=5, std=1),
    1: dict(distribution="uniform", low=4, high=6),
    2: dict(distribution="exponential", scale=0.2),
    3: dict(distribution="gamma", shape=9, scale=0.55),
    4: dict(distribution="logistic", loc=5, scale=0.3),
    5: dict(distribution="lognormal", mean=1.6, sigma=0.3),
    6: dict(distribution="weibull", a=2),
    7: dict(distribution="beta", a=2, b=5),
}


def bandit_shift_functions():
    return [{"function": lambda t: 0.1 * t, "target_param": "mean"},
            {"function": lambda t: -0.01 * t, "target_param": "std"}]


def bandit_case(R, name, k, arm_params, seed, n_steps, action_seed, reset_at=()):
    env = R.KArmedBandit(k=k, arm_params=arm_params, seed=seed)
    actions = np.random.default_rng(action_seed).integers(0, k, n_steps)
    rewards, timesteps, means0 = [], [], []
    for i, a in enumerate(actions):
        if i in reset_at:
            env.reset()
        _, r, d, t, _ = env.step(int(a))
        assert d is True and t is False
        rewards.append(r)
        timesteps.append(env.timestep)
        means0.append(env.arms[0].current_params.get("mean", np.nan))
    np.savez_compressed(os.path.join(OUT, f"bandit_{name}.npz"), action=actions, reward=np.array(rewards),
                        timestep=np.array(timesteps), arm0_mean=np.array(means0, np.float64),
                        reset_at=np.array(sorted(reset_at), np.int64), k=k, seed=seed,
                        numpy_version=np.__version__)
    print("bandit", name, "mean reward", float(np.mean(rewards)))


def pendulum_case(R, cls_name, name, kw, n_steps, action_seed, init_seed):
    cls = getattr(R, cls_name)
    dim = 4 if cls_name == "CartPole" else 2
    env = cls(**kw)
    rng = np.random.default_rng(action_seed)
    irng = np.random.default_rng(init_seed)
    wide = kw.pop("_wide_init", False) if "_wide_init" in kw else False
    rec = dict(pre=[], action=[], post=[], reward=[], done=[], trunc=[], acc=[], step=[])

    def new_init():
        if cls_name == "CartPole":
            return irng.uniform(-0.05, 0.05, dim)
        return irng.uniform([-np.pi, -20.0], [np.pi, 20.0]) if irng.random() < 0.5 else irng.uniform(-0.01, 0.01, dim)

    env.reset(initial_state=new_init())
    for _ in range(n_steps):
        a = rng.uniform(-3, 3, (1,)).astype(np.float32)
        pre = np.array(env.state, np.float64)
        obs, r, d, t, info = env.step(a)
        rec["pre"].append(pre); rec["action"].append(a[0]); rec["post"].append(obs); rec["reward"].append(r)
        rec["done"].append(d); rec["trunc"].append(t); rec["step"].append(info["steps"])
        rec["acc"].append([info["x_acc"], info["theta_acc"]] if cls_name == "CartPole" else [info["alpha_dot_dot"]])
        if d or t:
            env.reset(initial_state=new_init())
    out = {k: np.array(v) for k, v in rec.items()}
    out["config"] = json.dumps(kw)
    out["numpy_version"] = np.__version__
    np.savez_compressed(os.path.join(OUT, f"{cls_name.lower()}_{name}.npz"), **out)
    print(cls_name, name, "episodes", int((out["done"] | out["trunc"]).sum()), "done", int(out["done"].sum()))


def gambler_case(R, name, kw, n_steps, seed):
    """GamblersProblem trajectory under np.random.seed(seed): the uniform each step consumed is read from a clone of the
    global RandomState just before the step."""
    np.random.seed(seed)
    env = R.GamblersProblem(**kw)
    arng = np.random.default_rng(seed + 1000)
    rec = dict(pre=[], action=[], u=[], post=[], reward=[], done=[])
    for i in range(n_steps):
        pre = env.state
        # mostly legal bets, sometimes above the capital (the reference only asserts action <= goal_amount), and a few
        # steps taken from a finished state (the reference keeps stepping)
        hi = pre if arng.random() < 0.8 else kw.get("goal_amount", 100)
        a = int(arng.integers(0, max(hi, 0) + 1))
        clone = np.random.RandomState()
        clone.set_state(np.random.get_state())
        u = clone.random_sample()
        obs, r, d, t, _ = env.step(a)
        assert t is False
        rec["pre"].append(pre); rec["action"].append(a); rec["u"].append(u); rec["post"].append(obs)
        rec["reward"].append(r); rec["done"].append(d)
        if d and arng.random() < 0.7:
            env.reset()
    out = {k: np.array(v) for k, v in rec.items()}
    out["config"] = json.dumps(kw)
    out["numpy_version"] = np.__version__
    np.savez_compressed(os.path.join(OUT, f"gambler_{name}.npz"), **out)
    print("gambler", name, "done", int(out["done"].sum()), "wins", int(out["reward"].sum()))


def car_rental_case(R, name, kw, n_steps, seed):
    """JacksCarRental trajectory: construct, np.random.seed(seed), reset("equal") (draws nothing), then n_steps steps.
    A replay needs only RandomState(seed): every double the four scipy poisson.rvs calls of a step consume comes from it."""
    env = R.JacksCarRental(**kw)
    np.random.seed(seed)
    env.reset("equal")
    arng = np.random.default_rng(seed + 2000)
    n_act = 2 * kw.get("max_move_cars", 5) + 1
    rec = dict(pre=[], action=[], post=[], reward=[])
    for _ in range(n_steps):
        pre = env.state
        a = int(arng.integers(0, n_act))
        obs, r, d, t, _ = env.step(a)
        assert d is False and t is False
        rec["pre"].append(pre); rec["action"].append(a); rec["post"].append(obs); rec["reward"].append(r)
    out = {k: np.array(v) for k, v in rec.items()}
    out["config"] = json.dumps(kw)
    out["seed"] = seed
    out["numpy_version"] = np.__version__
    import scipy
    out["scipy_version"] = scipy.__version__
    small = R.JacksCarRental(max_cars=4, max_move_cars=2, request_lambda=[2, 2], return_lambda=[2, 2])
    mdp = small.build_mdp()                                      # test_car_rental.py fixture config (seconds)
    keys = sorted(mdp)
    out["mdp_keys"] = np.array([[k[0][0], k[0][1], k[1]] for k in keys])
    out["mdp_next"] = np.array([mdp[k][0] for k in keys])
    out["mdp_reward"] = np.array([mdp[k][1] for k in keys])
    np.savez_compressed(os.path.join(OUT, f"carrental_{name}.npz"), **out)
    print("carrental", name, "mean reward", float(np.mean(out["reward"])))


# ---- the reference driven by the ENGINE's Philox streams ("identical injected random draws and actions") -----------------
# These files let the GPU box compare the device DIRECTLY with outputs of the unmodified reference: env i of the batch is
# one reference instance whose random draws are replaced by the engine's Philox words of (seed, global env index, step) and
# whose actions are the engine's in-kernel random policy, so `Vec*.rollout(K)` / `step()` must reproduce them.
class _ChoiceFromEngine:
    """np_random stand-in for GridWorld: choice(n, p=probs) consumes the engine's transition word of the current step with
    numpy's own rule, searchsorted(cumsum(p) / cumsum(p)[-1], u, 'right') (Generator.choice), u = r * 2**-32."""

    def __init__(self):
        self.r = None

    def choice(self, n, p=None):
        # Generator.choice validates first: |sum(p) - 1| > sqrt(eps) raises before any draw (numpy _generator.pyx)
        if abs(math.fsum(p) - 1.0) > math.sqrt(np.finfo(np.float64).eps):
            raise ValueError("probabilities do not sum to 1")
        u = float(self.r) * 2.0 ** -32
        cdf = np.cumsum(np.asarray(p, np.float64))
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, u, side="right"))


def engine_gridworld_case(R, name, cfg, n_envs, n_steps, seed, env_offset):
    from oracle import philox
    kw = dict(cfg)
    kw["special_transitions"] = {(k[0], R.Action(k[1])): v for k, v in kw["special_transitions"].items()}
    g = np.arange(n_envs) + env_offset
    wt = np.stack([philox.group_word(seed, g, t, philox.STREAM_TRANSITION) for t in range(n_steps)])   # [T, N]
    wp = np.stack([philox.group_word(seed, g, t, philox.STREAM_POLICY) for t in range(n_steps)])
    act = (wp >> np.uint32(30)).astype(np.int64)
    obs = np.zeros((n_steps, n_envs, 2), np.int8)
    rew = np.zeros((n_steps, n_envs), np.float32)
    flags = np.zeros((n_steps, n_envs), np.uint8)        # bit0 done, bit1 truncated, bit2 the reference raised
    for i in range(n_envs):
        env = R.GridWorld(seed=0, **kw)
        shim = _ChoiceFromEngine()
        env.np_random = shim
        env.reset()
        for t in range(n_steps):
            shim.r = wt[t, i]
            try:
                o, r, d, tr, _ = env.step(int(act[t, i]))
            except ValueError:
                flags[t, i] = 4                              # state untouched, no reward
                obs[t, i] = env.state
                continue
            rew[t, i] = r
            flags[t, i] = (1 if d else 0) | (2 if tr else 0)
            if d or tr:
                o, _ = env.reset()                           # same-step auto-reset: next observation = start state
            obs[t, i] = o
    np.savez_compressed(os.path.join(OUT, f"engine_gridworld_{name}.npz"), obs=obs, reward=rew, flags=flags,
                        seed=seed, env_offset=env_offset, config=_jsonable_grid_cfg(cfg))
    print("engine gridworld", name, "episodes", int((flags & 3 > 0).sum()), "raised", int((flags & 4 > 0).sum()))


def engine_gambler_case(R, name, kw, n_envs, n_steps, seed, env_offset):
    from oracle import philox
    g = np.arange(n_envs) + env_offset
    goal = kw["goal_amount"]
    state = np.zeros((n_steps, n_envs), np.int32)
    rew = np.zeros((n_steps, n_envs), np.float32)
    done = np.zeros((n_steps, n_envs), bool)
    real_rand = np.random.rand
    try:
        for i in range(n_envs):
            env = R.GamblersProblem(**kw)
            for t in range(n_steps):
                r32 = int(philox.group_word(seed, g[i:i + 1], t, philox.STREAM_TRANSITION)[0])
                wp = int(philox.group_word(seed, g[i:i + 1], t, philox.STREAM_POLICY)[0])
                hi = max(min(env.state, goal - env.state), 0)
                bet = (wp * (hi + 1)) >> 32                  # the engine's random policy: uniform legal bet
                np.random.rand = lambda r32=r32: r32 * 2.0 ** -32
                s, r, d, _, _ = env.step(int(bet))
                rew[t, i], done[t, i] = r, d
                if d:
                    s, _ = env.reset()
                state[t, i] = s
    finally:
        np.random.rand = real_rand
    np.savez_compressed(os.path.join(OUT, f"engine_gambler_{name}.npz"), state=state, reward=rew, done=done, seed=seed,
                        env_offset=env_offset, config=json.dumps(kw))
    print("engine gambler", name, "episodes", int(done.sum()), "wins", int(rew.sum()))


def engine_car_rental_case(R, name, kw, n_envs, n_steps, seed, env_offset):
    """scipy.stats.poisson.rvs is replaced by numpy's legacy Poisson algorithm (oracle.acml.legacy_poisson, bit-exact vs
    RandomState.poisson) over the engine's Philox doubles of (env, step)."""
    import scipy.stats
    from oracle import acml, philox
    cr_mod = __import__("rl_envs_forge.envs.acml.car_rental.car_rental", fromlist=["poisson"])
    n_act = 2 * kw.get("max_move_cars", 5) + 1
    g = np.arange(n_envs) + env_offset
    state = np.zeros((n_steps, n_envs, 2), np.int32)
    rew = np.zeros((n_steps, n_envs), np.float32)
    real = cr_mod.poisson

    class _Rvs:
        src = None

        def rvs(self, lam):
            return acml.legacy_poisson(self.src, lam)

        def pmf(self, k, lam):
            return scipy.stats.poisson.pmf(k, lam)
    fake = _Rvs()
    try:
        cr_mod.poisson = fake
        for i in range(n_envs):
            env = R.JacksCarRental(init_state_option="equal", **kw)
            for t in range(n_steps):
                wp = int(philox.group_word(seed, g[i:i + 1], t, philox.STREAM_POLICY)[0])
                fake.src = acml.PhiloxDoubles(seed, int(g[i]), t)
                s, r, _, _, _ = env.step(int((wp * n_act) >> 32))
                state[t, i], rew[t, i] = s, r
    finally:
        cr_mod.poisson = real
    np.savez_compressed(os.path.join(OUT, f"engine_carrental_{name}.npz"), state=state, reward=rew, seed=seed,
                        env_offset=env_offset, config=json.dumps(kw))
    print("engine carrental", name, "mean reward", float(rew.mean()))


def engine_bandit_case(R, name, k, arm_params, n_envs, n_steps, seed, env_offset):
    """Arm.sample draws from `np_random`; here that is oracle.bandit.LegacyDistributions (bit-exact vs RandomState) over the
    engine's per-(bandit, timestep) Philox primitives, so the float64 rewards are what the reference returns for the
    engine's draws.  Custom arms only (default arms draw their means in __init__ from the RandomState)."""
    from oracle import bandit as OB, philox
    g = np.arange(n_envs) + env_offset
    rew = np.zeros((n_steps, n_envs), np.float64)
    act = np.zeros((n_steps, n_envs), np.int32)
    for i in range(n_envs):
        env = R.KArmedBandit(k=k, arm_params={a: dict(p) for a, p in arm_params.items()}, seed=0)
        for t in range(n_steps):
            wp = int(philox.group_word(seed, g[i:i + 1], t, philox.STREAM_POLICY)[0])
            a = (wp * k) >> 32
            env.np_random = OB.LegacyDistributions(OB.PhiloxPrimitives(seed, int(g[i]), t))
            for arm in env.arms:                             # arms sample through the env's generator
                if hasattr(arm, "np_random"):
                    arm.np_random = env.np_random
            _, r, _, _, _ = env.step(int(a))
            rew[t, i], act[t, i] = r, a
    np.savez_compressed(os.path.join(OUT, f"engine_bandit_{name}.npz"), reward=rew, action=act, seed=seed,
                        env_offset=env_offset, k=k)
    print("engine bandit", name, "mean reward", float(rew.mean()))


def engine_cases(R):
    engine_gridworld_case(R, "config2_p08", dict(CONFIG2_LAYOUT, p_success=0.8), 384, 256, seed=41, env_offset=1 << 20)
    engine_gridworld_case(R, "config2_p1", dict(CONFIG2_LAYOUT, p_success=1.0), 128, 256, seed=42, env_offset=7)
    engine_gambler_case(R, "default", dict(goal_amount=100, win_probability=0.4, start_capital=10), 256, 200, seed=43,
                        env_offset=1000)
    engine_car_rental_case(R, "default", dict(), 128, 60, seed=44, env_offset=12345)
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items() if "distribution" in p}
    engine_bandit_case(R, "config3_custom_arms", len(arms), arms, 128, 30, seed=45, env_offset=99)


def labyrinth_case(R, name, kw, n_steps, action_seed):
    """Labyrinth trajectory on a maze generated by the reference: every step's pre/post position, reward, done and the
    full observation are recorded.  A finished episode continues from the start position on the SAME maze
    (`set_state`; the reference's own reset() would generate a new maze, which is host work outside the stepping path)."""
    env = R.Labyrinth(**kw)
    rng = np.random.default_rng(action_seed)
    grid = np.array(env.maze.grid, np.uint8)
    start, target = tuple(env.maze.start_position), tuple(env.maze.target_position)
    free = [(r, c) for r in range(env.rows) for c in range(env.cols) if grid[r, c] != 0 and (r, c) != target]
    rec = dict(pre=[], action=[], post=[], reward=[], done=[], obs=[], info=[])
    for i in range(n_steps):
        if i % 40 == 39:                                   # visit the whole maze, the target's neighbourhood included
            near = [p for p in free if abs(p[0] - target[0]) + abs(p[1] - target[1]) == 1]
            env.set_state(near[rng.integers(len(near))] if near and rng.random() < 0.5 else free[rng.integers(len(free))])
        pre = env.player.position
        a = int(rng.integers(0, 4))
        obs, r, d, t, info = env.step(a)
        assert t is False
        rec["pre"].append(pre); rec["action"].append(a); rec["post"].append(env.player.position)
        rec["reward"].append(r); rec["done"].append(d); rec["obs"].append(np.array(obs, np.uint8))
        rec["info"].append({"": 0, "Invalid move!": 1, "Reached the target!": 2}[info.get("info", "")])
        if d:
            env.set_state(start)
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(grid=grid, start=np.array(start), target=np.array(target), config=json.dumps(kw),
               initial_obs=np.array(R.Labyrinth(**kw).state, np.uint8))
    np.savez_compressed(os.path.join(OUT, f"labyrinth_{name}.npz"), **out)
    print("labyrinth", name, grid.shape, "done", int(out["done"].sum()), "bumps", int((out["info"] == 1).sum()))


# "donut" rooms are left out: room.py:228 calls random.randint with a float bound, a TypeError on Python >= 3.12
_ROOMS = ["rectangle", "oval", "t_shape", "l_shape", "triangle"]


def labyrinth_cases(R):
    labyrinth_case(R, "10x10_full", dict(rows=10, cols=10, seed=1, room_types=_ROOMS), 1500, 1)
    labyrinth_case(R, "20x20_full", dict(rows=20, cols=20, seed=42, room_types=_ROOMS), 1500, 2)                # test_labyrinth.py:222-233
    labyrinth_case(R, "12x17_vision2", dict(rows=12, cols=17, seed=7, state_vision_range=2, room_types=_ROOMS), 1500, 3)
    labyrinth_case(R, "30x30_vision3", dict(rows=30, cols=30, seed=42, state_vision_range=3, room_types=_ROOMS), 1500, 4)   # :215-220
    labyrinth_case(R, "16x16_vision1", dict(rows=16, cols=16, seed=9, state_vision_range=1, room_types=_ROOMS), 800, 5)


def acml_cases(R):
    gambler_case(R, "default", dict(goal_amount=100, win_probability=0.4, start_capital=10), 4000, seed=21)  # test fixture
    gambler_case(R, "fair_small", dict(goal_amount=16, win_probability=0.5, start_capital=3), 3000, seed=22)
    gambler_case(R, "sure_win", dict(goal_amount=30, win_probability=1.0, start_capital=1), 300, seed=23)
    car_rental_case(R, "default", dict(), 3000, seed=31)
    car_rental_case(R, "small", dict(max_cars=4, max_move_cars=2, request_lambda=[2, 2], return_lambda=[2, 2]), 3000, seed=32)
    car_rental_case(R, "zero_lambda", dict(max_cars=10, max_move_cars=3, request_lambda=[0, 9.5], return_lambda=[1, 0],
                                           rental_credit=7, move_car_cost=3), 2000, seed=33)


def transition_probs_cases(R):
    """`GridWorld.transition_probs` (grid_world.py:82, :401-430) of the unmodified reference, densified: one file, per case
    the constructor config, the slip distribution, the rows present and tp[row = cell * 4 + action, next cell]."""
    crng = np.random.default_rng(77)
    cases = [("default", dict(), None), ("slip_half", dict(p_success=0.5), None),
             ("config2_p08", dict(CONFIG2_LAYOUT, p_success=0.8), None),
             ("config2_custom_slip", dict(CONFIG2_LAYOUT, p_success=0.7), {0: 0.5, 1: 0.25, 2: 0.25, 3: 0.0}),
             ("p0", dict(rows=4, cols=6, terminal_states={(3, 5): 2.0}, p_success=0.0), None)]
    cases += [(f"random{i:02d}", random_grid_cfg(crng), None) for i in range(8)]
    out = {"names": json.dumps([c[0] for c in cases])}
    for name, cfg, slip in cases:
        kw = dict(cfg)
        if kw.get("special_transitions"):
            kw["special_transitions"] = {(k[0], R.Action(k[1])): v for k, v in kw["special_transitions"].items()}
        if slip is not None:
            kw["slip_distribution"] = {R.Action(a): w for a, w in slip.items()}
        env = R.GridWorld(seed=0, **kw)
        S = env.rows * env.cols
        present = np.zeros(S * 4, bool)
        tp = np.zeros((S * 4, S), np.float64)
        for (st, a), outs in env.transition_probs.items():
            row = (st[0] * env.cols + st[1]) * 4 + int(a)
            present[row] = True
            for to, pr in outs.items():
                tp[row, to[0] * env.cols + to[1]] = pr
        out[name + "_config"] = _jsonable_grid_cfg(cfg)
        out[name + "_slip"] = json.dumps(None if slip is None else {str(a): w for a, w in slip.items()})
        out[name + "_present"] = present
        out[name + "_tp"] = tp
    out["numpy_version"] = np.__version__
    np.savez_compressed(os.path.join(OUT, "tprobs_cases.npz"), **out)
    print("transition_probs:", len(cases), "cases")


def main():
    os.makedirs(OUT, exist_ok=True)
    R = load_reference()
    import sys
    if "tprobs" in sys.argv[1:]:
        transition_probs_cases(R)
        return
    if "engine_big" in sys.argv[1:]:
        # a 1M-step subsample of BASELINE config 2 (reference under the shim on a subsample of the population):
        # 4096 envs x 256 steps of the unmodified reference under the engine's streams, limit 200 so that truncation occurs
        engine_gridworld_case(R, "config2_p08_4096", dict(CONFIG2_LAYOUT, p_success=0.8, episode_length_limit=200), 4096, 256,
                              seed=20261019, env_offset=3 << 18)
        return
    if "engine" in sys.argv[1:]:
        engine_cases(R)
        return
    if "acml" in sys.argv[1:] or "labyrinth" in sys.argv[1:]:
        if "acml" in sys.argv[1:]:
            acml_cases(R)
        if "labyrinth" in sys.argv[1:]:
            labyrinth_cases(R)
        return
    # -- GridWorld ---------------------------------------------------------------------------------------
    gridworld_case(R, "default", dict(), 3000, seed=0, action_seed=0)                        # BASELINE config 1
    gridworld_case(R, "slip_seed42", dict(p_success=0.5), 3000, seed=42, action_seed=1)       # test_grid_world.py:20-22
    gridworld_case(R, "config2_p1", dict(CONFIG2_LAYOUT, p_success=1.0), 6000, seed=3, action_seed=2)
    gridworld_case(R, "config2_p08", dict(CONFIG2_LAYOUT, p_success=0.8), 6000, seed=4, action_seed=3)
    gridworld_case(R, "limit10", dict(episode_length_limit=10), 200, seed=5, action_seed=4)  # test_grid_world.py:154-166
    gridworld_case(R, "runtime_special", dict(), 1500, seed=6, action_seed=5,
                   runtime_specials=[((0, 0), 2, (4, 4), 0.5), ((2, 2), 1, None, None)])      # test_grid_world.py:50-67
    gridworld_case(R, "p0", dict(rows=4, cols=6, terminal_states={(3, 5): 2.0}, p_success=0.0), 1500, seed=7, action_seed=6)
    crng = np.random.default_rng(2024)
    for i in range(12):
        gridworld_case(R, f"random{i:02d}", random_grid_cfg(crng), 800, seed=100 + i, action_seed=200 + i)
    # -- KArmedBandit ------------------------------------------------------------------------------------
    bandit_case(R, "default_seed0", 10, None, 0, 500, action_seed=0)                          # README known answer
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}
    arms[0]["param_functions"] = bandit_shift_functions()
    bandit_case(R, "config3", 10, arms, 11, 90, action_seed=1, reset_at=(40,))
    arms_lo = {0: dict(distribution="gamma", shape=0.4, scale=2.0), 1: dict(distribution="beta", a=0.5, b=0.7),
               2: dict(distribution="gamma", shape=1.0, scale=3.0), 3: dict(distribution="weibull", a=0.7)}
    bandit_case(R, "small_shapes", 4, arms_lo, 5, 1500, action_seed=2)
    # -- pendulums ---------------------------------------------------------------------------------------
    pendulum_case(R, "CartPole", "default", dict(), 3000, 0, 1)
    pendulum_case(R, "CartPole", "continuous", dict(continuous_reward=True, tau=0.01, episode_length_limit=50), 1500, 2, 3)
    pendulum_case(R, "CartPole", "nonlinear_angle", dict(continuous_reward=True, nonlinear_reward=True,
                                                          reward_decay_rate=30.0, angle_termination=0.1), 1500, 4, 5)
    pendulum_case(R, "PendulumDisk", "default", dict(), 3000, 6, 7)
    pendulum_case(R, "PendulumDisk", "continuous", dict(continuous_reward=True, tau=0.01, episode_length_limit=50), 1500, 8, 9)
    pendulum_case(R, "PendulumDisk", "nonlinear_angle", dict(continuous_reward=True, nonlinear_reward=True,
                                                              reward_decay_rate=30.0, angle_termination=2.0), 1500, 10, 11)
    # -- ACML envs ------------------------------------------------------------------
    acml_cases(R)
    # -- Labyrinth stepping on reference-generated mazes ------------------------------
    labyrinth_cases(R)
    # -- the reference driven by the engine's own Philox streams (direct device-vs-reference vectors) ---------------------
    engine_cases(R)
    # -- GridWorld.transition_probs export ----------------------------------------------------------
    transition_probs_cases(R)


if __name__ == "__main__":
    main()
