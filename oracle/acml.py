"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference ACML envs' hot path.

Follows mariusdgm/rl-envs-forge v5.22.0:
    GamblersProblem  rl_envs_forge/envs/acml/gambler_problem/gambler_problem.py  step :36-66, reset :68-76,
                     build_mdp :89-119
    JacksCarRental   rl_envs_forge/envs/acml/car_rental/car_rental.py  _move_cars_and_calculate_cost :54-74,
                     _simulate_rentals_and_returns :76-104, step :106-121, reset :123-143

Randomness that lives in third-party code NOT under /root/reference:
  * `np.random.rand()` (gambler_problem.py:51): one double from numpy's global legacy RandomState (MT19937).
  * `scipy.stats.poisson.rvs(lam)` (car_rental.py:89-92) -> the global RandomState's `poisson`, i.e. numpy's legacy
    `random_poisson`: lam == 0 -> 0 without a draw; lam < 10 -> multiplication method (X = number of uniforms whose
    running product stays above exp(-lam)); lam >= 10 -> PTRS, Hoermann's transformed rejection with numpy's own
    `random_loggam` (two doubles per trial).
    numpy pinned 1.26.4 (poetry.lock:786-787), scipy from the same lock; the container has numpy 2.3 / scipy 1.18,
    whose legacy stream is frozen by numpy's compatibility policy.
  * `np.random.randint(0, max_cars + 1)` (car_rental.py:136-139): initial state; parity tests inject initial states.

Primitive sources (as in oracle/bandit.py): a real `RandomState` (pins the restatement against numpy itself and against
golden runs of the reference with `np.random.seed`), or the engine's Philox words (what the device is compared to).
The engine draws a Poisson double from TWO Philox words exactly like numpy's legacy `next_double` draws it from two
MT19937 outputs: ((a >> 5) * 2**26 + (b >> 6)) / 2**53.

Parity status: PINNED -- tests/test_oracle_acml.py (RandomState.poisson bit-exact; golden trajectories of the unmodified
reference, tests/golden/gambler_*.npz and carrental_*.npz).  Nothing in the product package may import this module.
"""
import math

import numpy as np

from . import philox

STREAM_ACML = 6  # per (env, t), block j: the Poisson uniforms of JacksCarRental (two words per double)


# -- primitive sources ------------------------------------------------------------------------------------------------
class MTDoubles:
    """next_double of a real numpy RandomState (the reference's global generator)."""

    def __init__(self, rs):
        self.rs = rs

    def double(self):
        return float(self.rs.random_sample())


class PhiloxDoubles:
    """The engine's per-(env, step) double stream: words taken in order from STREAM_ACML blocks 0, 1, ..."""

    def __init__(self, seed, g, t):
        self.seed, self.g, self.t = seed, int(g), int(t)
        self._words = []
        self.consumed = 0

    def _word(self):
        j, k = divmod(self.consumed, 4)
        while len(self._words) <= j:
            blk = philox.draw_block(self.seed, self.g, self.t, STREAM_ACML, block=len(self._words))
            self._words.append([int(w) for w in blk])
        self.consumed += 1
        return self._words[j][k]

    def double(self):
        a, b = self._word() >> 5, self._word() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0


_LOGGAM_A = (8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04, -5.952380952380952e-04,
             8.417508417508418e-04, -1.917526917526918e-03, 6.410256410256410e-03, -2.955065359477124e-02,
             1.796443723688307e-01, -1.39243221690590e+00)


def random_loggam(x):
    """numpy's random_loggam (distributions.c): log-gamma by an asymptotic series after shifting x to >= 7."""
    if x == 1.0 or x == 2.0:
        return 0.0
    n = int(7 - x) if x < 7.0 else 0
    x0 = x + n
    x2 = (1.0 / x0) * (1.0 / x0)
    gl0 = _LOGGAM_A[9]
    for k in range(8, -1, -1):
        gl0 *= x2
        gl0 += _LOGGAM_A[k]
    gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * math.log(x0) - x0
    if x < 7.0:
        for _ in range(n):
            gl -= math.log(x0 - 1.0)
            x0 -= 1.0
    return gl


def ptrs_constants(lam):
    """Per-rate constants of numpy's random_poisson_ptrs (Hoermann's transformed rejection), float64 like the C code."""
    slam = math.sqrt(lam)
    b = 0.931 + 2.53 * slam
    a = -0.059 + 0.02483 * b
    invalpha = 1.1239 + 1.1328 / (b - 3.4)
    vr = 0.9277 - 3.6224 / (b - 2)
    return dict(loglam=math.log(lam), a=a, b=b, log_invalpha=math.log(invalpha), vr=vr)


def legacy_poisson(src, lam):
    """numpy legacy random_poisson, float64 like the C code: lam == 0 -> 0 without a draw; lam < 10 -> multiplication
    method; lam >= 10 -> PTRS (two doubles per trial)."""
    if lam >= 10:
        c = ptrs_constants(lam)
        while True:
            u = src.double() - 0.5
            v = src.double()
            us = 0.5 - abs(u)
            if us == 0.0:             # u == -0.5: 2a/us is +inf in C, k = floor(-inf) < 0 -> next trial
                continue
            k = math.floor((2 * c["a"] / us + c["b"]) * u + lam + 0.43)
            if us >= 0.07 and v <= c["vr"]:
                return k
            if k < 0 or (us < 0.013 and v > us):
                continue
            lv = math.log(v) if v > 0.0 else -math.inf
            if lv + c["log_invalpha"] - math.log(c["a"] / (us * us) + c["b"]) <= -lam + k * c["loglam"] - random_loggam(k + 1):
                return k
    if lam == 0:
        return 0
    enlam = math.exp(-lam)
    x, prod = 0, 1.0
    while True:
        prod *= src.double()
        if prod > enlam:
            x += 1
        else:
            return x


# -- GamblersProblem -----------------------------------------------------------------------------------------------
def gambler_step(state, action, u, goal_amount, win_probability):
    """One reference step (gambler_problem.py:36-66) given the uniform `u` np.random.rand() returned.
    -> (state, reward, done).  The reference asserts 0 <= action <= goal_amount and nothing else (a bet may exceed the
    capital; a finished env keeps stepping from 0 / goal_amount)."""
    assert 0 <= action <= goal_amount, "Invalid action"
    done, reward = False, 0
    if u < win_probability:
        state += action
        if state >= goal_amount:
            state, reward, done = goal_amount, 1, True
    else:
        state -= action
        if state <= 0:
            state, done = 0, True
    return state, reward, done


def gambler_step_vec(state, action, r32, goal_amount, win_probability):
    """Vectorised gambler_step with u = r32 * 2**-32 (exact in float64).  Invalid actions leave the env untouched and are
    flagged.  -> (state, reward f32, done, invalid)."""
    state = np.asarray(state, np.int64).copy()
    action = np.asarray(action, np.int64)
    u = np.asarray(r32, np.uint32).astype(np.float64) * 2.0 ** -32
    invalid = (action < 0) | (action > goal_amount)
    win = u < win_probability
    nxt = np.where(win, state + action, state - action)
    done = np.where(win, nxt >= goal_amount, nxt <= 0)
    reward = (win & (nxt >= goal_amount)).astype(np.float32)
    nxt = np.where(win, np.minimum(nxt, goal_amount), np.maximum(nxt, 0))
    state = np.where(invalid, state, nxt)
    return state, np.where(invalid, np.float32(0), reward), done & ~invalid, invalid


def gambler_build_mdp(goal_amount, win_probability):
    """build_mdp (gambler_problem.py:89-119)."""
    d = {}
    for s in range(1, goal_amount):
        for a in range(1, min(s, goal_amount - s) + 1):
            win_state = min(s + a, goal_amount)
            lose_state = max(s - a, 0)
            d[(s, a)] = {"win": (win_state, 1 if win_state == goal_amount else 0, win_state == goal_amount, win_probability),
                         "loss": (lose_state, 0, lose_state == 0, 1 - win_probability)}
    return d


# -- JacksCarRental ------------------------------------------------------------------------------------------------
class CarRentalParams:
    def __init__(self, max_cars=20, max_move_cars=5, rental_credit=10, move_car_cost=2, request_lambda=(3, 4),
                 return_lambda=(3, 2)):
        self.max_cars, self.max_move_cars = max_cars, max_move_cars
        self.rental_credit, self.move_car_cost = rental_credit, move_car_cost
        self.request_lambda, self.return_lambda = tuple(request_lambda), tuple(return_lambda)


def car_rental_move(p, state, action):
    """_move_cars_and_calculate_cost (car_rental.py:54-74); `action` already shifted to [-max_move, max_move]."""
    a, b = state
    moved = min(a if action > 0 else b, abs(action), p.max_move_cars)
    sign = (action > 0) - (action < 0)
    a -= moved * sign
    b += moved * sign
    return (max(a, 0), max(b, 0)), abs(action) * p.move_car_cost


def car_rental_step(p, state, action, src):
    """One reference step (car_rental.py:106-121): raw action in [0, 2*max_move]; four Poisson draws from `src` in the
    reference's order (requests A, requests B, returns A, returns B).  -> (state, reward)."""
    moved, cost = car_rental_move(p, state, action - p.max_move_cars)
    a, b = moved
    req_a = legacy_poisson(src, p.request_lambda[0])
    req_b = legacy_poisson(src, p.request_lambda[1])
    ret_a = legacy_poisson(src, p.return_lambda[0])
    ret_b = legacy_poisson(src, p.return_lambda[1])
    rent_a, rent_b = min(a, req_a), min(b, req_b)
    reward = (rent_a + rent_b) * p.rental_credit
    a = min(a - rent_a + ret_a, p.max_cars)
    b = min(b - rent_b + ret_b, p.max_cars)
    return (a, b), reward - cost


def car_rental_engine_step(p, seed, g, t, state, action):
    """What the device must produce for env g at step t: the reference step fed by the engine's Philox doubles."""
    return car_rental_step(p, (int(state[0]), int(state[1])), int(action), PhiloxDoubles(seed, g, t))


def car_rental_engine_step_vec(p, seed, g, t, state, action):
    """Vectorised car_rental_engine_step over envs g (same arithmetic; checked against the scalar version in the tests).
    state int [n,2], action int [n] -> (state int64 [n,2], reward int64 [n])."""
    g = np.asarray(g, np.uint64)
    n = g.shape[0]
    st = np.asarray(state, np.int64)
    act = np.asarray(action, np.int64) - p.max_move_cars
    a, b = st[:, 0].copy(), st[:, 1].copy()
    moved = np.minimum(np.minimum(np.where(act > 0, a, b), np.abs(act)), p.max_move_cars)
    sign = np.sign(act)
    a = np.maximum(a - moved * sign, 0)
    b = np.maximum(b + moved * sign, 0)
    cost = np.abs(act) * p.move_car_cost
    cursor = np.zeros(n, np.int64)      # words consumed per env
    blocks = []                         # lazily grown list of [n,4] word blocks

    def words():                        # one word per env, each at its own cursor
        j, lane = cursor // 4, cursor % 4
        while len(blocks) <= int(j.max()):
            blocks.append(np.stack(philox.draw_block(seed, g, t, STREAM_ACML, block=len(blocks)), axis=-1).astype(np.uint64))
        allw = np.stack(blocks, axis=1)                              # [n, nblocks, 4]
        w = allw[np.arange(n), j, lane]
        return w

    def doubles(active):
        nonlocal cursor
        wa = words()
        cursor = cursor + active
        wb = words()
        cursor = cursor + active
        return ((wa >> np.uint64(5)).astype(np.float64) * 67108864.0 + (wb >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0

    def poisson_ptrs(lam):
        c = ptrs_constants(lam)
        out = np.zeros(n, np.int64)
        active = np.ones(n, bool)
        while active.any():
            u = doubles(active.astype(np.int64)) - 0.5
            v = doubles(active.astype(np.int64))
            us = 0.5 - np.abs(u)
            with np.errstate(divide="ignore", invalid="ignore"):
                k = np.floor((2 * c["a"] / us + c["b"]) * u + lam + 0.43)
                quick = (us >= 0.07) & (v <= c["vr"])
                skip = (k < 0) | ((us < 0.013) & (v > us)) | ~np.isfinite(k)
                kk = np.where(skip | quick, 1.0, k)
                lg = np.array([random_loggam(float(x) + 1.0) for x in kk])
                slow = (np.log(v) + c["log_invalpha"] - np.log(c["a"] / (us * us) + c["b"])) <= (-lam + kk * c["loglam"] - lg)
            acc = active & (quick | (~skip & slow))
            out[acc] = k[acc].astype(np.int64)
            active = active & ~acc
        return out

    def poisson(lam):
        if lam >= 10:
            return poisson_ptrs(lam)
        x = np.zeros(n, np.int64)
        if lam == 0:
            return x
        enlam = math.exp(-lam)
        prod = np.ones(n)
        active = np.ones(n, bool)
        while active.any():
            u = doubles(active.astype(np.int64))
            prod = np.where(active, prod * u, prod)
            more = active & (prod > enlam)
            x += more
            active = more
        return x

    req_a, req_b = poisson(p.request_lambda[0]), poisson(p.request_lambda[1])
    ret_a, ret_b = poisson(p.return_lambda[0]), poisson(p.return_lambda[1])
    rent_a, rent_b = np.minimum(a, req_a), np.minimum(b, req_b)
    reward = (rent_a + rent_b) * p.rental_credit - cost
    a = np.minimum(a - rent_a + ret_a, p.max_cars)
    b = np.minimum(b - rent_b + ret_b, p.max_cars)
    return np.stack([a, b], axis=1), reward


def car_rental_expected_reward_mdp(p, poisson_max_value=12):
    """build_mdp (car_rental.py:229-301) restated with numpy sums: {(state, action): (next_state for zero requests and
    returns, expected reward, False)}.  The reference accumulates prob_total * reward term by term; the sums here differ
    from it only by floating-point association."""
    from scipy.stats import poisson as sp
    n = p.max_cars + 1
    if poisson_max_value < p.max_cars:
        raise IndexError("list index out of range")   # the reference indexes poisson_probs[...][request] up to max_cars
    pr = [[sp.pmf(i, lam) for i in range(n)] for lam in p.request_lambda]
    pt = [[sp.pmf(i, lam) for i in range(n)] for lam in p.return_lambda]
    ret_mass = float(np.sum(np.outer(pt[0], pt[1])))
    out = {}
    for a0 in range(n):
        for b0 in range(n):
            for action in range(-p.max_move_cars, p.max_move_cars + 1):
                (a, b), cost = car_rental_move(p, (a0, b0), action)
                ra = np.minimum(a, np.arange(n))[:, None]
                rb = np.minimum(b, np.arange(n))[None, :]
                reward = (ra + rb) * p.rental_credit - cost
                exp_r = float(np.sum(np.outer(pr[0], pr[1]) * reward)) * ret_mass
                out[((a0, b0), action)] = ((min(a, p.max_cars), min(b, p.max_cars)), exp_r, False)
    return out
