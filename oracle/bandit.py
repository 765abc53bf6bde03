"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the reference KArmedBandit hot path.

Follows mariusdgm/rl-envs-forge v5.22.0, rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py:
    Arm.update_param :53-64, Arm.sample :66-122, _init_arms :155-173, step :175-200, reset :202-217.

The sampling arithmetic itself lives in a third-party dependency that is NOT under /root/reference:
numpy's legacy `RandomState` distributions (numpy pinned 1.26.4, poetry.lock:786-787; call sites
k_armed_bandit.py:79,84,88,97,106,111,117,122,162).  `LegacyDistributions` restates those published algorithms
(polar Box-Muller with a one-value cache, -log(1-U) exponential, Marsaglia-Tsang / Ahrens-Dieter gamma,
Johnk / two-gamma beta, inverse-cdf logistic and Weibull) over an abstract *primitive source* that yields
U(0,1) doubles (`u()`) and N(0,1) doubles (`n()`):

  * `MTPrimitives(RandomState)`  -- U = RandomState.random_sample(), N = legacy polar gauss built on those U.
        With it the restatement is bit-exact against numpy's own RandomState.normal/gamma/beta/... ; that is the
        pin (tests/test_oracle_bandit.py, tests/golden/bandit_*.npz, README known answer 0.5442007795281013).
  * `PhiloxPrimitives`           -- U = u23(word), N = plain Box-Muller of two words, words taken in order from the
        engine's STREAM_BANDIT Philox substream of (env, timestep).  Same consumption order as the device kernel
        (csrc/bandit.cu), evaluated in float64.  This is what the device rewards are compared against.

Parity status: PINNED (see above).  Nothing in the product package may import this module.
"""
import math

import numpy as np

from . import philox

FAMILIES = ("normal", "uniform", "exponential", "gamma", "beta", "logistic", "weibull", "lognormal")


# -- primitive sources -------------------------------------------------------------------------------------
class MTPrimitives:
    """U from a real numpy RandomState; N by the legacy polar method with its one-value cache."""

    def __init__(self, rs):
        self.rs = rs
        self._cached = None

    def u(self):
        return float(self.rs.random_sample())

    def n(self):
        if self._cached is not None:
            v, self._cached = self._cached, None
            return v
        while True:
            x1 = 2.0 * self.u() - 1.0
            x2 = 2.0 * self.u() - 1.0
            r2 = x1 * x1 + x2 * x2
            if 0.0 < r2 < 1.0:
                break
        f = math.sqrt(-2.0 * math.log(r2) / r2)
        self._cached = f * x1
        return f * x2


class PhiloxPrimitives:
    """Scalar view of the engine's per-(env, timestep) primitive stream (float64 evaluation)."""

    def __init__(self, seed, g, t):
        self.seed, self.g, self.t = seed, g, t
        self._words = []
        self.consumed = 0

    def _word(self):
        j, k = divmod(self.consumed, 4)
        while len(self._words) <= j:
            blk = philox.draw_block(self.seed, self.g, self.t, philox.STREAM_BANDIT, block=len(self._words))
            self._words.append([int(w) for w in blk])
        self.consumed += 1
        return self._words[j][k]

    def u(self):
        return float(philox.u23(self._word()))

    def n(self):
        u1 = float(philox.u23(self._word()))
        u2 = float(philox.u23(self._word()))
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


# -- numpy legacy distributions, restated ------------------------------------------------------------------
class LegacyDistributions:
    """The subset of RandomState's API that Arm.sample calls, over a primitive source."""

    def __init__(self, prim):
        self.p = prim

    def random_sample(self):
        return self.p.u()

    def randn(self, k):
        return np.array([self.p.n() for _ in range(k)])

    def normal(self, loc=0.0, scale=1.0):
        if scale < 0:
            raise ValueError("scale < 0")
        return loc + scale * self.p.n()

    def lognormal(self, mean=0.0, sigma=1.0):
        if sigma < 0:
            raise ValueError("sigma < 0")
        return math.exp(mean + sigma * self.p.n())

    def uniform(self, low=0.0, high=1.0):
        return low + (high - low) * self.p.u()

    def standard_exponential(self):
        return -math.log(1.0 - self.p.u())

    def exponential(self, scale=1.0):
        if scale < 0:
            raise ValueError("scale < 0")
        return scale * self.standard_exponential()

    def logistic(self, loc=0.0, scale=1.0):
        if scale < 0:
            raise ValueError("scale < 0")
        u = self.p.u()
        while u <= 0.0:
            u = self.p.u()
        return loc + scale * math.log(u / (1.0 - u))

    def weibull(self, a):
        if a < 0:
            raise ValueError("a < 0")
        if a == 0.0:
            return 0.0
        return math.pow(self.standard_exponential(), 1.0 / a)

    def standard_gamma(self, shape):
        if shape == 1.0:
            return self.standard_exponential()
        if shape == 0.0:
            return 0.0
        if shape < 1.0:
            while True:
                u = self.p.u()
                v = self.standard_exponential()
                if u <= 1.0 - shape:
                    x = math.pow(u, 1.0 / shape)
                    if x <= v:
                        return x
                else:
                    y = -math.log((1.0 - u) / shape)
                    x = math.pow(1.0 - shape + shape * y, 1.0 / shape)
                    if x <= v + y:
                        return x
        b = shape - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * b)
        while True:
            while True:
                x = self.p.n()
                v = 1.0 + c * x
                if v > 0.0:
                    break
            v = v * v * v
            u = self.p.u()
            if u < 1.0 - 0.0331 * (x * x) * (x * x):
                return b * v
            if math.log(u) < 0.5 * x * x + b * (1.0 - v + math.log(v)):
                return b * v

    def gamma(self, shape, scale=1.0):
        if shape < 0:
            raise ValueError("shape < 0")
        if scale < 0:
            raise ValueError("scale < 0")
        return scale * self.standard_gamma(shape)

    def beta(self, a, b):
        if a <= 0:
            raise ValueError("a <= 0")
        if b <= 0:
            raise ValueError("b <= 0")
        if a <= 1.0 and b <= 1.0:
            while True:
                u = self.p.u()
                v = self.p.u()
                x = math.pow(u, 1.0 / a)
                y = math.pow(v, 1.0 / b)
                if x + y <= 1.0 and u + v > 0.0:
                    if x + y > 0:
                        return x / (x + y)
                    lx, ly = math.log(u) / a, math.log(v) / b
                    lm = max(lx, ly)
                    lx -= lm
                    ly -= lm
                    return math.exp(lx - math.log(math.exp(lx) + math.exp(ly)))
        ga = self.standard_gamma(a)
        gb = self.standard_gamma(b)
        return ga / (ga + gb)


# -- Arm / KArmedBandit, restated ----------------------------------------------------------------------------
class ArmPort:
    def __init__(self, distribution="normal", param_functions=None, **kwargs):
        if distribution not in FAMILIES:
            raise ValueError(f"Invalid distribution: {distribution}. Please use one of {list(FAMILIES)}.")
        self.distribution = distribution
        self.init_params = dict(kwargs)
        self.current_params = dict(kwargs)
        self.param_functions = param_functions or []

    def update_param(self, timestep):  # k_armed_bandit.py:53-64 -- absolute shift from the initial value
        for fd in self.param_functions:
            self.current_params[fd["target_param"]] = self.init_params[fd["target_param"]] + fd["function"](timestep)

    def sample(self, rng):  # k_armed_bandit.py:66-122
        p, d = self.current_params, self.distribution
        if d == "normal":
            return rng.normal(p.get("mean", 0), p.get("std", 1))
        if d == "uniform":
            return rng.uniform(p.get("low", 0), p.get("high", 1))
        if d == "exponential":
            return rng.exponential(p.get("scale", 1))
        if d == "gamma":
            if p.get("shape") is None or p.get("scale") is None:
                raise ValueError("For gamma distribution, both 'shape' and 'scale' params are required.")
            return rng.gamma(p["shape"], p["scale"])
        if d == "beta":
            if p.get("a") is None or p.get("b") is None:
                raise ValueError("For beta distribution, both 'a' and 'b' params are required.")
            return rng.beta(p["a"], p["b"])
        if d == "logistic":
            return rng.logistic(p.get("loc", 0), p.get("scale", 1))
        if d == "weibull":
            if p.get("a") is None:
                raise ValueError("For weibull distribution, 'a' param is required.")
            return rng.weibull(p["a"])
        return rng.lognormal(p.get("mean", 0), p.get("sigma", 1))


class BanditPort:
    """Scalar restatement of KArmedBandit.  `make_rng(seed)` builds the RandomState-like object (so tests can
    plug a real RandomState, a LegacyDistributions(MTPrimitives) or a Philox-driven source)."""

    def __init__(self, k=10, arm_params=None, seed=None, make_rng=None):
        self.k, self.seed, self.arm_params = k, seed, arm_params
        self.make_rng = make_rng or (lambda s: LegacyDistributions(MTPrimitives(np.random.RandomState(s))))
        self.np_random = self.make_rng(seed)
        self.timestep = 0
        self._init_arms()

    def _init_arms(self):  # k_armed_bandit.py:155-173
        means = self.np_random.randn(self.k)
        self.arms = [ArmPort("normal", mean=m, std=1) for m in means]
        for idx, params in (self.arm_params or {}).items():
            if idx < 0 or idx > self.k - 1:
                raise ValueError(f"Invalid arm index: {idx}, the arm index range is between 0 and {self.k-1}.")
            self.arms[idx] = ArmPort(**params)

    def step(self, action):  # k_armed_bandit.py:175-200
        assert 0 <= action < self.k, "Invalid action, must be between 0 and k-1."
        reward = self.arms[action].sample(self.np_random)
        self.timestep += 1
        for arm in self.arms:
            arm.update_param(self.timestep)
        return action, reward, True, False, {}

    def reset(self, seed=None):  # k_armed_bandit.py:202-217 -- timestep deliberately survives
        if seed is not None:
            self.seed = seed
        self.np_random = self.make_rng(self.seed)
        self._init_arms()
        return 0, {}

    @property
    def state(self):
        return self.timestep, [a.current_params for a in self.arms]


# -- engine-level oracle: what the device must produce for (seed, env g, timestep t, arm spec) -----------------
def default_arm_mean(seed, g, arm):
    """Default-arm mean of env g (float64): Box-Muller of words 0,1 of STREAM_BANDIT_MEANS block at counter=arm."""
    w = philox.draw_block(seed, g, arm, philox.STREAM_BANDIT_MEANS)
    u1, u2 = philox.u23(w[0]), philox.u23(w[1])
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def engine_sample(seed, g, t, family, p0, p1):
    """float64 sample the reference's sampler returns when its primitives are the engine's Philox stream of
    (env g, timestep t).  family = name; (p0, p1) = (mean,std) | (low,high) | (scale,-) | (shape,scale) | (a,b)
    | (loc,scale) | (a,-) | (mean,sigma).  Returns (sample, words_consumed)."""
    prim = PhiloxPrimitives(seed, int(g), int(t))
    rng = LegacyDistributions(prim)
    if family == "normal":
        v = rng.normal(p0, p1)
    elif family == "uniform":
        v = rng.uniform(p0, p1)
    elif family == "exponential":
        v = rng.exponential(p0)
    elif family == "gamma":
        v = rng.gamma(p0, p1)
    elif family == "beta":
        v = rng.beta(p0, p1)
    elif family == "logistic":
        v = rng.logistic(p0, p1)
    elif family == "weibull":
        v = rng.weibull(p0)
    elif family == "lognormal":
        v = rng.lognormal(p0, p1)
    else:
        raise ValueError(family)
    return v, prim.consumed


# -- vectorised engine-level oracle (same algorithms, N envs at once; checked against `engine_sample`) ------------
class _VecPrimitives:
    """Per-env word cursor over the STREAM_BANDIT substream of (env, t); float64 evaluation."""

    def __init__(self, seed, g, t):
        self.seed, self.g, self.t = seed, np.asarray(g, np.uint64), np.asarray(t, np.uint64)
        self.n = len(self.g)
        self.cur = np.zeros(self.n, np.int64)
        self._blocks = []

    def _words(self, mask):
        idx = np.nonzero(mask)[0]
        j, k = self.cur[idx] >> 2, self.cur[idx] & 3
        need = int(j.max()) + 1 if len(idx) else 0
        while len(self._blocks) < need:
            blk = philox.draw_block(self.seed, self.g, self.t, philox.STREAM_BANDIT, block=len(self._blocks))
            self._blocks.append(np.stack(blk, axis=1))
        out = np.zeros(self.n, np.uint32)
        if len(idx):
            stack = np.stack(self._blocks, axis=0)  # [blocks, n, 4]
            out[idx] = stack[j, idx, k]
        self.cur[idx] += 1
        return out

    def u(self, mask):
        return philox.u23(self._words(mask))

    def n_(self, mask):
        u1 = self.u(mask)
        u2 = self.u(mask)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)

    def std_exp(self, mask):
        return -np.log(1.0 - self.u(mask))


def _vec_std_gamma(pr, shape, active):
    out = np.zeros(pr.n)
    pending = active.copy()
    with np.errstate(all="ignore"):
        if shape == 1.0:
            return np.where(active, pr.std_exp(active), 0.0)
        if shape == 0.0:
            return out
        if shape < 1.0:
            while pending.any():
                u = pr.u(pending)
                v = pr.std_exp(pending)
                low = u <= 1.0 - shape
                x1 = np.power(u, 1.0 / shape)
                y = -np.log((1.0 - u) / shape)
                x2 = np.power(1.0 - shape + shape * y, 1.0 / shape)
                x = np.where(low, x1, x2)
                acc = pending & np.where(low, x1 <= v, x2 <= v + y)
                out[acc] = x[acc]
                pending &= ~acc
            return out
        b = shape - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * b)
        while pending.any():
            x = np.zeros(pr.n)
            v = np.zeros(pr.n)
            inner = pending.copy()
            while inner.any():
                xn = pr.n_(inner)
                vn = 1.0 + c * xn
                x[inner], v[inner] = xn[inner], vn[inner]
                inner &= ~(vn > 0.0)
            v = v * v * v
            u = pr.u(pending)
            acc = (u < 1.0 - 0.0331 * (x * x) * (x * x)) | (np.log(u) < 0.5 * x * x + b * (1.0 - v + np.log(v)))
            acc &= pending
            out[acc] = (b * v)[acc]
            pending &= ~acc
    return out


def engine_sample_vec(seed, g, t, family, p0, p1=0.0):
    """Vectorised `engine_sample`: float64 samples for envs g at step counter t (scalar or per-env).  Returns
    (samples float64 [n], words consumed int64 [n])."""
    g = np.asarray(g)
    t = np.broadcast_to(np.asarray(t, np.uint64), g.shape)
    pr = _VecPrimitives(seed, g, t)
    all_ = np.ones(pr.n, bool)
    with np.errstate(all="ignore"):
        if family == "normal":
            v = p0 + p1 * pr.n_(all_)
        elif family == "uniform":
            v = p0 + (p1 - p0) * pr.u(all_)
        elif family == "exponential":
            v = p0 * pr.std_exp(all_)
        elif family == "gamma":
            v = p1 * _vec_std_gamma(pr, float(p0), all_)
        elif family == "beta":
            a, b = float(p0), float(p1)
            if a <= 1.0 and b <= 1.0:
                v = np.zeros(pr.n)
                pending = all_.copy()
                while pending.any():
                    u = pr.u(pending)
                    w = pr.u(pending)
                    x = np.power(u, 1.0 / a)
                    y = np.power(w, 1.0 / b)
                    acc = pending & (x + y <= 1.0)
                    pos = acc & (x + y > 0)
                    v[pos] = (x / (x + y))[pos]
                    tiny = acc & ~pos
                    if tiny.any():
                        lx, ly = np.log(u) / a, np.log(w) / b
                        lm = np.maximum(lx, ly)
                        lx, ly = lx - lm, ly - lm
                        v[tiny] = np.exp(lx - np.log(np.exp(lx) + np.exp(ly)))[tiny]
                    pending &= ~acc
            else:
                ga = _vec_std_gamma(pr, a, all_)
                gb = _vec_std_gamma(pr, b, all_)
                v = ga / (ga + gb)
        elif family == "logistic":
            u = pr.u(all_)
            v = p0 + p1 * np.log(u / (1.0 - u))
        elif family == "weibull":
            v = np.zeros(pr.n) if p0 == 0.0 else np.power(pr.std_exp(all_), 1.0 / p0)
        elif family == "lognormal":
            v = np.exp(p0 + p1 * pr.n_(all_))
        else:
            raise ValueError(family)
    return v, pr.cur.copy()
