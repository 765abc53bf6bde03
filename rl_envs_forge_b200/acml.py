"""VecGamblersProblem / VecJacksCarRental -- N instances of the reference's ACML envs stepped by sm_100a kernels.

Drop-in surface for
  `GamblersProblem`  rl_envs_forge/envs/acml/gambler_problem/gambler_problem.py  (kwargs :11-16, step :36-66,
                     reset :68-76, spaces :30-31, build_mdp :89-119)
  `JacksCarRental`   rl_envs_forge/envs/acml/car_rental/car_rental.py  (kwargs :10-20, step :106-121, reset :123-143,
                     spaces :46-49, poisson tables :165-227, build_mdp :229-301)
State lives in device memory (capital int32 [N]; cars int32 [N,2]).  Randomness: the reference draws from numpy's global
legacy RandomState (`np.random.rand`, `scipy.stats.poisson.rvs`); here every draw comes from the engine's Philox streams
keyed by (seed, global env index, step), and parity is judged under identical injected draws (GamblersProblem: the
`draws` argument; JacksCarRental: the oracle consumes the same Philox doubles, tests/test_gpu_acml.py).
"""
import math

import numpy as np

from . import _lib, spaces
from .vec_env import VecEnv, resolve_seed


def _host_actions(env, actions, n):
    """Host action array -> pinned int32 tensor [n] (caller-pinned int32 tensors are used in place)."""
    torch = env._torch
    if torch.is_tensor(actions):
        a = actions.reshape(-1).to(torch.int32)
    else:
        a = np.asarray(actions)
        a = (a.reshape(1) if a.shape == () else a.reshape(-1)).astype(np.int32, copy=False)
    return env._host_in("action", a, torch.int32, (n,))


class _AcmlBase(VecEnv):
    """Shared step plumbing: outputs, host transports (zerocopy / staged), final observations."""
    _OBS_SHAPE = ()

    def _alloc_outputs(self):
        torch, n, dev = self._torch, self.num_envs, self.device
        self._obs = torch.zeros((n,) + self._OBS_SHAPE, dtype=torch.int32, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._term = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._final_obs = None

    def _step_common(self, actions, launch, return_final_obs):
        """launch(p_action, obs, reward, term, trunc, final_obs) enqueues the kernel; buffers are device tensors or, for the
        zero-copy host transport, pinned host tensors the kernel writes directly."""
        torch, n = self._torch, self.num_envs
        host_io = not (torch.is_tensor(actions) and actions.is_cuda)
        mode = None
        if host_io:
            mode = "zerocopy" if self.host_transport == "zerocopy" else "staged"
            h_act = _host_actions(self, actions, n)
            p_act = h_act.data_ptr() if mode == "zerocopy" else _lib.ptr(self._to_device("action", h_act, torch.int32, (n,)))
        else:
            if actions.numel() != n:
                raise ValueError(f"actions must hold {n} values, got {tuple(actions.shape)}")
            act = actions.reshape(-1)
            act = (act if act.dtype == torch.int32 else act.to(torch.int32)).contiguous()
            p_act = _lib.ptr(act)
        fobs = None
        if mode == "zerocopy":
            obs = self._pinned("o_obs", (n,) + self._OBS_SHAPE, torch.int32)
            rew = self._pinned("o_reward", (n,), torch.float32)
            term = self._pinned("o_term", (n,), torch.uint8)
            trunc = self._pinned("o_trunc", (n,), torch.uint8)
            if return_final_obs:
                fobs = self._pinned("o_final_obs", (n,) + self._OBS_SHAPE, torch.int32)
        else:
            obs, rew, term, trunc = self._obs, self._reward, self._term, self._trunc
            if return_final_obs:
                if self._final_obs is None:
                    self._final_obs = torch.zeros((n,) + self._OBS_SHAPE, dtype=torch.int32, device=self.device)
                fobs = self._final_obs
        launch(p_act, obs, rew, term, trunc, fobs)
        self._t += 1
        if self.strict:
            self._strict_check()
        info = {}
        if host_io:
            if mode == "zerocopy":
                torch.cuda.current_stream(self.device).synchronize()
                out = [x.numpy() for x in (obs, rew, term, trunc)] + ([fobs.numpy()] if return_final_obs else [])
            else:
                named = [("obs", obs), ("reward", rew), ("term", term), ("trunc", trunc)]
                out = self._to_host(named + ([("final_obs", fobs)] if return_final_obs else []))
            if return_final_obs:
                info["final_observation"] = out[4]
            return out[0], out[1], out[2].view(np.bool_), out[3].view(np.bool_), info
        if return_final_obs:
            info["final_observation"] = fobs
        return obs, rew, term.view(torch.bool), trunc.view(torch.bool), info

    def _strict_check(self):
        pass

    def _rollout_buffers(self, K, actions, return_rewards):
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
        return actions, rew, flg

    def get_state(self):
        return {"state": self._state.clone(), "ep_len": self.ep_len.clone(), "ep_ret": self.ep_ret.clone(), "t": self._t}

    def set_state(self, s):
        self._state.copy_(s["state"])
        self.ep_len.copy_(s["ep_len"])
        self.ep_ret.copy_(s["ep_ret"])
        self._t = int(s.get("t", self._t))


class VecGamblersProblem(_AcmlBase):
    """N GamblersProblem instances.  kwargs as in the reference (+ num_envs, device, seed, env_offset, autoreset, strict,
    and `episode_length_limit`, an extension: the reference never truncates)."""

    def __init__(self, num_envs=1, goal_amount=100, win_probability=0.4, start_capital=1, *, episode_length_limit=None,
                 device=None, seed=None, env_offset=0, autoreset=True, strict=False):
        if not isinstance(goal_amount, (int, np.integer)) or goal_amount < 0:
            raise ValueError("goal_amount must be a non-negative integer")
        if not 0.0 <= float(win_probability) <= 1.0:
            raise ValueError("win_probability must lie in [0, 1]")
        super().__init__(num_envs, device, seed, env_offset, autoreset)
        torch = self._torch
        self.goal_amount, self.win_probability, self.start_capital = int(goal_amount), win_probability, int(start_capital)
        self.episode_length_limit = episode_length_limit
        self.strict = bool(strict)
        self.single_action_space = spaces.Discrete(self.goal_amount + 1)       # gambler_problem.py:30
        self.single_observation_space = spaces.Discrete(self.goal_amount + 1)  # :31
        self.action_space = spaces.Batched(self.single_action_space, num_envs)
        self.observation_space = spaces.Batched(self.single_observation_space, num_envs)
        p = _lib.GamblerParams()
        p.goal_amount, p.start_capital = self.goal_amount, self.start_capital
        p.episode_limit = int(episode_length_limit) if episode_length_limit is not None else 0
        # np.random.rand() < p  <=>  r < ceil(p * 2^32) for u = r * 2^-32 (p * 2^32 is exact in binary64)
        p.win_threshold = min(int(math.ceil(float(win_probability) * 4294967296.0)), 1 << 32)
        self._params = p
        self._state = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        self._alloc_outputs()
        self.reset()

    @property
    def state(self):
        """int32 [N] capital (the reference's `state`); assignable with a scalar, array or tensor."""
        return self._state

    @state.setter
    def state(self, value):
        v = self._torch.as_tensor(value, device=self.device).to(self._torch.int32).reshape(-1)
        self._state.copy_(v.expand(self.num_envs) if v.numel() == 1 else v)

    def reset(self, *, seed=None, mask=None):
        """GamblersProblem.reset (:68-76): capital <- start_capital for all envs (or those selected by `mask`)."""
        if seed is not None:
            self.seed, self._seed = seed, resolve_seed(seed)
        m = None if mask is None else self._torch.as_tensor(mask, device=self.device).to(self._torch.uint8).contiguous()
        self._call(self._lib.ef_gambler_reset, self.start_capital, _lib.ptr(self._state), _lib.ptr(self.ep_len),
                   _lib.ptr(self.ep_ret), _lib.ptr(self._obs), _lib.ptr(m), self.num_envs)
        self._reset_epoch += 1
        return self._obs, {}

    def step(self, actions, draws=None, *, return_final_obs=False):
        """GamblersProblem.step (:36-66) for all envs: actions int [N] (the bets).  `draws`: optional uint32-valued [N]
        injected draws r (the env wins iff r * 2**-32 < win_probability); default: Philox.
        Returns (capital int32 [N], reward fp32 [N], terminated bool [N], truncated bool [N], info)."""
        torch, n = self._torch, self.num_envs
        drw = None
        if draws is not None:
            if torch.is_tensor(draws) and draws.is_cuda:
                drw = (draws if draws.dtype == torch.int32 else draws.to(torch.int64).to(torch.int32)).contiguous()
            else:
                drw = self._to_device("draw", np.asarray(draws).astype(np.uint32, copy=False).view(np.int32), torch.int32, (n,))

        def launch(p_act, obs, rew, term, trunc, fobs):
            self._call(self._lib.ef_gambler_step, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len),
                       _lib.ptr(self.ep_ret), p_act, _lib.ptr(drw), self._seed, self._t_arg(), self.env_offset,
                       _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(term), _lib.ptr(trunc), _lib.ptr(fobs),
                       _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), n, int(self.autoreset))
        return self._step_common(actions, launch, return_final_obs)

    def rollout(self, K, actions=None, *, return_rewards=False):
        """K fused steps in one kernel (capital in registers, auto-reset on).  actions: CUDA int32 [K,N] or None for the
        in-kernel policy (a legal bet, uniform in 0..min(s, goal - s)).  Returns (rewards fp32 [K,N], flags uint8 [K,N]:
        bit0 terminated, bit1 truncated) if return_rewards else None; episode statistics accumulate in `stats`."""
        actions, rew, flg = self._rollout_buffers(K, actions, return_rewards)
        self._call(self._lib.ef_gambler_rollout, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len),
                   _lib.ptr(self.ep_ret), _lib.ptr(actions), self._seed, self._t_arg(), self.env_offset, int(K), _lib.ptr(rew),
                   _lib.ptr(flg), _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), self.num_envs)
        self._t += int(K)
        return (rew, flg) if return_rewards else None

    def _strict_check(self):
        try:
            self.check_errors()
        except ValueError as e:
            raise AssertionError("Invalid action") from e

    def _describe_error(self, count, env, detail, kind_id, kind):
        return f"Invalid action {detail} (env {env}; {count} rejected env-step(s))"

    def build_mdp(self):
        """build_mdp (:89-119): {(state, bet): {"win": (next, reward, done, p), "loss": (...)}}."""
        d, goal, p = {}, self.goal_amount, self.win_probability
        for s in range(1, goal):
            for a in range(1, min(s, goal - s) + 1):
                win_state, lose_state = min(s + a, goal), max(s - a, 0)
                d[(s, a)] = {"win": (win_state, 1 if win_state == goal else 0, win_state == goal, p),
                             "loss": (lose_state, 0, lose_state == 0, 1 - p)}
        return d


def _poisson_pmf(k, lam):
    """scipy.stats.poisson.pmf(k, lam) = exp(k*log(lam) - lam - lgamma(k+1)) (scipy is not a dependency of the package)."""
    if lam == 0:
        return 1.0 if k == 0 else 0.0
    return math.exp(k * math.log(lam) - lam - math.lgamma(k + 1))


class VecJacksCarRental(_AcmlBase):
    """N JacksCarRental instances.  kwargs as in the reference (+ num_envs, device, seed, env_offset, autoreset and the
    extension `episode_length_limit`: the reference never ends an episode)."""
    _OBS_SHAPE = (2,)

    def __init__(self, num_envs=1, max_cars=20, max_move_cars=5, rental_credit=10, move_car_cost=2, request_lambda=(3, 4),
                 return_lambda=(3, 2), init_state_option="random", poisson_max_value=12, *, episode_length_limit=None,
                 device=None, seed=None, env_offset=0, autoreset=True):
        lams = [float(x) for x in list(request_lambda) + list(return_lambda)]
        if len(lams) != 4 or any(not (x >= 0) for x in lams):
            raise ValueError("request_lambda and return_lambda must each hold two non-negative rates")
        if any(x > 1e6 for x in lams):
            raise NotImplementedError("Poisson rates above 1e6 are not supported by the device path")
        for v in (max_cars, max_move_cars, rental_credit, move_car_cost):
            if not isinstance(v, (int, np.integer)):
                raise TypeError("max_cars, max_move_cars, rental_credit and move_car_cost must be integers")
        super().__init__(num_envs, device, seed, env_offset, autoreset)
        torch = self._torch
        self.max_cars, self.max_move_cars = int(max_cars), int(max_move_cars)
        self.rental_credit, self.move_car_cost = int(rental_credit), int(move_car_cost)
        self.request_lambda, self.return_lambda = list(request_lambda), list(return_lambda)
        self.episode_length_limit = episode_length_limit
        self.strict = False
        self.single_action_space = spaces.Discrete(2 * self.max_move_cars + 1)                       # car_rental.py:46
        self.single_observation_space = spaces.Tuple((spaces.Discrete(self.max_cars + 1),
                                                      spaces.Discrete(self.max_cars + 1)))            # :47-49
        self.action_space = spaces.Batched(self.single_action_space, num_envs)
        self.observation_space = spaces.Batched(self.single_observation_space, num_envs)
        p = _lib.CarRentalParams()
        p.max_cars, p.max_move_cars = self.max_cars, self.max_move_cars
        p.rental_credit, p.move_car_cost = self.rental_credit, self.move_car_cost
        p.episode_limit = int(episode_length_limit) if episode_length_limit is not None else 0
        for j, lam in enumerate(lams):
            p.lam[j] = lam
            p.exp_neg_lam[j] = math.exp(-lam)       # what numpy's random_poisson_mult computes (libm exp)
            if lam >= 10:                           # numpy's random_poisson_ptrs constants (float64, libm sqrt / log)
                slam = math.sqrt(lam)
                b = 0.931 + 2.53 * slam
                a = -0.059 + 0.02483 * b
                p.log_lam[j], p.ptrs_a[j], p.ptrs_b[j] = math.log(lam), a, b
                p.ptrs_log_invalpha[j] = math.log(1.1239 + 1.1328 / (b - 3.4))
                p.ptrs_vr[j] = 0.9277 - 3.6224 / (b - 2)
        self._params = p
        self._precompute_poisson_probabilities(poisson_max_value)
        self._state = torch.zeros((self.num_envs, 2), dtype=torch.int32, device=self.device)
        self._alloc_outputs()
        self.reset(init_state_option)

    @property
    def state(self):
        """int32 [N,2] (cars at A, cars at B); assignable with a pair, an [N,2] array or tensor."""
        return self._state

    @state.setter
    def state(self, value):
        v = self._torch.as_tensor(np.asarray(value) if not self._torch.is_tensor(value) else value, device=self.device)
        v = v.to(self._torch.int32).reshape(-1, 2)
        self._state.copy_(v.expand(self.num_envs, 2) if v.shape[0] == 1 else v)

    def reset(self, option="random", *, seed=None, mask=None):
        """JacksCarRental.reset (:123-143): "random" -> each location uniform in 0..max_cars (Philox reset stream);
        "equal" -> max_cars // 2 each.  Any other option leaves the state untouched, like the reference."""
        if seed is not None:
            self.seed, self._seed = seed, resolve_seed(seed)
        if option in ("random", "equal"):
            self._params.init_equal = 1 if option == "equal" else 0
            m = None if mask is None else self._torch.as_tensor(mask, device=self.device).to(self._torch.uint8).contiguous()
            self._call(self._lib.ef_car_rental_reset, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len),
                       _lib.ptr(self.ep_ret), self._seed, self._reset_epoch, self.env_offset, _lib.ptr(self._obs),
                       _lib.ptr(m), self.num_envs)
            self._reset_epoch += 1
        return self._state, {"Reset Option": option}

    def step(self, actions, *, return_final_obs=False):
        """JacksCarRental.step (:106-121) for all envs: actions int [N] in 0..2*max_move_cars.
        Returns (cars int32 [N,2], reward fp32 [N], terminated (all False), truncated, info)."""
        n = self.num_envs

        def launch(p_act, obs, rew, term, trunc, fobs):
            self._call(self._lib.ef_car_rental_step, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len),
                       _lib.ptr(self.ep_ret), p_act, self._seed, self._t_arg(), self.env_offset, _lib.ptr(obs),
                       _lib.ptr(rew), _lib.ptr(term), _lib.ptr(trunc), _lib.ptr(fobs), _lib.ptr(self._stats_raw),
                       _lib.ptr(self._clock), n, int(self.autoreset))
        return self._step_common(actions, launch, return_final_obs)

    def rollout(self, K, actions=None, *, return_rewards=False):
        """K fused steps in one kernel (cars in registers).  actions: CUDA int32 [K,N] or None (uniform random action).
        Returns (rewards fp32 [K,N], flags uint8 [K,N]: bit1 truncated) if return_rewards else None."""
        actions, rew, flg = self._rollout_buffers(K, actions, return_rewards)
        self._call(self._lib.ef_car_rental_rollout, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len),
                   _lib.ptr(self.ep_ret), _lib.ptr(actions), self._seed, self._t_arg(), self.env_offset, int(K), _lib.ptr(rew),
                   _lib.ptr(flg), _lib.ptr(self._stats_raw), _lib.ptr(self._clock), self.num_envs)
        self._t += int(K)
        return (rew, flg) if return_rewards else None

    # -- host-side tables of the reference (dynamic-programming users read them) ------------------------------------------
    def _precompute_poisson_probabilities(self, max_value=12):
        """car_rental.py:165-198."""
        r, t = self.request_lambda, self.return_lambda
        self.poisson_probs = {
            "request": {"A": [_poisson_pmf(i, r[0]) for i in range(max_value + 1)],
                        "B": [_poisson_pmf(i, r[1]) for i in range(max_value + 1)]},
            "return": {"A": [_poisson_pmf(i, t[0]) for i in range(max_value + 1)],
                       "B": [_poisson_pmf(i, t[1]) for i in range(max_value + 1)]},
        }

    def _get_probability_of_rental_requests(self, request_A, request_B):
        return self.poisson_probs["request"]["A"][request_A] * self.poisson_probs["request"]["B"][request_B]

    def _get_probability_of_car_returns(self, return_A, return_B):
        return self.poisson_probs["return"]["A"][return_A] * self.poisson_probs["return"]["B"][return_B]

    def _move_cars_and_calculate_cost(self, state, action):
        """car_rental.py:54-74 (host helper; `action` in -max_move_cars .. max_move_cars)."""
        a, b = state
        moved = min(a if action > 0 else b, abs(action), self.max_move_cars)
        sign = (action > 0) - (action < 0)
        return (max(a - moved * sign, 0), max(b + moved * sign, 0)), abs(action) * self.move_car_cost

    def build_mdp(self):
        """build_mdp (:229-301): {(state, action): (next state for zero requests and returns, expected reward, False)}.
        The reference's seven nested Python loops are evaluated as numpy sums (same terms, different association)."""
        n = self.max_cars + 1
        pr, pt = self.poisson_probs["request"], self.poisson_probs["return"]
        if len(pr["A"]) < n:
            raise IndexError("list index out of range")   # the reference indexes the tables up to max_cars
        req = np.outer(pr["A"][:n], pr["B"][:n])
        ret_mass = float(np.sum(np.outer(pt["A"][:n], pt["B"][:n])))
        out = {}
        k = np.arange(n)
        for a0 in range(n):
            for b0 in range(n):
                for action in range(-self.max_move_cars, self.max_move_cars + 1):
                    (a, b), cost = self._move_cars_and_calculate_cost((a0, b0), action)
                    reward = (np.minimum(a, k)[:, None] + np.minimum(b, k)[None, :]) * self.rental_credit - cost
                    out[((a0, b0), action)] = ((min(a, self.max_cars), min(b, self.max_cars)),
                                               float(np.sum(req * reward)) * ret_mass, False)
        return out
