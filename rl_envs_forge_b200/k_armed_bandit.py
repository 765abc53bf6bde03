"""VecKArmedBandit -- N k-armed bandits stepped by one sm_100a kernel launch.

Drop-in surface for the reference classes `Arm` and `KArmedBandit`
(rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py): `KArmedBandit(k, arm_params, seed)` (:126-153), default normal
arms with randn means (:160-165), custom arms with index check (:167-173), `step(action)` returning
`(action, reward, True, False, {})` (:175-200), `reset(seed)` that rebuilds the arms but keeps `timestep` (:202-217),
`state` (:273-282), and `Arm` with absolute parameter shifts `init + f(t)` (:53-64).

The arm set is shared by all N bandits (default arms have a per-bandit mean re-derived in-kernel from Philox).  User
`param_functions` are arbitrary Python callables: they are evaluated on the host, per timestep, into the small arm
table the kernel reads; sampling itself (8 families) runs on the device in fp32.

`rng="mt19937"` is the stream-exact mode: every bandit owns numpy's legacy `RandomState(seed_i)` (what the reference
builds at :145 / :215) with `seed_i = seed + global env index` (or `seed` given as a sequence of N ints), default means are
that generator's `randn(k)`, and `Arm.sample` runs numpy's legacy float64 algorithms on the device
(`ef_bandit_step_mt19937`): `VecKArmedBandit(1, seed=s, rng="mt19937")` replays `KArmedBandit(seed=s)`.  2.5 KB of
generator state per bandit, so it is capped at 2**18 bandits; the Philox mode is the throughput path.
"""
import numpy as np

from . import _lib, spaces
from .vec_env import VecEnv, resolve_seed

IMPLEMENTED_DISTRIBUTIONS = ["normal", "uniform", "exponential", "gamma", "beta", "logistic", "weibull", "lognormal"]

# family -> ((param name, default or REQUIRED), (second param ...)) in the order packed into ef_arm.p0 / p1 (Arm.sample :66-122)
_REQUIRED = object()
_PARAMS = {
    "normal": (("mean", 0), ("std", 1)),
    "uniform": (("low", 0), ("high", 1)),
    "exponential": (("scale", 1),),
    "gamma": (("shape", _REQUIRED), ("scale", _REQUIRED)),
    "beta": (("a", _REQUIRED), ("b", _REQUIRED)),
    "logistic": (("loc", 0), ("scale", 1)),
    "weibull": (("a", _REQUIRED),),
    "lognormal": (("mean", 0), ("sigma", 1)),
}
_MISSING_MSG = {
    "gamma": "For gamma distribution, both 'shape' and 'scale' params are required.",
    "beta": "For beta distribution, both 'a' and 'b' params are required.",
    "weibull": "For weibull distribution, 'a' param is required.",
}


class Arm:
    """Host descriptor of one arm (reference `Arm`, :11-64).  Sampling happens on the device."""

    IMPLEMENTED_DISTRIBUTIONS = IMPLEMENTED_DISTRIBUTIONS

    def __init__(self, distribution="normal", param_functions=None, **kwargs):
        self.distribution = distribution
        if self.distribution not in self.IMPLEMENTED_DISTRIBUTIONS:
            raise ValueError(
                f"Invalid distribution: {self.distribution}. Please use one of {self.IMPLEMENTED_DISTRIBUTIONS}.")
        self.init_params = kwargs.copy()
        self.current_params = kwargs.copy()
        self.param_functions = param_functions if param_functions else []
        self.per_env_mean = False  # True for default arms: `mean` differs per bandit and lives on the device

    def update_param(self, timestep):
        for fd in self.param_functions:
            self.current_params[fd["target_param"]] = self.init_params[fd["target_param"]] + fd["function"](timestep)

    def packed(self):
        """(family id, p0, p1) as the kernel wants them; raises what Arm.sample / numpy would raise at sample time."""
        d, p = self.distribution, self.current_params
        vals = []
        for name, default in _PARAMS[d]:
            v = p.get(name) if default is _REQUIRED else p.get(name, default)
            if v is None and self.per_env_mean and name == "mean":
                v = 0.0
            if v is None:
                raise ValueError(_MISSING_MSG.get(d, f"parameter {name!r} of a {d} arm is None"))
            vals.append(float(v))
        vals += [0.0] * (2 - len(vals))
        p0, p1 = vals
        if not (np.isfinite(p0) and np.isfinite(p1)):
            # numpy's samplers return NaN for NaN parameters; on the device a NaN shape would never be accepted by a
            # rejection loop -- refuse it here, loudly
            raise ValueError(f"parameters of a {d} arm must be finite, got {p0!r}, {p1!r}")
        # numpy legacy argument checks (RandomState.normal/exponential/gamma/beta/logistic/weibull/lognormal)
        if d in ("normal", "logistic") and p1 < 0:
            raise ValueError("scale < 0")
        if d == "lognormal" and p1 < 0:
            raise ValueError("sigma < 0")
        if d == "exponential" and p0 < 0:
            raise ValueError("scale < 0")
        if d == "gamma":
            if p0 < 0:
                raise ValueError("shape < 0")
            if p1 < 0:
                raise ValueError("scale < 0")
        if d == "beta":
            if p0 <= 0:
                raise ValueError("a <= 0")
            if p1 <= 0:
                raise ValueError("b <= 0")
        if d == "weibull" and p0 < 0:
            raise ValueError("a < 0")
        if self.per_env_mean:
            return _lib.ARM_FAMILIES["default_normal"], 0.0, p1
        return _lib.ARM_FAMILIES[d], p0, p1


class VecKArmedBandit(VecEnv):
    MT19937_MAX_ENVS = 1 << 18

    def __init__(self, num_envs=1, k=10, arm_params=None, seed=None, *, device=None, env_offset=0, strict=False,
                 rng="philox", mean_table=True):
        """mean_table: keep the default arms' per-env means (`np_random.randn(k)` of one reference instance, :160-165) in an
        fp32 [N,k] device table (4 k bytes per env) that step()/rollout() look up; False re-derives a mean from its Philox
        stream whenever a default arm is pulled -- same bits, no table, one more Philox call on those steps."""
        if rng not in ("philox", "mt19937"):
            raise ValueError("rng must be 'philox' or 'mt19937'")
        self.rng = rng
        self._seed_list = None
        if rng == "mt19937":
            if num_envs > self.MT19937_MAX_ENVS:
                raise ValueError(f"rng='mt19937' keeps 2.5 KB of generator state per bandit: at most {self.MT19937_MAX_ENVS} envs")
            if seed is not None and not isinstance(seed, (int, np.integer)):
                self._seed_list = [int(x) for x in seed]
                if len(self._seed_list) != num_envs:
                    raise ValueError(f"seed sequence must hold {num_envs} integers")
                seed = self._seed_list[0]
        super().__init__(num_envs, device, seed, env_offset, autoreset=True)
        torch = self._torch
        self.k = k
        self.arm_params = arm_params
        self.timestep = 0
        self.strict = bool(strict)
        self._steps_since_reset = 0
        self.single_action_space = spaces.Discrete(k)
        self.single_observation_space = spaces.Discrete(1)
        self.action_space = spaces.Batched(self.single_action_space, num_envs)
        self.observation_space = spaces.Batched(self.single_observation_space, num_envs)
        n, dev = self.num_envs, self.device
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._obs = torch.zeros(n, dtype=torch.int32, device=dev)
        self._done = torch.ones(n, dtype=torch.bool, device=dev)
        self._trunc = torch.zeros(n, dtype=torch.bool, device=dev)
        self._zero_obs = torch.zeros(n, dtype=torch.int32, device=dev)
        self._arm_dev = None
        self._arm_key = None
        self.arms = None
        self.mean_table = bool(mean_table)
        self._means_dev, self._means_key = None, None
        self._init_arms()
        if rng == "mt19937":
            self._reward64 = torch.zeros(n, dtype=torch.float64, device=dev)
            self._init_mt19937()

    # -- stream-exact mode: per-env numpy RandomState ---------------------------------------------------------------------
    def _init_mt19937(self):
        """Seed every env's generator with numpy itself and draw its default means (`randn(k)`, :162) -- what
        `KArmedBandit.__init__` / `reset` do (:145, :215) -- then move the generator state to the device."""
        from . import mt19937
        torch, dev = self._torch, self.device
        base = (self._seed & 0xFFFFFFFF) if self.seed is None else int(self.seed)
        seeds = mt19937.env_seeds(base, self._seed_list, self.env_offset, self.num_envs)
        key, pos, has_gauss, gauss, means = mt19937.initial_state(seeds, self.k)
        self._mt_key = torch.from_numpy(key.view(np.int32)).to(dev)
        self._mt_pos = torch.from_numpy(pos).to(dev)
        self._mt_has_gauss = torch.from_numpy(has_gauss).to(dev)
        self._mt_gauss = torch.from_numpy(gauss).to(dev)
        self._mt_means = torch.from_numpy(means).to(dev)

    def _arm_table64(self):
        """Device [k] ef_arm64 table of the parameters in force now (float64, unrounded)."""
        torch = self._torch
        tab = np.zeros(self.k, dtype=[("family", "<i4"), ("reserved", "<i4"), ("p0", "<f8"), ("p1", "<f8")])
        for i, a in enumerate(self.arms):
            tab[i]["family"], tab[i]["p0"], tab[i]["p1"] = a.packed()
        raw = torch.from_numpy(tab.view(np.uint8).copy())
        stage, slot = self._pinned_ring("h_arms64", tuple(raw.shape), torch.uint8)   # see _arm_table
        stage.copy_(raw)
        dev = self._device_scratch("arms64", tuple(raw.shape), torch.uint8).copy_(stage, non_blocking=True)
        self._mark_staged(slot)
        return dev

    @property
    def reward64(self):
        """float64 [N] unrounded samples of the last step (rng='mt19937' only)."""
        return self._reward64

    def _step_mt19937(self, actions):
        torch, n = self._torch, self.num_envs
        host_io = not (torch.is_tensor(actions) and actions.is_cuda)
        if host_io:
            a = (actions.numpy() if torch.is_tensor(actions) else np.asarray(actions)).reshape(-1)
            if self.strict:
                assert ((0 <= a) & (a < self.k)).all(), "Invalid action, must be between 0 and k-1."
            act = self._to_device("action", self._host_in("action", a.astype(np.int32, copy=False), torch.int32, (n,)),
                                  torch.int32, (n,))
        else:
            if tuple(actions.shape) != (n,):
                raise ValueError(f"actions must have shape ({n},), got {tuple(actions.shape)}")
            act = (actions if actions.dtype == torch.int32 else actions.to(torch.int32)).contiguous()
        arms = self._arm_table64()
        self._call(self._lib.ef_bandit_step_mt19937, _lib.ptr(arms), self.k, _lib.ptr(act), _lib.ptr(self._mt_key),
                   _lib.ptr(self._mt_pos), _lib.ptr(self._mt_has_gauss), _lib.ptr(self._mt_gauss), _lib.ptr(self._mt_means),
                   self.env_offset, _lib.ptr(self._obs), _lib.ptr(self._reward), _lib.ptr(self._reward64),
                   _lib.ptr(self._stats_raw), _lib.ptr(self.err), n)
        self._advance(1)
        if self.strict:
            try:
                self.check_errors()
            except ValueError as e:
                raise AssertionError("Invalid action, must be between 0 and k-1.") from e
        if host_io:
            obs, rew = self._to_host([("obs", self._obs), ("reward", self._reward)])
            return obs, rew, np.ones(n, bool), np.zeros(n, bool), {}
        return self._obs, self._reward, self._done, self._trunc, {}

    # -- arms (:155-173) ----------------------------------------------------------------------------------------------
    def _init_arms(self):
        self.arms = []
        for _ in range(self.k):
            arm = Arm(distribution="normal", mean=None, std=1)
            arm.per_env_mean = True
            self.arms.append(arm)
        if self.arm_params:
            for idx, params in self.arm_params.items():
                if idx < 0 or (idx > self.k - 1):
                    raise ValueError(f"Invalid arm index: {idx}, the arm index range is between 0 and {self.k-1}.")
                self.arms[idx] = Arm(**params)
        self._arm_key = None

    @property
    def default_means(self):
        """fp32 [N,k] default-arm means (what `np_random.randn(k)` is to one reference instance); float64 and drawn by
        numpy's own generator with rng='mt19937'."""
        if self.rng == "mt19937":
            return self._mt_means
        out = self._torch.empty((self.num_envs, self.k), dtype=self._torch.float32, device=self.device)
        self._call(self._lib.ef_bandit_default_means, self.k, self._seed, self.env_offset, _lib.ptr(out), self.num_envs)
        return out

    def _means_ptr(self):
        """Device pointer of the cached default-mean table for ef_bandit_step_means, or None (no default arm in play, or
        mean_table=False).  Filled by ef_bandit_default_means -- the very values the kernel would re-derive -- and rebuilt
        when reset() re-keys the generator."""
        if not self.mean_table or not any(a.per_env_mean for a in self.arms):
            return None
        key = (self._seed, self.env_offset, self.k)
        if self._means_key != key:
            if self._means_dev is None:
                self._means_dev = self._torch.empty((self.num_envs, self.k), dtype=self._torch.float32, device=self.device)
            self._call(self._lib.ef_bandit_default_means, self.k, self._seed, self.env_offset, _lib.ptr(self._means_dev),
                       self.num_envs)
            self._means_key = key
        return _lib.ptr(self._means_dev)

    @property
    def state(self):
        return self.timestep, [arm.current_params for arm in self.arms]

    def _arm_table(self, K):
        """Device [K,k] ef_arm table for timesteps timestep .. timestep+K-1 (parameters in force when sampling)."""
        torch = self._torch
        dynamic = any(a.param_functions for a in self.arms)
        if not dynamic:
            key = (K, tuple((a.distribution, tuple(sorted(a.current_params.items(), key=lambda kv: kv[0])), a.per_env_mean)
                            for a in self.arms))
            if key == self._arm_key and self._arm_dev is not None:
                return self._arm_dev
        tab = np.zeros((K, self.k, 4), np.float32)
        itab = tab.view(np.int32)
        saved = [dict(a.current_params) for a in self.arms]
        for j in range(K):
            if j > 0:  # parameters in force for the (j+1)-th step: shifted by f(timestep + j)
                for a in self.arms:
                    a.update_param(self.timestep + j)
            for i, a in enumerate(self.arms):
                fam, p0, p1 = a.packed()
                itab[j, i, 0] = fam
                tab[j, i, 1], tab[j, i, 2] = p0, p1
        for a, s in zip(self.arms, saved):
            a.current_params = s
        # the table changes every step when an arm has param_functions, and nothing makes the host wait for the upload:
        # rotate pinned buffers behind events so that step i+1 cannot overwrite what step i's DMA is still reading
        stage, slot = self._pinned_ring("h_arms", (K, self.k, 4), torch.float32)
        stage.copy_(torch.from_numpy(tab))
        if self._arm_dev is None or tuple(self._arm_dev.shape) != (K, self.k, 4):
            self._arm_dev = torch.empty((K, self.k, 4), dtype=torch.float32, device=self.device)
        self._arm_dev.copy_(stage, non_blocking=True)
        self._mark_staged(slot)
        self._arm_key = None if dynamic else key
        return self._arm_dev

    # -- hot path ---------------------------------------------------------------------------------------------------
    def step(self, actions):
        """KArmedBandit.step (:175-200) for all bandits: (obs = actions, reward fp32 [N], done = True, truncated = False, {})."""
        if self.rng == "mt19937":
            return self._step_mt19937(actions)
        torch = self._torch
        n = self.num_envs
        host_io = not (torch.is_tensor(actions) and actions.is_cuda)
        mode = self.host_transport if host_io else None
        obs, rew = self._obs, self._reward
        if host_io:
            a = (actions.numpy() if torch.is_tensor(actions) else np.asarray(actions)).reshape(-1)
            if self.strict:
                assert ((0 <= a) & (a < self.k)).all(), "Invalid action, must be between 0 and k-1."
            src = actions.reshape(-1) if torch.is_tensor(actions) and actions.dtype == torch.int32 else a.astype(np.int32, copy=False)
            h_act = self._host_in("action", src, torch.int32, (n,))
            if mode == "staged":
                p_act = _lib.ptr(self._to_device("action", h_act, torch.int32, (n,)))
            else:   # the kernel reads the actions from pinned host memory directly
                p_act = h_act.data_ptr()
                h_obs = self._pinned("o_obs", (n,), torch.int32)
                h_rew = self._pinned("o_reward", (n,), torch.float32)
                if mode == "zerocopy":   # ... and writes obs + reward there too
                    obs, rew = h_obs, h_rew
        else:
            if tuple(actions.shape) != (n,):
                raise ValueError(f"actions must have shape ({n},), got {tuple(actions.shape)}")
            act = (actions if actions.dtype == torch.int32 else actions.to(torch.int32)).contiguous()
            p_act = _lib.ptr(act)
        arms = self._arm_table(1)
        if mode == "pipelined":
            d_act = _lib.ptr(self._device_scratch("action", (n,), torch.int32)) if self.host_stage_actions else None
            with torch.cuda.device(self.device):
                _lib.check(self._lib.ef_bandit_step_host(
                    self._pipe(), _lib.ptr(arms), self.k, p_act, d_act, self._seed, self._steps_since_reset, self.env_offset,
                    _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(h_obs), _lib.ptr(h_rew), _lib.ptr(self._stats_raw),
                    _lib.ptr(self.err), n, self.host_chunks, self._stream()))
            obs, rew = h_obs, h_rew
        else:
            self._call(self._lib.ef_bandit_step_means, _lib.ptr(arms), self.k, p_act, self._means_ptr(), self._seed,
                       self._steps_since_reset, self.env_offset, 1, _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(self._stats_raw),
                       _lib.ptr(self.err), n)
        self._advance(1)
        if self.strict:
            try:
                self.check_errors()
            except ValueError as e:
                raise AssertionError("Invalid action, must be between 0 and k-1.") from e
        if host_io:
            if mode != "staged":
                torch.cuda.current_stream(self.device).synchronize()
                return obs.numpy(), rew.numpy(), np.ones(n, bool), np.zeros(n, bool), {}
            obs, rew = self._to_host([("obs", self._obs), ("reward", self._reward)])
            return obs, rew, np.ones(n, bool), np.zeros(n, bool), {}
        return self._obs, self._reward, self._done, self._trunc, {}

    def rollout(self, K, actions=None):
        """K consecutive steps in one launch.  actions CUDA int32 [K,N] or None (uniform random arm).  Returns
        (obs int32 [K,N], reward fp32 [K,N])."""
        torch = self._torch
        n = self.num_envs
        if actions is not None:
            if not (torch.is_tensor(actions) and actions.is_cuda and actions.dtype == torch.int32
                    and tuple(actions.shape) == (K, n)):
                raise ValueError(f"actions must be a CUDA int32 tensor of shape ({K}, {n})")
            actions = actions.contiguous()
        if self.rng == "mt19937":      # one launch per step (the generators advance step by step); small-N mode
            if actions is None:
                raise ValueError("rng='mt19937' needs explicit actions (the in-kernel random policy is a Philox stream)")
            obs = torch.empty((K, n), dtype=torch.int32, device=self.device)
            rew = torch.empty((K, n), dtype=torch.float32, device=self.device)
            for j in range(K):
                o, r = self._step_mt19937(actions[j])[:2]
                obs[j].copy_(o)
                rew[j].copy_(r)
            return obs, rew
        arms = self._arm_table(K)
        obs = torch.empty((K, n), dtype=torch.int32, device=self.device)
        rew = torch.empty((K, n), dtype=torch.float32, device=self.device)
        self._call(self._lib.ef_bandit_step_means, _lib.ptr(arms), self.k, _lib.ptr(actions), self._means_ptr(), self._seed,
                   self._steps_since_reset, self.env_offset, int(K), _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(self._stats_raw),
                   _lib.ptr(self.err), n)
        self._advance(K)
        return obs, rew

    def _advance(self, K):
        self.timestep += K
        self._steps_since_reset += K
        self._t += K
        for arm in self.arms:  # :195-196 -- absolute shift, so one update at the final timestep suffices
            arm.update_param(self.timestep)

    def reset(self, seed=None):
        """KArmedBandit.reset (:202-217): re-key the generator, rebuild the arms (parameters back to their initial
        values); `timestep` is NOT reset.  Returns (zeros int32 [N], {})."""
        if seed is not None:
            self.seed = seed
        # the reference builds RandomState(self.seed): a fixed seed replays the same means and reward stream,
        # seed=None draws fresh entropy
        if self.rng == "mt19937" and seed is not None and not isinstance(seed, (int, np.integer)):
            self._seed_list = [int(x) for x in seed]
            if len(self._seed_list) != self.num_envs:
                raise ValueError(f"seed sequence must hold {self.num_envs} integers")
            self.seed = self._seed_list[0]
        elif seed is not None:
            self._seed_list = None
        self._seed = resolve_seed(self.seed)
        self._steps_since_reset = 0
        self._reset_epoch += 1
        self._init_arms()
        if self.rng == "mt19937":
            self._init_mt19937()
        return self._zero_obs, {}

    def _describe_error(self, count, env, detail, kind_id, kind):
        return f"Invalid action {detail}, must be between 0 and k-1 (env {env}; {count} rejected env-step(s))"
