"""Shared host code of VecCartPole / VecPendulumDisk (see cart_pole.py, pendulum_disk.py)."""
import numpy as np

from . import _lib, spaces
from .vec_env import VecEnv, resolve_seed


class VecPendulumBase(VecEnv):
    DIM = 0            # 4 (CartPole) / 2 (PendulumDisk)
    _FN = ""           # C-ABI prefix: "ef_cartpole" / "ef_pendulum_disk"
    _ACC_DIM = 0       # info accelerations per env

    def __init__(self, num_envs, tau, continuous_reward, episode_length_limit, angle_termination, initial_state,
                 nonlinear_reward, reward_decay_rate, reward_angle_range, device, seed, env_offset, autoreset, strict,
                 alias_obs):
        super().__init__(num_envs, device, seed, env_offset, autoreset)
        torch = self._torch
        self.tau = tau
        self.continuous_reward = continuous_reward
        self.nonlinear_reward = nonlinear_reward
        self.reward_decay_rate = reward_decay_rate
        self.episode_length_limit = episode_length_limit
        self.angle_termination = angle_termination
        self.initial_state = initial_state
        self.reward_angle_range = reward_angle_range
        self.strict = bool(strict)
        if not isinstance(episode_length_limit, (int, np.integer)):
            raise TypeError("episode_length_limit must be an int (the reference compares current_step >= limit)")
        self.single_action_space = spaces.Box(low=-3.0, high=3.0, shape=(1,), dtype=np.float32)
        self.action_space = spaces.Batched(self.single_action_space, num_envs)
        self.observation_space = spaces.Batched(self.single_observation_space, num_envs)
        n, dev = self.num_envs, self.device
        self._state = torch.zeros((n, self.DIM), dtype=torch.float32, device=dev)
        # The observation of these envs IS the (post-reset) state (cart_pole.py:166, pendulum_disk.py:158).  alias_obs=True
        # (default) returns the state tensor itself as the observation -- like every output of step(), valid until the next
        # call, not to be modified in place -- and saves writing DIM*4 bytes per env-step a second time; alias_obs=False
        # keeps a separate observation buffer.
        self._obs = self._state if alias_obs else torch.zeros((n, self.DIM), dtype=torch.float32, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._term = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._final_obs = None
        self._acc = None
        self._fast_args = None   # cached pointer arguments of the lean step() path
        self._params = self._make_params()
        self._fn_step = getattr(self._lib, self._FN + "_step")
        self._fn_rollout = getattr(self._lib, self._FN + "_rollout")
        self._fn_rollout_traj = getattr(self._lib, self._FN + "_rollout_traj")
        self._fn_reset = getattr(self._lib, self._FN + "_reset")
        self._fn_step_host = getattr(self._lib, self._FN + "_step_host")
        self.reset()

    # -- params -------------------------------------------------------------------------------------------------
    def _fill_common(self, p, angle_threshold):
        p.tau = self.tau
        p.angle_threshold = angle_threshold
        p.angle_termination = -1.0 if self.angle_termination is None else float(self.angle_termination)
        p.reward_mode = 0 if not self.continuous_reward else (2 if self.nonlinear_reward else 1)
        p.reward_lo, p.reward_hi = float(self.reward_angle_range[0]), float(self.reward_angle_range[1])
        p.reward_decay = float(self.reward_decay_rate)
        p.reward_inv_span = 1.0 / angle_threshold
        denom = 1.0 - np.exp(-float(self.reward_decay_rate) * angle_threshold)
        p.reward_inv_norm = float(1.0 / denom) if denom != 0 else float("inf")
        p.episode_limit = int(self.episode_length_limit)
        p.init_lo, p.init_hi = -0.01, 0.01  # _initialize_state: np.random.uniform(-0.01, 0.01)
        self._set_initial_state(p, self.initial_state)

    def _set_initial_state(self, p, initial_state):
        if initial_state is None:
            p.fixed_init = 0
            return
        v = np.asarray(initial_state, dtype=np.float64).reshape(-1)
        if v.shape[0] != self.DIM:
            raise ValueError(f"initial_state must have {self.DIM} components")
        p.fixed_init = 1
        for i in range(self.DIM):
            p.init_state[i] = float(v[i])

    # -- reference-style attributes -----------------------------------------------------------------------------------
    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, value):
        torch = self._torch
        v = torch.as_tensor(np.asarray(value, dtype=np.float32) if not torch.is_tensor(value) else value)
        v = v.to(device=self.device, dtype=torch.float32).reshape(-1, self.DIM)
        self._state.copy_(v.expand(self.num_envs, self.DIM) if v.shape[0] == 1 else v)

    @property
    def current_step(self):
        return self.ep_len

    @property
    def total_reward(self):
        return self.ep_ret

    def get_state(self):
        return {"state": self._state.clone(), "ep_len": self.ep_len.clone(), "ep_ret": self.ep_ret.clone(), "t": self._t}

    def set_state(self, s):
        self._state.copy_(s["state"])
        self.ep_len.copy_(s["ep_len"])
        self.ep_ret.copy_(s["ep_ret"])
        self._t = int(s.get("t", self._t))

    # -- hot path ---------------------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None, initial_state=None, mask=None):
        """reset (cart_pole.py:91-105 / pendulum_disk.py:94-106).  initial_state: one state for all envs or [N,DIM]."""
        torch = self._torch
        if seed is not None:
            self.seed, self._seed = seed, resolve_seed(seed)
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        per_env = None
        params = self._params
        if initial_state is not None:
            arr = initial_state if torch.is_tensor(initial_state) else np.asarray(initial_state, dtype=np.float64)
            if arr.ndim == 2 and arr.shape[0] == self.num_envs and self.num_envs > 1:
                per_env = torch.as_tensor(arr).to(device=self.device, dtype=torch.float32)
            else:
                params = self._make_params()
                self._set_initial_state(params, np.asarray(arr.cpu() if torch.is_tensor(arr) else arr).reshape(-1))
        self._call(self._fn_reset, params, _lib.ptr(self._state), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret),
                   self._seed, self._reset_epoch, self.env_offset, _lib.ptr(m), self.num_envs)
        self._reset_epoch += 1
        if per_env is not None:
            if m is None:
                self._state.copy_(per_env)
            else:
                self._state.copy_(torch.where(m.bool()[:, None], per_env, self._state))
        if self._obs is not self._state:
            self._obs.copy_(self._state)
        return self._obs, {}

    def check_errors(self, clear=True):
        """Raise AssertionError if any env-step since the last check carried an action the reference asserts on
        (`assert self.action_space.contains(action)`, cart_pole.py:108 / pendulum_disk.py:109): NaN or outside [-3, 3].
        Such env-steps were rejected on the device: env untouched, reward 0, flags 0, counted in `err`."""
        e = [int(x) & 0xFFFFFFFF for x in self.err.cpu().tolist()]
        if clear and e[0]:
            self.err.zero_()
        if e[0]:
            packed = e[2] | (e[3] << 32)      # (1 + global env index) << 32 | kind << 24 | bits 8-31 of the fp32 action
            env, detail = (packed >> 32) - 1, packed & 0xFFFFFF
            approx = float(np.array([detail << 8], dtype=np.uint32).view(np.float32)[0])
            raise AssertionError(f"{e[0]} env-step(s) rejected: action outside [-3, 3] invalid; first offender env {env}, "
                                 f"action ~ {approx!r}")

    def _check_actions_strict(self, a):
        # Box(-3, 3, (1,), float32).contains (cart_pole.py:108): float32-castable dtype and |a| <= 3
        torch = self._torch
        if a.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise AssertionError(f"{a.dtype} actions invalid: not castable to float32")
        if bool((a.abs() > 3.0).any()) or bool(torch.isnan(a).any()):
            raise AssertionError("action outside [-3, 3] invalid")

    def step(self, actions, *, return_final_obs=False, return_info=False):
        """step (cart_pole.py:107-180 / pendulum_disk.py:108-172) for all envs.  actions: fp32 [N] or [N,1]."""
        torch = self._torch
        n = self.num_envs
        if (not return_final_obs and not return_info and not self.strict and torch.is_tensor(actions) and actions.is_cuda
                and actions.dtype == torch.float32 and actions.numel() == n and actions.is_contiguous()):
            # lean path of the usual call (CUDA fp32 actions in, CUDA tensors out): pointer arguments cached
            fixed = self._fast_args
            if fixed is not None and (fixed[3] is not self.ep_len or fixed[4] is not self.ep_ret):
                fixed = None     # a counter tensor was replaced by assignment: rebuild the cached pointers
            if fixed is None:
                fixed = self._fast_args = (
                    (self._params, self._state.data_ptr(), self.ep_len.data_ptr(), self.ep_ret.data_ptr()),
                    (self._obs.data_ptr(), self._reward.data_ptr(), self._term.data_ptr(), self._trunc.data_ptr(), None, None,
                     self._stats_raw.data_ptr(), self.err.data_ptr()),
                    (self._obs, self._reward, self._term.view(torch.bool), self._trunc.view(torch.bool)),
                    self.ep_len, self.ep_ret)
            head, tail, outs = fixed[:3]
            clock = self._clock
            self._call_lean(self._fn_step, *head, actions.data_ptr(), self._seed, 0 if clock is not None else self._t,
                            self.env_offset, *tail, None if clock is None else clock.data_ptr(), n, int(self.autoreset))
            self._t += 1
            return outs[0], outs[1], outs[2], outs[3], {}
        host_io = not (torch.is_tensor(actions) and actions.is_cuda)
        mode = self.host_transport if host_io else None
        if mode == "pipelined" and self._clock is not None:
            mode = "staged"
        zero_copy = mode == "zerocopy"
        if host_io:
            if torch.is_tensor(actions):
                a = actions.reshape(-1)
                if self.strict:
                    self._check_actions_strict(a)
                a = a.to(torch.float32)
            else:
                a = np.asarray(actions)
                if self.strict and a.dtype == np.float64:
                    raise AssertionError("float64 actions invalid: not castable to float32")
                a = a.reshape(-1).astype(np.float32, copy=False)
                if self.strict:
                    self._check_actions_strict(torch.from_numpy(a))
            h_act = self._host_in("action", a, torch.float32, (n,))
            p_act = (_lib.ptr(self._to_device("action", h_act, torch.float32, (n,))) if mode == "staged"
                     else h_act.data_ptr())
        else:
            act = actions.reshape(-1)
            if act.shape[0] != n:
                raise ValueError(f"actions must hold {n} values, got {tuple(actions.shape)}")
            if self.strict:
                self._check_actions_strict(act)
            act = act.to(torch.float32).contiguous()
            p_act = _lib.ptr(act)
        fobs = acc = None
        if zero_copy:
            # outputs are written by the kernel straight into pinned host memory (UVA pointers): no staging copies
            obs = self._pinned("o_obs", (n, self.DIM), torch.float32)
            rew = self._pinned("o_reward", (n,), torch.float32)
            term = self._pinned("o_term", (n,), torch.uint8)
            trunc = self._pinned("o_trunc", (n,), torch.uint8)
            if return_final_obs:
                fobs = self._pinned("o_final_obs", (n, self.DIM), torch.float32)
            if return_info:
                acc = self._pinned("o_acc", (n, self._ACC_DIM), torch.float32)
        else:
            obs, rew, term, trunc = self._obs, self._reward, self._term, self._trunc
            if return_final_obs:
                if self._final_obs is None:
                    self._final_obs = torch.zeros((n, self.DIM), dtype=torch.float32, device=self.device)
                fobs = self._final_obs
            if return_info:
                if self._acc is None:
                    self._acc = torch.zeros((n, self._ACC_DIM), dtype=torch.float32, device=self.device)
                acc = self._acc
        if mode == "pipelined":
            # ef_*_step_host: chunked; kernels write the device buffers, the copy engine drains them into pinned memory
            hosts = [self._pinned("o_obs", (n, self.DIM), torch.float32), self._pinned("o_reward", (n,), torch.float32),
                     self._pinned("o_term", (n,), torch.uint8), self._pinned("o_trunc", (n,), torch.uint8),
                     self._pinned("o_final_obs", (n, self.DIM), torch.float32) if return_final_obs else None,
                     self._pinned("o_acc", (n, self._ACC_DIM), torch.float32) if return_info else None]
            d_act = _lib.ptr(self._device_scratch("action", (n,), torch.float32)) if self.host_stage_actions else None
            with torch.cuda.device(self.device):
                _lib.check(self._fn_step_host(
                    self._pipe(), self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret), p_act,
                    d_act, self._seed, self._t, self.env_offset, _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(term),
                    _lib.ptr(trunc), _lib.ptr(fobs), _lib.ptr(acc), *[_lib.ptr(h) for h in hosts],
                    _lib.ptr(self._stats_raw), _lib.ptr(self.err), n, int(self.autoreset), self.host_chunks, self._stream()))
            obs, rew, term, trunc, fobs, acc = hosts
        else:
            self._call(self._fn_step, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret),
                       p_act, self._seed, self._t_arg(), self.env_offset, _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(term),
                       _lib.ptr(trunc), _lib.ptr(fobs), _lib.ptr(acc), _lib.ptr(self._stats_raw), _lib.ptr(self.err),
                       _lib.ptr(self._clock), n, int(self.autoreset))
        self._t += 1
        info = {}
        if host_io:
            if mode != "staged":
                torch.cuda.current_stream(self.device).synchronize()
                out = [t.numpy() for t in (obs, rew, term, trunc)]
                out += [fobs.numpy()] if return_final_obs else []
                out += [acc.numpy()] if return_info else []
            else:
                named = [("obs", self._obs), ("reward", self._reward), ("term", self._term), ("trunc", self._trunc)]
                if return_final_obs:
                    named.append(("final_obs", fobs))
                if return_info:
                    named.append(("acc", acc))
                out = self._to_host(named)
            k = 4
            if return_final_obs:
                info["final_observation"] = out[k]
                k += 1
            if return_info:
                info.update(self._info(out[k], out[3].view(np.bool_), h_act.numpy()))
            return out[0], out[1], out[2].view(np.bool_), out[3].view(np.bool_), info
        if return_final_obs:
            info["final_observation"] = fobs
        if return_info:
            info.update(self._info(acc, self._trunc.view(torch.bool), act))
            info["steps"] = self.ep_len
        return self._obs, self._reward, self._term.view(torch.bool), self._trunc.view(torch.bool), info

    def rollout(self, K, actions=None, *, return_rewards=False, return_obs=False):
        """K fused steps (state in registers, auto-reset on).  actions: CUDA fp32 [K,N] or None (U(-3,3) Philox policy).
        Returns None, (rewards fp32 [K,N], flags uint8 [K,N]: bit0 terminated, bit1 truncated) if return_rewards, or
        (obs fp32 [K,N,DIM], rewards, flags) if return_obs -- the observations K step() calls would have returned."""
        torch = self._torch
        n = self.num_envs
        if actions is not None:
            if not (torch.is_tensor(actions) and actions.is_cuda and actions.dtype == torch.float32
                    and tuple(actions.shape) == (K, n)):
                raise ValueError(f"actions must be a CUDA float32 tensor of shape ({K}, {n})")
            actions = actions.contiguous()
        rew = flg = obs = None
        if return_rewards or return_obs:
            rew = torch.empty((K, n), dtype=torch.float32, device=self.device)
            flg = torch.empty((K, n), dtype=torch.uint8, device=self.device)
        if return_obs:
            obs = torch.empty((K, n, self.DIM), dtype=torch.float32, device=self.device)
            self._call(self._fn_rollout_traj, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret),
                       _lib.ptr(actions), self._seed, self._t_arg(), self.env_offset, int(K), _lib.ptr(obs), _lib.ptr(rew),
                       _lib.ptr(flg), _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), n)
        else:
            self._call(self._fn_rollout, self._params, _lib.ptr(self._state), _lib.ptr(self.ep_len), _lib.ptr(self.ep_ret),
                       _lib.ptr(actions), self._seed, self._t_arg(), self.env_offset, int(K), _lib.ptr(rew), _lib.ptr(flg),
                       _lib.ptr(self._stats_raw), _lib.ptr(self.err), _lib.ptr(self._clock), n)
        self._t += int(K)
        if self._obs is not self._state:
            self._obs.copy_(self._state)
        if return_obs:
            return obs, rew, flg
        return (rew, flg) if return_rewards else None
