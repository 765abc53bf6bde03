"""The persistent step server (`ef_gridworld_serve`): K GridWorld steps of one env set without a launch per step, the
actions coming from a policy kernel of its own that runs concurrently and talks to the server through sequence words.
Held to the plain `step()` path: same actions -> bit-identical trajectories, outputs and statistics."""
import numpy as np
import pytest
import torch

from oracle.gen_golden import CONFIG2_LAYOUT
from tests.test_gpu_gridworld import _vec_kwargs

pytestmark = pytest.mark.gpu


def _same_stats(a, b):
    """Integer entries (episodes, length sum, env-steps) exact; the float64 sums of fp32 returns agree up to the order of
    additions (the server and the step kernel give an env to different threads)."""
    sa, sb = a.stats.cpu(), b.stats.cpu()
    for i in (0, 1, 4):
        assert sa[i] == sb[i]
    assert torch.allclose(sa, sb, rtol=1e-12, atol=0)


def _demo_policy(obs, k):
    return ((obs[:, 0] * 3 + obs[:, 1] + k) & 3).to(torch.int32)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n,K", [(4096, 40), (1 << 20, 64), (1 << 18, 300)])
def test_served_steps_equal_plain_steps(n, K):
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _vec_kwargs(dict(CONFIG2_LAYOUT, p_success=0.8))
    served = VecGridWorld(n, seed=21, **cfg)
    plain = VecGridWorld(n, seed=21, **cfg)
    obs_s, rew_s, term_s, trunc_s = served.serve_with_demo_policy(K)
    obs = plain.reset()[0]
    for k in range(K):
        obs, rew, term, trunc, _ = plain.step(_demo_policy(obs, k))
    assert torch.equal(obs_s, obs) and torch.equal(rew_s, rew) and torch.equal(term_s, term) and torch.equal(trunc_s, trunc)
    assert torch.equal(served.pos, plain.pos) and torch.equal(served.ep_len, plain.ep_len)
    assert torch.equal(served.ep_ret, plain.ep_ret)
    _same_stats(served, plain)
    assert float(served.stats[0]) > 0
    assert int(served.err[0].item()) == int(plain.err[0].item())
    assert served.step_count == plain.step_count == K
    # the env goes on with ordinary steps (and another served stretch) from where the server stopped
    a = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")
    o1, o2 = served.step(a), plain.step(a)
    assert torch.equal(o1[0], o2[0]) and torch.equal(o1[1], o2[1])
    served.serve_with_demo_policy(5)
    obs = o2[0]
    for k in range(5):
        obs = plain.step(_demo_policy(obs, k))[0]
    assert torch.equal(served.pos, plain.pos)
    _same_stats(served, plain)


@pytest.mark.timeout(120)
def test_server_gives_up_without_a_producer():
    """No producer running: the server must not hang the GPU -- it gives up after ~5 s, reports it, and leaves the env
    exactly as it was."""
    from rl_envs_forge_b200 import EnvForgeError, VecGridWorld
    cfg = _vec_kwargs(dict(CONFIG2_LAYOUT, p_success=0.8))
    env = VecGridWorld(4096, seed=1, **cfg)
    before = env.pos.clone()
    env.serve_begin(3)
    with pytest.raises(EnvForgeError, match="gave up"):
        env.serve_end()
    assert torch.equal(env.pos, before) and env.step_count == 0
    env.step(torch.zeros(4096, dtype=torch.int32, device="cuda"))      # still usable


@pytest.mark.timeout(120)
def test_largest_shared_memory_table_is_served_or_refused_never_hung():
    """A 24x32 grid with slip has the largest table the kernels stage in shared memory (96 KB).  The server keeps 17 KB of
    its own next to it and needs EVERY chunk's CTA resident at once: it must either serve that grid (bit-identical to plain
    steps) or refuse it with EF_ERR_UNSUPPORTED before launching -- never start a grid that cannot be co-resident."""
    from rl_envs_forge_b200 import VecGridWorld
    from rl_envs_forge_b200._lib import EnvForgeError
    n, K = 1 << 16, 12
    cfg = dict(rows=24, cols=32, start_state=(0, 0), terminal_states={(23, 31): 1.0}, p_success=0.7,
               episode_length_limit=50)
    served = VecGridWorld(n, seed=5, **cfg)
    plain = VecGridWorld(n, seed=5, **cfg)
    try:
        obs_s, rew_s, term_s, trunc_s = served.serve_with_demo_policy(K)
    except EnvForgeError as e:
        assert "unsupported" in str(e).lower()
        return
    obs = plain.reset()[0]
    for k in range(K):
        obs, rew, term, trunc, _ = plain.step(_demo_policy(obs, k))
    assert torch.equal(obs_s, obs) and torch.equal(rew_s, rew) and torch.equal(term_s, term) and torch.equal(trunc_s, trunc)
    assert torch.equal(served.pos, plain.pos) and torch.equal(served.ep_ret, plain.ep_ret)
