"""The C ABI driven by a plain C program (tests/c/abi_host.c: no Python, no torch; device buffers from cudaMalloc, default
stream): its outputs must equal VecGridWorld.step on the same table, seed and actions bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from oracle.gen_golden import CONFIG2_LAYOUT

pytestmark = pytest.mark.gpu
CDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")


def _binary():
    from rl_envs_forge_b200 import _build
    _build.build()
    subprocess.run(["make", "-C", CDIR], check=True, capture_output=True)
    return os.path.join(CDIR, "abi_host")


@pytest.mark.parametrize("p,packed,n", [(0.8, 1, 20_003), (1.0, 0, 4096)])
def test_plain_c_host_equals_python_wrapper(tmp_path, p, packed, n):
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorld
    steps, seed, limit = 60, 321, 25
    cfg = dict(CONFIG2_LAYOUT, p_success=p, episode_length_limit=limit)
    cfg["special_transitions"] = {(k[0], Action(int(k[1]))): v for k, v in cfg["special_transitions"].items()}
    env = VecGridWorld(n, seed=seed, state_layout="packed" if packed else "split", **cfg)
    d = env._desc
    tab = env.spec.device_table(extra_cells=1)
    rng = np.random.default_rng(n)
    actions = rng.integers(0, 4, (steps, n)).astype(np.int32)
    actions[rng.random((steps, n)) < 0.002] = 6                       # a few invalid actions: counted, env untouched
    header = np.array([d.rows, d.cols, d.n_slots, d.idx_cap, d.thr[0], d.thr[1], d.thr[2], d.start_cell, d.episode_limit,
                       d.table_cells, n, steps, seed, packed, 0, 0], np.int64).astype(np.uint32).view(np.int32)
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(src, "wb") as f:
        f.write(header.tobytes())
        f.write(np.ascontiguousarray(tab).tobytes())
        f.write(actions.tobytes())
    r = subprocess.run([_binary(), str(src), str(dst)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("ok:")
    raw = np.fromfile(dst, np.uint8)
    per = 8 * n + 4 * n + n + n
    for t in range(steps):
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(actions[t]).cuda())
        blk = raw[t * per:(t + 1) * per]
        assert np.array_equal(blk[:8 * n].view(np.int32).reshape(n, 2), obs.cpu().numpy()), t
        assert np.array_equal(blk[8 * n:12 * n].view(np.float32), rew.cpu().numpy()), t
        assert np.array_equal(blk[12 * n:13 * n].view(np.bool_), term.cpu().numpy()), t
        assert np.array_equal(blk[13 * n:14 * n].view(np.bool_), trunc.cpu().numpy()), t
    tail = raw[steps * per:]
    stats = tail[:64].view(np.float64)
    err = tail[64:80].view(np.uint32)
    mine = env.stats.cpu().numpy()
    assert stats[0] == mine[0] and stats[1] == mine[1] and stats[4] == mine[4] and stats[0] > 0
    assert np.allclose(stats[2:4], mine[2:4], rtol=1e-12)
    assert int(err[0]) == int(env.err.cpu().numpy()[0]) > 0
