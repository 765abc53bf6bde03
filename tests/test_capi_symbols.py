"""The C-ABI library loads without a GPU and exports exactly the entry points include/envforge_b200.h declares."""
import ctypes
import os
import re

import pytest

from rl_envs_forge_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "envforge_b200.h")


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ef_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_status_strings(lib):
    assert lib.ef_abi_version() == _lib.EF_ABI_VERSION
    assert lib.ef_status_string(0) == b"EF_OK"
    assert lib.ef_status_string(1) == b"EF_ERR_INVALID_ARGUMENT"


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.GridDesc) == 48
    assert _lib.GridDesc.table.offset == 40
    assert ctypes.sizeof(_lib.LabyrinthDesc) == 64 and _lib.LabyrinthDesc.grids.offset == 32
    assert ctypes.sizeof(_lib.GamblerParams) == 24 and _lib.GamblerParams.win_threshold.offset == 16
    assert ctypes.sizeof(_lib.CarRentalParams) == 248 and _lib.CarRentalParams.lam.offset == 24
    assert ctypes.sizeof(_lib.PendulumParams) == 4 + 32 + 4 * 4 + 4 * 4 + 4 * 4 + 16  # 100 bytes


def test_argument_validation_without_gpu(lib):
    # null descriptor / null state are rejected before any CUDA call is made
    assert lib.ef_gridworld_step(None, None, None, None, None, None, 0, 0, 0, None, None, None, None, None, None,
                                 None, None, 8, 1, None) == 1
    d = _lib.GridDesc()
    assert lib.ef_gridworld_rollout(ctypes.byref(d), None, None, None, None, 0, 0, 0, 4, None, None, None, None, None,
                                    8, None) == 1
    assert lib.ef_cartpole_step(None, None, None, None, None, 0, 0, 0, None, None, None, None, None, None, None, None,
                                None, 8, 1, None) == 1
    assert lib.ef_bandit_step(None, 10, None, 0, 0, 0, 1, None, None, None, None, 8, None) == 1
    # packed GridWorld state (ep_len == NULL) is only accepted with auto-reset on and 0 < episode_limit <= 65535
    d.rows = d.cols = 2
    d.n_slots, d.table = 1, 0x1000
    fake = ctypes.c_void_p(0x2000)
    for limit, autoreset, want in ((0, 1, 1), (70000, 1, 1), (10, 0, 1), (10, 1, 0)):
        d.episode_limit = limit
        got = lib.ef_gridworld_step(ctypes.byref(d), fake, None, fake, None, None, 0, 0, 0, None, None, None, None, None,
                                    None, None, None, 0, autoreset, None)          # n == 0: returns before any launch
        assert got == want, (limit, autoreset, got)
    # host-buffer entry points: a missing pipe / host action buffer is rejected up front
    assert lib.ef_bandit_step_host(None, None, 10, None, None, 0, 0, 0, None, None, None, None, None, None, 8, 4, None) == 1
    assert lib.ef_gridworld_step_host(None, ctypes.byref(d), *([None] * 5), 0, 0, 0, *([None] * 12), 8, 1, 4, None) == 1
    assert lib.ef_host_pipe_create(None) == 1 and lib.ef_host_pipe_destroy(None) == 0
    # narrow wire format: same descriptor checks, grids wider than a byte are refused before any launch
    d.episode_limit = 10
    assert lib.ef_gridworld_step_u8(None, *([None] * 5), 0, 0, 0, *([None] * 8), 8, 1, None) == 1
    assert lib.ef_gridworld_step_u8(ctypes.byref(d), fake, None, fake, None, None, 0, 0, 0, *([None] * 8), 0, 1, None) == 0
    d.rows, d.cols = 300, 2
    assert lib.ef_gridworld_step_u8(ctypes.byref(d), fake, None, fake, None, None, 0, 0, 0, *([None] * 8), 8, 1, None) == 3
    d.rows = d.cols = 2
    assert lib.ef_gridworld_step_u8_host(None, ctypes.byref(d), *([None] * 5), 0, 0, 0, *([None] * 12), 8, 1, 4, None) == 1
    x = _lib.GridStepArgs()
    x.grid, x.pos, x.ep_ret, x.n, x.autoreset, x.wire = ctypes.pointer(d), 0x2000, 0x2000, 0, 1, _lib.EF_WIRE_U8
    assert lib.ef_gridworld_step_v(ctypes.byref(x)) == 0 and lib.ef_gridworld_step_v(None) == 1
    # mixed-layout rollout and the stream-exact bandit step: null state / generator buffers are rejected
    assert lib.ef_gridworld_rollout_layouts(ctypes.byref(d), 2, None, None, *([None] * 4), 0, 0, 0, 4, *([None] * 6), 8, None) == 1
    assert lib.ef_gridworld_rollout_layouts(ctypes.byref(d), 0, None, fake, fake, None, fake, None, 0, 0, 0, 4, *([None] * 6), 8,
                                            None) == 1
    assert lib.ef_bandit_step_mt19937(None, 10, *([None] * 6), 0, *([None] * 5), 8, None) == 1
    assert lib.ef_bandit_step_mt19937(fake, 10, fake, fake, fake, fake, fake, None, 0, *([None] * 5), 0, None) == 0   # n == 0
    assert lib.ef_bandit_step_mt19937(fake, 10, fake, fake, fake, fake, fake, None, 2 ** 32, *([None] * 5), 8, None) == 1


def test_no_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rl_envs_forge_b200 as r
    for cls in (r.VecGridWorld, r.VecCartPole, r.VecPendulumDisk, r.VecKArmedBandit, r.VecGamblersProblem,
                r.VecJacksCarRental):
        with pytest.raises(r.EnvForgeError):
            cls(4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rl_envs_forge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_header_is_plain_c_and_the_c_host_links(lib, tmp_path):
    """include/envforge_b200.h must compile as C11 (no C++ / torch types in the signatures), and the plain-C host of the ABI
    (tests/c/abi_host.c, run on the GPU box by tests/test_gpu_c_abi_host.py) must build and link against the library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "envforge_b200.h"\nint main(void) { return sizeof(ef_grid_desc) == 48 && sizeof(ef_arm64) == 24 '
                   '&& sizeof(ef_gridworld_step_args) == 160 && sizeof(ef_car_rental_params) == 248 ? 0 : 1; }\n')
    exe = tmp_path / "hdr"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}", str(src), "-o",
                    str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
    assert ctypes.sizeof(_lib.GridStepArgs) == 160
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(os.path.join(ROOT, "tests", "c", "abi_host"))
