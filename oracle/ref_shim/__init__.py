"""TEST INFRASTRUCTURE ONLY.

Import shim that makes the *unmodified* reference (mariusdgm/rl-envs-forge, mounted read-only at
/root/reference) importable in the build container, where gymnasium / matplotlib / pygame / seaborn
are not installed and the package itself is not pip-installed (rl_envs_forge/__init__.py:1-3 calls
importlib.metadata.version and would raise PackageNotFoundError).

Used only by oracle/gen_golden.py and the `needs_reference` tests.  /root/reference does not exist
on the GPU box; nothing that runs there may call `load_reference()`.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ENVFORGE_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rl_envs_forge", "envs"))


# oracle/_ref/: the four hot-path files of the reference, copied there by oracle/build_ref.py in the build container
# (git-ignored; it travels to the GPU box with the repo snapshot).  Enough for bench.py's CPU arm, not for gen_golden.
REF_COPY_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref")


def _install_stubs_and_package(root):
    for name in ("gymnasium", "matplotlib", "pygame", "seaborn"):
        try:
            importlib.import_module(name)
        except ImportError:
            if _STUBS not in sys.path:
                sys.path.insert(0, _STUBS)
            importlib.import_module(name)
    if "rl_envs_forge" not in sys.modules:
        pkg = types.ModuleType("rl_envs_forge")
        pkg.__path__ = [os.path.join(root, "rl_envs_forge")]
        pkg.__version__ = "5.22.0"
        sys.modules["rl_envs_forge"] = pkg


def load_reference_gridworld():
    """(GridWorld, Action, root) of the UNMODIFIED reference from /root/reference or, on the GPU box, from oracle/_ref;
    None when neither exists."""
    for root in (REFERENCE_ROOT, REF_COPY_ROOT):
        if os.path.isfile(os.path.join(root, "rl_envs_forge", "envs", "grid_world", "grid_world.py")):
            _install_stubs_and_package(root)
            gw = importlib.import_module("rl_envs_forge.envs.grid_world.grid_world")
            return gw.GridWorld, gw.Action, root
    return None


def load_reference():
    """Return a namespace with the reference env classes of the hot path, imported from REFERENCE_ROOT."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("gymnasium", "matplotlib", "pygame", "seaborn"):
        try:
            importlib.import_module(name)
        except ImportError:
            if _STUBS not in sys.path:
                sys.path.insert(0, _STUBS)
            importlib.import_module(name)
    if "rl_envs_forge" not in sys.modules:
        pkg = types.ModuleType("rl_envs_forge")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "rl_envs_forge")]
        pkg.__version__ = "5.22.0"
        sys.modules["rl_envs_forge"] = pkg
    ns = types.SimpleNamespace()
    gw = importlib.import_module("rl_envs_forge.envs.grid_world.grid_world")
    kb = importlib.import_module("rl_envs_forge.envs.k_armed_bandit.k_armed_bandit")
    cp = importlib.import_module("rl_envs_forge.envs.inverted_pendulum.cart_pole.cart_pole")
    pd = importlib.import_module("rl_envs_forge.envs.inverted_pendulum.pendulum_disk.pendulum_disk")
    ns.GridWorld, ns.Action = gw.GridWorld, gw.Action
    ns.KArmedBandit, ns.Arm = kb.KArmedBandit, kb.Arm
    ns.CartPole = cp.CartPole
    ns.PendulumDisk = pd.PendulumDisk
    gp = importlib.import_module("rl_envs_forge.envs.acml.gambler_problem.gambler_problem")
    cr = importlib.import_module("rl_envs_forge.envs.acml.car_rental.car_rental")
    ns.GamblersProblem, ns.JacksCarRental = gp.GamblersProblem, cr.JacksCarRental
    lb = importlib.import_module("rl_envs_forge.envs.labyrinth.labyrinth")
    ns.Labyrinth = lb.Labyrinth
    return ns
