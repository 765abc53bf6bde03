"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe for `oracle/_ref/` (git-ignored, NOT gpurun-ignored).

The reference is pure Python: nothing to compile.  When the reference tree is present (the build container), this copies
the four hot-path source files -- and nothing else -- from where they lie under /root/reference into `oracle/_ref/`, keeping
their package paths, so that `bench.py --impl reference` can time the UNMODIFIED `GridWorld.step` on the GPU box's host
cores (`cpu_baseline.kind == "reference"`).  The files never enter the git history; the product never imports them.

    python -m oracle.build_ref        (also run by __graft_entry__.build())
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ENVFORGE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = (  # the files of the hot path
    "rl_envs_forge/envs/grid_world/grid_world.py",
    "rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py",
    "rl_envs_forge/envs/inverted_pendulum/cart_pole/cart_pole.py",
    "rl_envs_forge/envs/inverted_pendulum/pendulum_disk/pendulum_disk.py",
)


def build_ref():
    """Returns the output directory, or None when the reference tree is absent (the GPU box: the prebuilt copy is used)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "rl_envs_forge", "envs")):
        return OUT if os.path.isdir(OUT) else None
    for rel in FILES:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
        d = os.path.dirname(dst)
        while os.path.abspath(d) != os.path.abspath(OUT):   # the reference's package __init__ files are empty
            init = os.path.join(d, "__init__.py")           # (the root one needs package metadata: the shim replaces it)
            if not os.path.exists(init):
                open(init, "w").close()
            d = os.path.dirname(d)
    return OUT


if __name__ == "__main__":
    print(build_ref())
