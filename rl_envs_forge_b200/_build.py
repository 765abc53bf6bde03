"""Builds rl_envs_forge_b200/csrc/*.cu into the in-tree C-ABI library libenvforge_b200.so (sm_100a only).

    python -m rl_envs_forge_b200._build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libenvforge_b200.so")
SOURCES = ["capi.cu", "stats.cu", "gridworld.cu", "pendulums.cu", "bandit.cu", "bandit_mt19937.cu", "acml.cu", "labyrinth.cu", "hostpipe.cu"]
HEADERS = ["ef_common.cuh", os.path.join("..", "..", "include", "envforge_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC", "-cudart", "static"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libenvforge_b200.so")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), tag=""):
    """Compile and link.  `defines` (e.g. ["EF_STEP_MINBLOCKS=4"]) + `tag` build a tuning variant next to the default
    library (libenvforge_b200<tag>.so), selected at run time with ENVFORGE_LIB."""
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    lib = LIB[:-3] + tag + ".so"
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = s[:-3] + tag + ".o"
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = ([nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
                   + ["-c", s, "-o", o])
            if verbose:
                print(" ".join(cmd))
            jobs.append(cmd)
    if jobs:  # translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            for r in pool.map(lambda c: subprocess.run(c, check=False), jobs):
                if r.returncode != 0:
                    raise subprocess.CalledProcessError(r.returncode, r.args)
    if force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-o", lib] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    tags = [a[6:] for a in sys.argv[1:] if a.startswith("--tag=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, tag=tags[0] if tags else ""))
