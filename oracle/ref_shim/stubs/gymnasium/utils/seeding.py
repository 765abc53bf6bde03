"""TEST INFRASTRUCTURE ONLY -- gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))."""
import numpy as np


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, int) and 0 <= seed):
        raise ValueError(f"Seed must be a python integer >= 0, actual: {seed!r}")
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    return np.random.Generator(np.random.PCG64(seed_seq)), np_seed
