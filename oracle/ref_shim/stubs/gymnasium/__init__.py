"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for gymnasium 0.29.1 (poetry.lock:443-444 of the
reference), which is not installed in this image.  Only the pieces the four hot-path reference
modules touch are provided: Env (np_random property), spaces.{Discrete,Box,Tuple},
utils.seeding.np_random.  Written from gymnasium's documented behaviour, not copied from it."""
import numpy as np
from . import spaces, utils  # noqa: F401
from .utils import seeding


class Env:
    metadata = {}
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, _ = seeding.np_random(seed)

    def close(self):
        pass
