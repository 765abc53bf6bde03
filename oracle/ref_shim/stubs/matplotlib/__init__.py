"""TEST INFRASTRUCTURE ONLY -- inert matplotlib stand-in (render paths are out of scope)."""
from . import pyplot, colors  # noqa: F401


def use(*a, **k):
    pass
