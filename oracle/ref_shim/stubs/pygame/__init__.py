"""TEST INFRASTRUCTURE ONLY -- inert pygame stand-in."""


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
