"""TEST INFRASTRUCTURE ONLY -- inert seaborn stand-in (the bandit module calls sns.set() at import)."""


def set(*a, **k):
    pass


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
