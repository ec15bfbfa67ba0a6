"""import-only stub: plotting is off the hot path (lifting_dwt_nets.py:391-412)."""


def __getattr__(name):
    def _noop(*a, **k):
        return None
    return _noop
