"""import-only stub (loggers/rate_dist.py)."""


class _Noop:
    def __init__(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


class visdomlogger:
    VisdomPlotLogger = _Noop
    VisdomLogger = _Noop


VisdomPlotLogger = _Noop
VisdomLogger = _Noop
