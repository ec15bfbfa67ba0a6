"""import-only stub (graphs/losses/rate_dist.py:11)."""


class Visdom:
    def __init__(self, *a, **k):
        pass
