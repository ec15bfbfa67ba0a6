"""Import the UNMODIFIED reference (``/root/reference``) on top of ``oracle/shims``.

TEST INFRASTRUCTURE.  Only usable in the dev container (the GPU box has no
``/root/reference``); used by ``tests/golden/make_golden.py`` to write the golden
vectors and by ``-m "not gpu"`` tests that re-check them when the reference is
present.  Nothing from the reference is copied: it is imported where it lies.
"""
import importlib
import json
import os
import sys

REFERENCE_ROOT = os.environ.get("LL_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
_REPO = os.path.dirname(_HERE)


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "graphs", "models", "LiftingBasedDWT_net.py"))


def load():
    """Return the reference's ``graphs.models.LiftingBasedDWT_net`` module."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("graphs.models.LiftingBasedDWT_net")


def default_config(**overrides):
    """``liftingDWT.json`` of the reference as an EasyDict, ``mode='test'`` so that
    every tensor stays on the CPU (the reference hard-codes ``.cuda()`` otherwise:
    wavelet_inverse_v2.py:48-51, lifting_dwt_nets.py:749-759)."""
    load()
    from easydict import EasyDict
    with open(os.path.join(REFERENCE_ROOT, "liftingDWT.json")) as f:
        cfg = EasyDict(json.load(f))
    cfg.mode = "test"
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg
