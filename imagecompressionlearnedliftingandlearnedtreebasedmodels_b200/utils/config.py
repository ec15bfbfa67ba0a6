"""Config surface of the codec (reference: utils/config.py:50-103, liftingDWT.json:1-53, main.py:7-33).

``get_config_from_json(path) -> (config, config_dict)`` and ``process_config(config)`` keep the reference's
names, return values and side effects (experiment directories under ``experiments/<exp_name>/`` and the
console + two rotating-file log handlers); ``EasyDict`` is a local stand-in for the ``easydict`` package the
reference imports (attribute access on a dict, applied recursively, ``AttributeError`` on a missing key).

New optional keys (absent keys keep the reference's behaviour): ``lift_precision`` ("tc16" | "tc" | "fp32"),
``ctx_precision`` ("bf16" | "fp32"), ``cuda_graph`` (bool).  ``default_config()`` returns the keys of
``liftingDWT.json`` as shipped in this package (same keys and values as the reference's file).
"""
import json
import logging
import os
from logging import Formatter
from logging.handlers import RotatingFileHandler

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_JSON = os.path.join(PKG_DIR, "liftingDWT.json")


class EasyDict(dict):
    """Attribute-style dict (recursive), the subset of ``easydict.EasyDict`` the codec relies on."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __delattr__(self, k):
        try:
            del self[k]
        except KeyError:
            raise AttributeError(k) from None

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v


_LOGGING_READY = False


def setup_logging(log_dir):
    """Console (INFO) + ``exp_debug.log`` (DEBUG) + ``exp_error.log`` (WARNING) on the root logger, once per process
    (utils/config.py:24-47)."""
    global _LOGGING_READY
    if _LOGGING_READY:
        return
    root = logging.getLogger()
    root.setLevel(logging.INFO)
    console = logging.StreamHandler()
    console.setLevel(logging.INFO)
    console.setFormatter(Formatter("[%(levelname)s]: %(message)s"))
    root.addHandler(console)
    file_fmt = Formatter("[%(levelname)s] - %(asctime)s - %(name)s - : %(message)s in %(pathname)s:%(lineno)d")
    for name, level in (("exp_debug.log", logging.DEBUG), ("exp_error.log", logging.WARNING)):
        h = RotatingFileHandler(os.path.join(log_dir, name), maxBytes=10 ** 6, backupCount=5)
        h.setLevel(level)
        h.setFormatter(file_fmt)
        root.addHandler(h)
    _LOGGING_READY = True


def get_config_from_json(json_file):
    """(config, config_dict) from a JSON file (utils/config.py:50-66).  A malformed file raises ``ValueError``
    (the reference prints a message and exits the interpreter; a library must not)."""
    with open(json_file, "r") as f:
        try:
            config_dict = json.load(f)
        except ValueError as e:
            raise ValueError(f"INVALID JSON file format in {json_file!r}: {e}") from None
    return EasyDict(config_dict), config_dict


def create_dirs(dirs):
    for d in dirs:
        os.makedirs(d, exist_ok=True)


def process_config(config, root="experiments", quiet=False):
    """Adds ``summary_dir / checkpoint_dir / out_dir / log_dir`` (trailing separator, as the reference's string
    concatenations expect), creates them and sets up logging (utils/config.py:69-103)."""
    try:
        name = config.exp_name
    except (AttributeError, KeyError):
        raise ValueError("the config has no exp_name") from None
    if not quiet:
        print(" *************************************** ")
        print("The experiment name is {}".format(name))
        print(" *************************************** ")
    for key, sub in (("summary_dir", "summaries"), ("checkpoint_dir", "checkpoints"), ("out_dir", "out"), ("log_dir", "logs")):
        config[key] = os.path.join(root, name, sub) + os.sep
    create_dirs([config.summary_dir, config.checkpoint_dir, config.out_dir, config.log_dir])
    setup_logging(config.log_dir)
    logging.getLogger().info("The pipeline of the project will begin now.")
    return config


def default_config(**overrides):
    """The package's ``liftingDWT.json`` (the reference's keys and values) with ``overrides`` applied."""
    config, _ = get_config_from_json(DEFAULT_JSON)
    config.update(overrides)
    return config


# BASELINE.json configs[0..4] as overrides of ``liftingDWT.json`` (shapes are the caller's: see bench.py)
BASELINE_CONFIGS = {
    "cfg1": dict(netType="CDF97", entropy_layer="conditioned2ZTsepSubbands", dwtlevels=4),
    "cfg2": dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=4),
    "cfg3": dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoderBerk",
                 entropy_layer="conditioned2ZTsepSubbands", dwtlevels=4),
    "cfg4": dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoderBerk",
                 entropy_layer="conditioned2ZTsepSubbands", dwtlevels=4, batch_size=8, patch_size=256),
    "cfg5": dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoderBerk",
                 entropy_layer="conditioned2ZTsepSubbands", dwtlevels=5),
}


def baseline_config(name, **overrides):
    cfg = default_config(**BASELINE_CONFIGS[name])
    cfg.update(overrides)
    return cfg
