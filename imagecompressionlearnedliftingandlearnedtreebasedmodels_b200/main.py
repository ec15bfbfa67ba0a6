"""Entry point (reference: main.py:7-33): ``python -m imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.main
<config.json>`` -> ``get_config_from_json`` -> (optional ``multi_agent`` sweep over ``config[multi_param]``) ->
``process_config`` -> ``globals()[config.agent](config).run(); .finalize()``.

The image-folder loaders are out of scope, so a run needs either ``synthetic_data: true`` in the JSON (seeded uniform
RGB batches of ``batch_size x 3 x patch_size x patch_size``, ``synthetic_batches`` per epoch) or a caller that
constructs the agent with its own ``data_loader``."""
import argparse
import os

from .agents import CompressionAgent, LiftingBasedDWTAgent  # noqa: F401  (looked up by name)
from .utils.config import get_config_from_json, process_config


class SyntheticImageLoader:
    """``train_loader / valid_loader / test_loader`` of seeded synthetic RGB batches (no dataset on this path)."""

    def __init__(self, config):
        from .utils.synthetic import synthetic_rgb
        n = int(config.get("synthetic_batches", 4))
        bs, ps = int(config.get("batch_size", 4)), int(config.get("patch_size", 256))
        seed = int(config.get("seed", 1337))
        self.train_loader = [synthetic_rgb(bs, ps, ps, seed + i) for i in range(n)]
        vp = int(config.get("val_patch_size", ps)) or ps
        self.valid_loader = [synthetic_rgb(1, vp, vp, seed + 1000 + i) for i in range(max(1, n // 2))]
        self.test_loader = self.valid_loader


def run_agent(config):
    agent_class = globals()[config.agent]
    loader = SyntheticImageLoader(config) if config.get("synthetic_data", False) else None
    agent = agent_class(config, data_loader=loader)
    agent.run()
    agent.finalize()
    return agent


def main(argv=None):
    ap = argparse.ArgumentParser(description="")
    ap.add_argument("config", metavar="config", help="The Configuration file in json format")
    args = ap.parse_args(argv)
    config, _ = get_config_from_json(args.config)
    if config.get("multi_agent", False):
        for v in config[config.multi_param]:
            config[config.multi_param] = v
            config.exp_name = os.path.join(config.multi_exp_name, "exp_" + str(v))
            run_agent(process_config(config))
    else:
        run_agent(process_config(config))


if __name__ == "__main__":
    main()
