"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

For each case the reference's own ``LiftingBasedDWTNetWrapper`` (imported on top of
``oracle/shims``) is built, loaded with synthetic-weights v2
(``oracle.model.keyed_weights`` -- rebuilt anywhere from the key list alone) and run
on a seeded input.  Saved per case: input, reconstruction, per-coefficient
self-information, quantised symbols, pre-quantiser coefficients.  Also saved per
config: the reference's state_dict key/shape list (drop-in check of the product's
module tree) and SHA-256 digests of the seed-1337 default initialisation
(synthetic-weights v1 parity of the product's construction order).
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import model as om, refload  # noqa: E402

CASES = {
    # name: (config overrides, input shape, training)
    "learned_berk_cond2zt_L2": (dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoderBerk",
                                     entropy_layer="conditioned2ZTsepSubbands", dwtlevels=2), (1, 3, 32, 32), False),
    "learned_ae1_ezwt_L2": (dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                                 entropy_layer="onlyEZWT", dwtlevels=2), (1, 3, 32, 48), False),
    "learned_diff_scale_lin_L2": (dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                                       entropy_layer="factorized", dwtlevels=2, block_property="different",
                                       scale=1, linearity_flag=0), (1, 3, 24, 40), False),
    "cdf97_cond2zt_L3": (dict(netType="CDF97", entropy_layer="conditioned2ZTsepSubbands", dwtlevels=3),
                         (1, 3, 32, 48), False),
    "cdf97_ztblock_L2": (dict(netType="CDF97", entropy_layer="DWTConditioned2EntropyLayerZTBlock", dwtlevels=2),
                         (2, 3, 16, 24), False),
    "cdf97_factorized_L2": (dict(netType="CDF97", entropy_layer="factorized", dwtlevels=2), (1, 3, 32, 32), False),
    "cdf97_cond2zt_L2_train": (dict(netType="CDF97", entropy_layer="conditioned2ZTsepSubbands", dwtlevels=2),
                               (1, 3, 32, 32), True),
    # clrch = 3: one net over the three colour channels (9 detail subbands per level, cgp groups = 9)
    "cdf97_cond2zt_L2_clrch3": (dict(netType="CDF97", entropy_layer="conditioned2ZTsepSubbands", dwtlevels=2, clrch=3),
                                (1, 3, 32, 32), False),
}


def cfg_tag(ov):
    return "_".join(f"{k}={ov[k]}" for k in sorted(ov))


def digest_groups(sd):
    groups = {}
    for k, v in sd.items():
        top = ".".join(k.split(".")[:2])
        groups.setdefault(top, hashlib.sha256()).update(v.detach().cpu().contiguous().numpy().tobytes())
    return {k: h.hexdigest() for k, h in groups.items()}


def main(only=None):
    """``only``: regenerate just these cases (their npz + meta entries), keeping the rest of meta.json."""
    m = refload.load()
    torch.Tensor.cuda = lambda self, *a, **k: self   # ZTBlock hard-codes .cuda() (LiftingBasedDWT_net.py:717-718)
    torch.set_num_threads(1)
    meta = {}
    if only:
        with open(os.path.join(HERE, "meta.json")) as f:
            meta = json.load(f)
    for name, (ov, shape, training) in CASES.items():
        if only and name not in only:
            continue
        cfg = refload.default_config(**ov)
        torch.manual_seed(1337)
        net = m.LiftingBasedDWTNetWrapper(cfg)
        v1_digest = digest_groups(net.state_dict())
        keys = [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]
        sd = om.keyed_weights(net.state_dict())
        net.load_state_dict(sd, strict=True)
        net.train(training)
        torch.manual_seed(4242)
        x = om.preprocess(torch.rand(*shape))
        torch.manual_seed(99)     # noise stream for training-mode cases
        # capture pre-quantiser coefficients and symbols by hooking the three plane nets
        arrays = {"x": x.numpy()}
        with torch.no_grad():
            xhat, si_xe, si_xo = net(x)
            arrays["xhat"] = xhat.numpy()
            arrays["si_xe"] = si_xe.numpy()
            for i, s in enumerate(si_xo):
                arrays[f"si_xo_{i}"] = s.numpy()
            if not training:
                planes = [(net.model, x)] if cfg.clrch == 3 else [(sub, x[:, c:c + 1]) for c, sub in
                                                                  enumerate((net.model0, net.model1, net.model2))]
                for c, (sub, xin) in enumerate(planes):
                    out_xe, out_xo = sub.autoencoder.encode(xin)
                    _, _, xe_q, xo_q = sub.entropymodel(out_xe, out_xo)
                    arrays[f"out_xe_{c}"] = out_xe.numpy()
                    arrays[f"xe_q_{c}"] = xe_q.numpy()
                    for i in range(cfg.dwtlevels):
                        arrays[f"out_xo_{c}_{i}"] = out_xo[i].numpy()
                        arrays[f"xo_q_{c}_{i}"] = xo_q[i].numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **arrays)
        bits = float(si_xe.sum() + sum(s.sum() for s in si_xo))
        meta[name] = {"config": ov, "shape": list(shape), "training": training, "keys": keys,
                      "v1_seed1337_sha256": v1_digest, "bpp": bits / (shape[0] * shape[2] * shape[3]),
                      "torch": torch.__version__}
        print(name, "bpp", meta[name]["bpp"], "n keys", len(keys))

    if only:
        with open(os.path.join(HERE, "meta.json"), "w") as f:
            json.dump(meta, f)
        return
    # lifting-only vectors: one level forward / inverse on a ragged plane, reference modules directly
    cfg = refload.default_config(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                                 entropy_layer="factorized", dwtlevels=1)
    torch.manual_seed(1337)
    net = m.LiftingBasedDWTNetWrapper(cfg).eval()
    sd = om.keyed_weights(net.state_dict())
    net.load_state_dict(sd, strict=True)
    torch.manual_seed(5)
    x = torch.rand(2, 1, 24, 40) - 0.5
    with torch.no_grad():
        LL, LH, HL, HH = net.model0.autoencoder.waveletForward[0].one_level_lifting(x)
        rec = net.model0.autoencoder.waveletInverse[0].one_level_lifting(LL, LH, HL, HH)
    np.savez_compressed(os.path.join(HERE, "lifting_one_level.npz"), x=x.numpy(), LL=LL.numpy(), LH=LH.numpy(),
                        HL=HL.numpy(), HH=HH.numpy(), rec=rec.numpy())
    meta["lifting_one_level"] = {"config": dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                                                entropy_layer="factorized", dwtlevels=1), "shape": [2, 1, 24, 40],
                                 "keys": [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]}
    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f)
    print("wrote", HERE)


if __name__ == "__main__":
    main(only=sys.argv[1:] or None)
