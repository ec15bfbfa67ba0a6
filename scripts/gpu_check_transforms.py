"""First-contact GPU check: lifting level fwd/inv, CDF 9/7 level, pointwise AE vs the oracle,
plus rough timings.  Dev script (gpurun), superseded by tests/ and bench.py."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops  # noqa: E402
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers import lifting_dwt_nets as ldn  # noqa: E402
from oracle import lifting as olift, model as om, subband_ae as oae, thirdparty as tp  # noqa: E402


def ev_time(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    res = {}
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=3)
    torch.manual_seed(1337)
    net = ldn.LiftingBasedNeuralWaveletv4(cfg)
    sd = om.keyed_weights({"m.autoencoder." + k: v for k, v in net.state_dict().items()})
    net.load_state_dict({k[len("m.autoencoder."):]: v for k, v in sd.items()}, strict=True)
    net = net.to(dev).eval()
    for shape in [(2, 1, 48, 80), (1, 1, 128, 192)]:
        torch.manual_seed(3)
        x = torch.rand(*shape) - 0.5
        with torch.no_grad():
            oxe, oxo = None, None
            yl, yh = olift.transform_forward(x, sd, "m.autoencoder.", cfg)
            gl, gh = net.transform(x.to(dev))
            errs = [(gl.cpu() - yl).abs().max().item()] + [(a.cpu() - b).abs().max().item() for a, b in zip(gh, yh)]
            rec = net.inverse_transform(gl, gh)
            orec = olift.transform_inverse(yl, yh, sd, "m.autoencoder.", cfg)
            e_inv = (rec.cpu() - orec).abs().max().item()
            e_pr = (rec.cpu() - x).abs().max().item()
            oe = oae.encode(yh[0], sd, "m.autoencoder.Yh_ae.0.")
            ge, gq = net.Yh_ae[0].encode_and_round(gh[0])
            e_ae = (ge.cpu() - oe).abs().max().item() / oe.abs().max().item()
            flips = (gq.cpu() != torch.round(oe)).sum().item()
            od = oae.decode(torch.round(oe), sd, "m.autoencoder.Yh_ae.0.")
            gd = net.Yh_ae[0].decode(torch.round(oe).to(dev))
            e_dec = (gd.cpu() - od).abs().max().item() / od.abs().max().item()
        res[f"lift{shape}"] = dict(fwd_err=errs, inv_err=e_inv, pr_err=e_pr, ae_rel=e_ae, ae_flips=flips, n=oe.numel(), dec_rel=e_dec)
        print(shape, res[f"lift{shape}"], flush=True)
    # CDF 9/7
    for shape in [(2, 3, 64, 96), (1, 1, 16, 24)]:
        x = torch.rand(*shape) - 0.5
        yl, yh = tp.dwt97_forward(x, 3)
        gl, gh = ops.dwt97_forward(x.to(dev), 3)
        errs = [(gl.cpu() - yl).abs().max().item()] + [(a.cpu() - b).abs().max().item() for a, b in zip(gh, yh)]
        rec = ops.dwt97_inverse(gl, gh)
        e_pr = (rec.cpu() - tp.dwt97_inverse(yl, yh)).abs().max().item()
        res[f"dwt{shape}"] = dict(fwd_err=errs, inv_err=e_pr)
        print(shape, res[f"dwt{shape}"], flush=True)
    # timings at config-2 shape (one colour plane net: B=16 planes of 512x768)
    x = torch.rand(16, 1, 512, 768, device=dev) - 0.5
    cfg4 = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=4)
    net4 = ldn.LiftingBasedNeuralWaveletv4(cfg4).to(dev).eval()
    with torch.no_grad():
        def fwd():
            return net4.transform(x)
        yl, yh = fwd()
        def inv():
            return net4.inverse_transform(yl, yh)
        t_f = ev_time(fwd, 3, 1)
        t_i = ev_time(inv, 3, 1)
        lvl0 = ev_time(lambda: net4.waveletForward[0].level(x), 3, 1)
    px = 16 * 512 * 768
    flop_fwd = 144532.0 * px
    res["time_lift"] = dict(fwd_ms=t_f, inv_ms=t_i, level0_ms=lvl0, tflops_fwd=flop_fwd / (t_f[0] * 1e-3) / 1e12,
                            mpix_s_fwd_inv_plane=px / ((t_f[0] + t_i[0]) * 1e-3) / 1e6)
    print(res["time_lift"], flush=True)
    x3 = torch.rand(16, 3, 512, 768, device=dev) - 0.5
    def dfwd():
        return ops.dwt97_forward(x3, 4)
    gl, gh = dfwd()
    t_df = ev_time(dfwd, 10, 3)
    t_di = ev_time(lambda: ops.dwt97_inverse(gl, gh), 10, 3)
    bytes_dir = 10.625 * 16 * 3 * 512 * 768
    res["time_dwt97"] = dict(fwd_ms=t_df, inv_ms=t_di, fwd_gbs=bytes_dir / (t_df[0] * 1e-3) / 1e9, inv_gbs=bytes_dir / (t_di[0] * 1e-3) / 1e9)
    print(res["time_dwt97"], flush=True)
    ae = net4.Yh_ae[0]
    t_ae = ev_time(lambda: ae.encode_and_round(yh[0]), 5, 2)
    res["time_ae1"] = dict(ms=t_ae, mcoef_s=yh[0].numel() / (t_ae[0] * 1e-3) / 1e6)
    print(res["time_ae1"], flush=True)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/check_transforms.json", "w"), indent=1)


if __name__ == "__main__":
    main()
