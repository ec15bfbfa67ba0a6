"""Full codec forward (transform -> quantise + rate -> inverse) at BASELINE config-3 shapes; component breakdown."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import LiftingBasedDWTNetWrapper
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda:0"
def ev(fn, n=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {}
for ae in ("SubbandAutoEncoder", "SubbandAutoEncoderBerk"):
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder=ae, entropy_layer="conditioned2ZTsepSubbands", dwtlevels=4)
    torch.manual_seed(1337)
    model = LiftingBasedDWTNetWrapper(cfg).to(dev).eval()
    x = om.preprocess(torch.rand(B, 3, 512, 768)).to(dev)
    with torch.no_grad():
        full = ev(lambda: model(x))
        m0 = model.model0
        xp = x[:, 0:1].contiguous()
        t_tr = ev(lambda: m0.autoencoder.transform(xp))
        yl, yh = m0.autoencoder.transform(xp)
        t_inv = ev(lambda: m0.autoencoder.inverse_transform(yl, yh))
        t_enc = ev(lambda: m0.autoencoder.encode(xp))
        oxe, oxo = m0.autoencoder.encode(xp)
        t_ent = ev(lambda: m0.entropymodel(oxe, oxo))
        _, _, qe, qo = m0.entropymodel(oxe, oxo)
        t_dec = ev(lambda: m0.autoencoder.decode(qe, qo))
    res[ae] = {"B": B, "full_ms": full, "MP_per_s": B * 512 * 768 / 1e6 / (full * 1e-3),
               "per_plane_ms": {"lifting_fwd": t_tr, "lifting_inv": t_inv, "encode(lift+AE)": t_enc, "entropy": t_ent, "decode(AE+lift)": t_dec,
                                "AE_down": t_enc - t_tr, "AE_up": t_dec - t_inv}}
    print(ae, json.dumps(res[ae]))
json.dump(res, open("gpurun_out/model_timing.json", "w"), indent=1)
