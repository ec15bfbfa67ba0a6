"""TC vs FP32 lifting kernels: one step on several view shapes, then a 4-level transform; timing."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import LiftingBasedNeuralWaveletv4
dev = "cuda:0"
cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=4)
torch.manual_seed(1337)
net = LiftingBasedNeuralWaveletv4(cfg)
sd = om.keyed_weights({"m.autoencoder." + k: v for k, v in net.state_dict().items()})
net.load_state_dict({k[len("m.autoencoder."):]: v for k, v in sd.items()}, strict=True)
net = net.to(dev).eval()
blobs = net.waveletForward[0]._blobs()
torch.manual_seed(3)
for shape in [(1, 16, 64), (2, 40, 100), (1, 8, 52), (3, 70, 53), (2, 128, 384)]:
    src = torch.rand(*shape, device=dev) - 0.5
    din = torch.rand(*shape, device=dev) - 0.5
    outs = {}
    for mode in ("fp32", "tc", "tc16"):
        o = torch.full(shape, 123.0, device=dev)
        ops.lift_step([(src, din, o)], blobs[0], 1.0, 0.1, False, mode)
        torch.cuda.synchronize()
        outs[mode] = o
    d = (outs["tc"] - outs["fp32"]).abs().max().item()
    d16 = (outs["tc16"] - outs["fp32"]).abs().max().item()
    net_scale = (outs["fp32"] - din).abs().max().item()
    print(f"step {shape}: max|tc-fp32| = {d:.3e}  max|tc16-fp32| = {d16:.3e}  (update scale {net_scale:.3f}), nan={bool(torch.isnan(outs['tc16']).any())}", flush=True)
# raw CNN output (sign = 0): isolates the net from the skip/din terms
src = torch.rand(2, 64, 128, device=dev) - 0.5
o = {}
for mode in ("fp32", "tc", "tc16"):
    t = torch.empty_like(src)
    ops.lift_step([(src, src, t)], blobs[1], 0.0, 0.1, False, mode)
    torch.cuda.synchronize(); o[mode] = t
print("net only: max|tc-fp32| =", (o["tc"] - o["fp32"]).abs().max().item(), "max|tc16-fp32| =", (o["tc16"] - o["fp32"]).abs().max().item(), "scale", o["fp32"].abs().max().item())
# one level-0 step alone
src = torch.rand(16, 256, 768, device=dev) - 0.5
din = torch.rand(16, 256, 768, device=dev) - 0.5
out = torch.empty_like(src)
for mode in ("tc", "tc16"):
    for _ in range(3): ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False, mode)
    e1.record(); torch.cuda.synchronize()
    print(f"level-0 step (16,256,768) {mode}: {e0.elapsed_time(e1) / 20:.3f} ms", flush=True)
x = torch.rand(16, 1, 512, 768, device=dev) - 0.5
res = {}
with torch.no_grad():
    for mode in ("fp32", "tc", "tc16"):
        net.lift_precision = mode
        for _ in range(2):
            yl, yh = net.transform(x); rec = net.inverse_transform(yl, yh)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            yl, yh = net.transform(x); rec = net.inverse_transform(yl, yh)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 3, yl, yh, rec)
        print(mode, "ms per plane-batch fwd+inv:", res[mode][0], "PR err", (rec - x).abs().max().item())
rel = lambda a, b: (a - b).abs().max().item() / b.abs().max().item()
for m in ("tc", "tc16"):
    print(m, "yl rel", rel(res[m][1], res["fp32"][1]), "yh rel", [rel(a, b) for a, b in zip(res[m][2], res["fp32"][2])])
