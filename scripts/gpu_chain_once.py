"""Coarsest-level causal chains of conditioned2ZT: tensor path vs fp32 SIMT chain, timing on 16 / 64 planes of 32x48."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models import LiftingBasedDWT_net as M
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C
layer = M.DWTConditioned2EntropyLayerZTsepSubbands(C.default_config(entropy_layer="conditioned2ZTsepSubbands", dwtlevels=4)).to("cuda:0").eval()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    for B in (16, 64):
        for key, seq, inn in (("xo", layer.csc_list[-1], 3), ("xe", layer.csc_xe, 1)):
            q = torch.randint(-6, 7, (B, inn, 32, 48), device="cuda:0").float()
            M.CTX_TC_CHAIN = True
            a = t(lambda: layer._chain_bits_input(key, seq, q))
            M.CTX_TC_CHAIN = False
            b = t(lambda: layer._chain_bits_input(key, seq, q))
            M.CTX_TC_CHAIN = True
            print(f"chain {key} batch {B}: tensor path {a:.3f} ms, fp32 SIMT chain {b:.3f} ms")
