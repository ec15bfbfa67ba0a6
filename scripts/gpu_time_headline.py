"""Headline step (configs[2], batch 64) under a few host-side knobs: scaling-network / context chunk sizes, plane streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers import lifting_dwt_nets as L
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models import LiftingBasedDWT_net as M
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
dev = torch.device("cuda", 0)
agent = bench.make_agent(bench.cfg_of("cfg3"), dev)
x = S.synthetic_rgb(64, 512, 768, 1337).to(dev)
def run(tag):
    ms = bench.ev_time(lambda: agent.validate_batch_async(x), 3, 2)
    v = agent._rd_scalars(agent.validate_batch_async(x).cpu(), x.numel())
    print(f"{tag:50s} {ms:8.1f} ms  {64 * 512 * 768 / 1e6 / (ms * 1e-3):6.2f} MP/s  bpp {v['bpp']:.6f}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    torch.cuda.reset_peak_memory_stats()
run("default")
M.CTX_GEMM_HEAD = False
run("context head / csc on the fp32 SIMT kernels")
M.CTX_FUSED_TAIL = False
run("... and cgp tail unfused (the path before)")
M.CTX_GEMM_HEAD = M.CTX_FUSED_TAIL = True
run("default again")
for ae in (16, 32, 64):
    L.SubbandAutoEncoderBerk.AE_BATCH_CHUNK = ae
    run(f"AE_BATCH_CHUNK={ae}")
L.SubbandAutoEncoderBerk.AE_BATCH_CHUNK = 32
for cc in (16, 32):
    M.CTX_BATCH_CHUNK = cc
    run(f"AE 32, CTX_BATCH_CHUNK={cc}")
M.CTX_BATCH_CHUNK = 16
agent.model.plane_streams = True
run("AE 32, CTX 16, plane streams on")
