"""BASELINE config 4: rate-distortion training step, batch 8 of 256x256 crops per GPU, data-parallel with a
gradient all-reduce (NCCL over NVLink) -- forward on the fused kernels, backward by recomputation.
  python scripts/bench_train.py                      (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py
Prints one JSON line on rank 0 (images/s over all ranks, weak scaling)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle import model as om   # default_cfg / preprocess only (config + input helpers)
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import parallel
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import LiftingBasedDWTNetWrapper
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.losses.rate_dist import TrainRDLoss

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
torch.backends.cudnn.benchmark = os.environ.get("CUDNN_BENCHMARK", "1") == "1"   # the reference agent sets it too (main.py)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
steps, warmup = int(os.environ.get("STEPS", 3)), int(os.environ.get("WARMUP", 2))
ae = os.environ.get("AE", "SubbandAutoEncoder")
cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder=ae, entropy_layer="conditioned2ZTsepSubbands", dwtlevels=4)
torch.manual_seed(1337)
model = LiftingBasedDWTNetWrapper(cfg).to(dev).train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-4)
crit = TrainRDLoss(cfg.lambda_)
g = torch.Generator().manual_seed(1337 + rank)
x = om.preprocess(torch.rand(8, 3, 256, 256, generator=g)).to(dev)

def step():
    opt.zero_grad(set_to_none=True)
    xhat, si_xe, si_xo = model(x)
    out = crit.forward3(x, xhat, si_xe, si_xo)
    loss = out[0] if isinstance(out, (tuple, list)) else out
    loss.backward()
    n = parallel.allreduce_gradients(params)
    opt.step()
    return loss, n

for _ in range(warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss, ncoll = step()
e1.record(); torch.cuda.synchronize()
ms = parallel.max_over_ranks(e0.elapsed_time(e1) / steps, device=dev)
if rank == 0:
    nparam = sum(p.numel() for p in params)
    print(json.dumps({"workload": "config 4: RD training step, batch 8 of 256x256 per GPU, learned lifting L=4 + " + ae + " + cond2ZT",
                      "n_gpus": world, "ms_per_step": ms, "images_per_s": 8 * world / (ms * 1e-3), "loss": float(loss),
                      "grad_allreduce_collectives": ncoll, "params": nparam, "grad_bytes": nparam * 4, "scaling": "weak"}))
if world > 1:
    dist.destroy_process_group()
