#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native learned-lifting hot path.

Workload (BASELINE.json configs[1]): learned lifting DWT (predict/update CNNs), 4 levels,
forward + inverse, batch 16 of 512x768 synthetic images, three colour planes (three
independent networks, clrch=1), fp32.  Metric: megapixels/s (image pixels B*H*W per step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* own arm: one process per GPU (torchrun for N>1), weak scaling -- every rank transforms its
  own batch, no data-path collective; barrier + device-event timing, max over ranks.
* ``--impl reference``: the reference's CPU path for the same workload = the oracle port
  (oracle/, torch-on-CPU restatement checked bit-for-bit against the unmodified reference in
  the dev container; the reference itself cannot travel to the GPU box), all host threads,
  each step a bounded sample (one image) of the workload.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, H, W, LEVELS = 16, 512, 768, 4
ALG_BYTES_PER_PLANE_PX = 2 * 8.0 * sum(4.0 ** -l for l in range(LEVELS))   # fwd + inv, SURVEY.md 8(d): 21.25
ALG_FLOP_PER_PLANE_PX = 2 * 144532.0                                        # fwd + inv learned lifting, SURVEY.md 8(d)
WORKLOAD = "learned lifting DWT 4-level forward+inverse, batch 16 of 512x768, 3 colour planes (configs[1])"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # under load = the upper half of the samples (the sampler also sees the idle edges)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def build_models(dev):
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import \
        LiftingBasedNeuralWaveletv4
    from oracle import model as om
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=LEVELS)
    torch.manual_seed(1337)
    nets = [LiftingBasedNeuralWaveletv4(cfg) for _ in range(3)]   # random-init weights of the architecture
    return [n.to(dev).eval() for n in nets], cfg


def synthetic_input(seed):
    from oracle import model as om
    g = torch.Generator().manual_seed(seed)
    return om.preprocess(torch.rand(B, 3, H, W, generator=g))


def run_own(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this package has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, ops
    _lib.load()
    nets, cfg = build_models(dev)
    x_host = synthetic_input(1337 + rank).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x_dev = x_host.to(dev)

    x_planes = [x_dev[:, c:c + 1].contiguous() for c in range(3)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]

    def step_resident():
        # the three colour planes are three independent networks: one CUDA stream each, so the small launches
        # of the deep levels of one plane overlap the kernels of another
        cur = torch.cuda.current_stream()
        outs = [None] * 3
        for c, net in enumerate(nets):
            st = streams[c]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                yl, yh = net.transform(x_planes[c])
                outs[c] = net.inverse_transform(yl, yh)
                outs[c].record_stream(cur)
        for st in streams:
            cur.wait_stream(st)
        return outs

    # end to end: the three colour planes are independent (three networks), so each one runs on its own stream --
    # H2D of its plane, transform, inverse, D2H of its reconstruction -- and the copies of one plane overlap the
    # kernels of another.  Plane-major pinned staging buffers (contiguous copies).
    xh_planes = [x_host[:, c:c + 1].contiguous().pin_memory() for c in range(3)]
    oh_planes = [torch.empty_like(t).pin_memory() for t in xh_planes]
    def step_e2e():
        cur = torch.cuda.current_stream()
        for c, net in enumerate(nets):
            st = streams[c]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                xd = xh_planes[c].to(dev, non_blocking=True)
                yl, yh = net.transform(xd)
                rec = net.inverse_transform(yl, yh)
                oh_planes[c].copy_(rec, non_blocking=True)
                for t in (xd, rec):
                    t.record_stream(st)
        for st in streams:
            cur.wait_stream(st)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = ops.launch_count()
            ev0.record()
            for _ in range(steps):
                fn()
            ev1.record()
            timed.launches = ops.launch_count() - n0
            barrier()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_resident, args.steps, args.warmup)
    launches = timed.launches
    clocks = sampler.stop() if sampler else None
    ms_e2e = timed(step_e2e, args.steps, max(1, args.warmup))
    for net in nets:                      # the exact FP32-FMA kernels, for the record (same results to ~1e-6)
        net.lift_precision = "fp32"
    ms_fp32 = timed(step_resident, max(1, args.steps // 2), 1) / max(1, args.steps // 2)
    for net in nets:
        net.lift_precision = "tc"
    with torch.no_grad():
        rec = torch.cat(step_resident(), dim=1)
    pr_err = float((rec - x_dev).abs().max().item())

    mp_step = B * H * W / 1e6
    ms_step = ms_total / args.steps
    value = world * mp_step / (ms_step * 1e-3)
    e2e_value = world * mp_step / (ms_e2e / args.steps * 1e-3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    plane_px = 3 * B * H * W
    alg_bytes = ALG_BYTES_PER_PLANE_PX * plane_px
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    flops = ALG_FLOP_PER_PLANE_PX * plane_px
    fp32_peak = ops.fma_peak_tflops()
    line = {
        "metric": "megapixels/sec encode+decode (learned lifting DWT forward+inverse)", "value": value, "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W, "levels": LEVELS,
                   "weights": "random init (seed 1337), block_property=same, SubbandAutoEncoder not in the timed path",
                   "lift_precision": "tc (conv2/conv3 on tcgen05, 3xTF32 split, fp32-level accuracy)",
                   "l2": "inputs+outputs+scratch per step (3 planes x 3 x 25 MB, read and rewritten 12x per level) exceed the 126 MB L2",
                   "parallelism": f"image-parallel x{world}, no collective; 3 colour planes on 3 CUDA streams per GPU"},
        "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": x_host.numel() * 4,
                "d2h_bytes_per_step": sum(t.numel() for t in oh_planes) * 4, "ms_per_step": ms_e2e / args.steps,
                "api": "LiftingBasedNeuralWaveletv4.transform / .inverse_transform on pinned host tensors, one CUDA stream per colour plane"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        # dominant kernel = ll::lift_step_tc_kernel: 94 % of its MACs (conv2/conv3) run on tcgen05 as 3xTF32
        # (3 TF32 MMAs per product, M = 80 of 128 rows used), so the tensor pipe executes
        # 3 * 128/80 * 0.94 = 4.5x the useful FLOPs; TF32 dense peak = half the measured BF16 peak.
        "roofline": {"bound": "tensor", "achieved": flops / (ms_step * 1e-3) / 1e12, "peak": pk["bf16_tflops"] / 2,
                     "unit": "TFLOP/s", "frac": flops / (ms_step * 1e-3) / 1e12 / (pk["bf16_tflops"] / 2),
                     # DRAM bytes of one launch of the dominant kernel (level-0 step on a (16,256,768) view: 12.6 MB source +
                     # 12.6 MB updated half read, 12.6 MB written) from the ncu --set full capture in
                     # profiles/r01_ncu_lift_tc_final.json: 25.3 MB read + the 12.6 MB result still in L2 at kernel end
                     "traffic": 25.3e6, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                     "peak_kind": pk_kind + " bf16 burst / 2 (TF32)",
                     "kernel": "ll::lift_step_tc_kernel",
                     "issued_tflops": flops * 0.94 * 3 * 128 / 80 / (ms_step * 1e-3) / 1e12,
                     "issued_frac": flops * 0.94 * 3 * 128 / 80 / (ms_step * 1e-3) / 1e12 / (pk["bf16_tflops"] / 2),
                     "note": "useful conv FLOPs (SURVEY.md 8d: 2 x 144532 per plane pixel, fwd+inv) / step time; "
                             "issued = tensor-pipe FLOPs incl. the 3xTF32 split and M padding"},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"],
                         "note": "algorithmic 21.25 B per plane pixel; the learned lifting is compute-bound "
                                 "(13.6 kFLOP per algorithmic byte), HBM is not its roofline"},
        "roofline_fp32": {"bound": "fp32_fma", "achieved": flops / (ms_fp32 * 1e-3) / 1e12, "peak": fp32_peak,
                          "unit": "TFLOP/s", "frac": flops / (ms_fp32 * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                          "ms_per_step": ms_fp32,
                          "peak_kind": "measured in this run: FFMA2 register-only loop, all SMs",
                          "note": "the same step with lift_precision='fp32' (ll::lift_step_kernel, every layer on the FP32 FMA pipe)"},
        "check": {"perfect_reconstruction_max_abs_err": pr_err},
    }
    # the dominant kernel timed alone, per launch (CUDA events on the launching stream, back-to-back launches):
    # one level-0 row step on a (16,256,768) view; 2 x 13 603 MAC per view pixel (3-tap pre-filter + the four 5x5 layers)
    with torch.no_grad():
        blobs = nets[0].waveletForward[0]._blobs()
        vsrc = torch.rand(B, H // 2, W, device=dev) - 0.5
        vdin = torch.rand(B, H // 2, W, device=dev) - 0.5
        vout = torch.empty_like(vsrc)
        for _ in range(3):
            ops.lift_step([(vsrc, vdin, vout)], blobs[0], 1.0, 0.1, False)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(20):
            ops.lift_step([(vsrc, vdin, vout)], blobs[0], 1.0, 0.1, False)
        ev1.record()
        torch.cuda.synchronize()
        k_ms = ev0.elapsed_time(ev1) / 20
        k_flop = 2.0 * 13603 * vsrc.numel()
        tf32_peak = pk["bf16_tflops"] / 2
        line["roofline"]["per_launch"] = {
            "view": [B, H // 2, W], "ms": k_ms, "useful_tflops": k_flop / (k_ms * 1e-3) / 1e12,
            "useful_frac": k_flop / (k_ms * 1e-3) / 1e12 / tf32_peak,
            "issued_tflops": k_flop * 0.94 * 3 * 128 / 80 / (k_ms * 1e-3) / 1e12,
            "issued_frac": k_flop * 0.94 * 3 * 128 / 80 / (k_ms * 1e-3) / 1e12 / tf32_peak,
            "algorithmic_bytes": 12 * vsrc.numel(), "hbm_gbs": 12 * vsrc.numel() / (k_ms * 1e-3) / 1e9,
            "note": "ll::lift_step_tc_kernel alone: 20 back-to-back launches between two CUDA events on the launching stream; "
                    "25.2 MB read (src, din) + 12.6 MB written per launch; frac against the TF32 dense peak (measured BF16 burst / 2)"}
        del vsrc, vdin, vout
        # headline roofline figures = the per-launch ones (algorithmic FLOPs of one launch / its measured duration);
        # the whole-step quotient (all levels, three overlapped streams) is kept beside them
        rf, pl = line["roofline"], line["roofline"]["per_launch"]
        rf["step_achieved"], rf["step_frac"] = rf["achieved"], rf["frac"]
        rf["step_issued_tflops"], rf["step_issued_frac"] = rf["issued_tflops"], rf["issued_frac"]
        rf["achieved"], rf["frac"] = pl["useful_tflops"], pl["useful_frac"]
        rf["issued_tflops"], rf["issued_frac"] = pl["issued_tflops"], pl["issued_frac"]
        rf["note"] = ("achieved = useful conv FLOPs of one ll::lift_step_tc_kernel launch (2 x 13 603 MAC per view pixel, level-0 "
                      "row step on a (16,256,768) view) / its average duration over 20 back-to-back launches (CUDA events); "
                      "issued = tensor-pipe FLOPs incl. the 3xTF32 split and M padding; step_* = the same quotients over the "
                      "whole timed step (SURVEY.md 8d: 2 x 144532 FLOP per plane pixel, fwd+inv, all levels)")
    line["dwt97"] = dwt97_probe(dev, pk)
    line["context_cnn"] = context_probe(dev, pk)
    line["codec_forward"] = codec_probe(dev)
    line["agent_pointwise"] = colour_probe(dev, pk)
    line["entropy_coder"] = coder_probe(dev)
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(budget_s=12.0)
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def dwt97_probe(dev, pk):
    """HBM-bound fixed-filter kernels (K1) on the same image shape: 4-level CDF 9/7 forward and
    inverse on 16x3 planes of 512x768; algorithmic 10.625 B per plane pixel per direction."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    x = torch.rand(B, 3, H, W, device=dev) - 0.5
    yl, yh = ops.dwt97_forward(x, LEVELS)
    res = {}
    for name, fn in (("fwd", lambda: ops.dwt97_forward(x, LEVELS)), ("inv", lambda: ops.dwt97_inverse(yl, yh))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n
        gbs = 8.0 * sum(4.0 ** -l for l in range(LEVELS)) * 3 * B * H * W / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"]}
    res["note"] = "4-level call; 150 MB in+out per level-0 launch (> L2 126 MB); back-to-back launches, CUDA events"
    # the dominant launch alone (level 0 = 75 % of the bytes), at this batch and at the config-3 batch (64 images):
    # 8 B per input sample, straight through the C ABI on preallocated buffers
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib
    lib = _lib.load()
    sp = torch.cuda.current_stream().cuda_stream
    for nb in (B, 4 * B):
        n = 3 * nb
        xi = torch.rand(n, H, W, device=dev) - 0.5
        ll = torch.empty(n, H // 2, W // 2, device=dev)
        hh = torch.empty(n, 3, H // 2, W // 2, device=dev)
        xr = torch.empty_like(xi)
        f = lambda: _lib.check(lib.ll_dwt97_fwd_level(xi.data_ptr(), H * W, ll.data_ptr(), H * W // 4, hh.data_ptr(), 3 * H * W // 4, n, H, W, sp))
        g = lambda: _lib.check(lib.ll_dwt97_inv_level(ll.data_ptr(), H * W // 4, hh.data_ptr(), 3 * H * W // 4, xr.data_ptr(), H * W, n, H, W, sp))
        for name, fn in (("fwd", f), ("inv", g)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(20):
                fn()
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / 20
            gbs = 8.0 * n * H * W / (ms * 1e-3) / 1e9
            res[f"level0_{name}_batch{nb}"] = {"ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"],
                                               "kernel": f"ll::dwt97_{name}_fast_kernel"}
        res[f"level0_batch{nb}_perfect_reconstruction_max_abs_err"] = float((xr - xi).abs().max().item())
        del xi, ll, hh, xr
    return res


def context_probe(dev, pk):
    """Tensor-core side of the path (K3): the 243->243 3x3 plc conv as a tcgen05 implicit GEMM on
    8 planes of 256x384 (level-0 subband of a 512x768 image), and the whole conditioned2ZT entropy
    model (eval forward, bf16 context path) on the subbands of 16 planes of 512x768."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        DWTConditioned2EntropyLayerZTsepSubbands
    from oracle import model as om

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / n

    nb, h, w = 8, H // 2, W // 2
    x = torch.zeros(nb, h, w, 256, dtype=torch.bfloat16, device=dev)
    x[..., :243] = torch.randn(nb, h, w, 243, device=dev).to(torch.bfloat16)
    wt = torch.randn(243, 243, 3, 3, device=dev) * 0.02
    bias = torch.randn(243, device=dev)
    wp = ops.pack_igemm_weight(wt, npad=256, kpad=256)
    out = torch.empty(nb, h, w, 256, dtype=torch.bfloat16, device=dev)
    ms = timed(lambda: ops.igemm_conv(x, wp, bias, 243, out_nhwc=out), 10)
    useful = 2.0 * nb * h * w * 243 * 243 * 9 / (ms * 1e-3) / 1e12
    issued = 2.0 * nb * h * w * 256 * 256 * 9 / (ms * 1e-3) / 1e12
    res = {"plc_igemm": {"ms": ms, "tflops_useful": useful, "tflops_issued": issued, "peak": pk["bf16_tflops"],
                         "frac_useful": useful / pk["bf16_tflops"], "frac_issued": issued / pk["bf16_tflops"],
                         "bound": "tensor", "dtype": "bf16 operands, fp32 accumulate (TMEM)",
                         "shape": "M=8x256x384 px, N=243 (pad 256), K=9x243 (pad 9x256); in+out 806 MB > L2"}}
    del x, out
    cfg = om.default_cfg(dwtlevels=LEVELS)
    torch.manual_seed(1337)
    em = DWTConditioned2EntropyLayerZTsepSubbands(cfg).to(dev).eval()
    xe = torch.randn(B, 1, H >> LEVELS, W >> LEVELS, device=dev) * 4
    xo = [torch.randn(B, 3, H >> (l + 1), W >> (l + 1), device=dev) * 4 for l in range(LEVELS)]
    with torch.no_grad():
        ms = timed(lambda: em(xe, xo), 3)
    res["cond2zt_entropy_model"] = {"ms_per_plane_batch16": ms, "mp_per_s_3_planes": B * H * W / 1e6 / (3 * ms * 1e-3),
                                    "flops_per_plane_px": 430482, "note": "quantise + context CNNs + Gaussian rate + bit sums"}
    try:
        from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils.cuda_graph import GraphedForward
        graphed = GraphedForward(em, xe, xo)
        res["cond2zt_entropy_model"]["ms_per_plane_batch16_cuda_graph"] = timed(lambda: graphed(xe, xo), 3)
        del graphed
    except Exception as e:  # noqa: BLE001
        res["cond2zt_entropy_model"]["cuda_graph_error"] = f"{type(e).__name__}: {e}"[:300]
    del em
    # the other three parallelisable entropy layers on the same subbands, eval forward of one plane
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import (
        DWTConditioned2EntropyLayerZTBlock, DWTFactorizedEntropyLayer, onlyEZWT)
    layers = {}
    for name, cls, flop in (("onlyEZWT", onlyEZWT, 354023), ("ZTBlock", DWTConditioned2EntropyLayerZTBlock, 47315),
                            ("factorized", DWTFactorizedEntropyLayer, 132)):
        torch.manual_seed(1337)
        layer = cls(cfg).to(dev).eval()
        with torch.no_grad():
            ms = timed(lambda: layer(xe, xo), 3)
            out = layer(xe, xo)
        bits = float(out[0].double().sum() + sum(s.double().sum() for s in out[1]))
        layers[name] = {"ms_per_plane_batch16": ms, "mp_per_s_3_planes": B * H * W / 1e6 / (3 * ms * 1e-3),
                        "flops_per_plane_px": flop, "bits_per_coefficient": bits / (B * H * W)}
        try:   # the same call replayed as one CUDA graph (ZTBlock is several hundred small launches)
            from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils.cuda_graph import GraphedForward
            graphed = GraphedForward(layer, xe, xo)
            layers[name]["ms_per_plane_batch16_cuda_graph"] = timed(lambda: graphed(xe, xo), 3)
            gout = graphed(xe, xo)
            layers[name]["cuda_graph_matches_eager"] = bool(torch.equal(gout[0], out[0]) and
                                                            all(torch.equal(a, b) for a, b in zip(gout[1], out[1])))
            del graphed, gout
        except Exception as e:  # noqa: BLE001
            layers[name]["cuda_graph_error"] = f"{type(e).__name__}: {e}"[:300]
        del layer
    res["other_entropy_layers"] = layers
    return res


def codec_probe(dev):
    """Whole codec forward (configs[2] shape at batch 16): transform -> quantise + rate estimate -> inverse transform,
    three colour planes, learned lifting L=4 + conditioned2ZT, with both scaling networks."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        LiftingBasedDWTNetWrapper
    from oracle import model as om
    res = {}
    x = synthetic_input(7).to(dev)
    for ae in ("SubbandAutoEncoder", "SubbandAutoEncoderBerk"):
        cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder=ae, entropy_layer="conditioned2ZTsepSubbands",
                             dwtlevels=LEVELS)
        torch.manual_seed(1337)
        model = LiftingBasedDWTNetWrapper(cfg).to(dev).eval()
        torch.cuda.empty_cache()
        with torch.no_grad():
            for _ in range(2):                     # weight packing, allocator growth
                model(x)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(3):
                xhat, si_xe, si_xo = model(x)
            ev1.record()
            torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 3
        bits = float(si_xe.double().sum() + sum(s.double().sum() for s in si_xo))
        res[ae] = {"ms_per_batch16": ms, "mp_per_s": B * H * W / 1e6 / (ms * 1e-3), "bpp": bits / (B * H * W)}
        del model
    # configs[2] end to end through the agent mirror: pinned host RGB batch -> H2D -> RGB->YCbCr -> codec forward ->
    # YCbCr->RGB, clamp, squared error -> one D2H of the three scalars (bpp, PSNR)
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.agents import LiftingBasedDWTAgent
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                         entropy_layer="conditioned2ZTsepSubbands", dwtlevels=LEVELS)
    torch.manual_seed(1337)
    agent = LiftingBasedDWTAgent(cfg, device=dev)
    rgb = torch.rand(B, 3, H, W).pin_memory()
    for _ in range(2):
        out = agent.validate_batch(rgb)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = agent.validate_batch(rgb)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    res["agent_validate_batch"] = {"ms_per_batch16": ms, "mp_per_s": B * H * W / 1e6 / (ms * 1e-3), "bpp": out["bpp"], "psnr": out["psnr"],
                                   "h2d_bytes": rgb.numel() * 4, "d2h_bytes": 24,
                                   "note": "wall clock around LiftingBasedDWTAgent.validate_batch (host RGB in, python floats out), SubbandAutoEncoder"}
    # the same call at configs[2]'s own batch: 64 images of 512x768 (302 MB of host RGB per call)
    rgb64 = torch.rand(64, 3, H, W).pin_memory()
    out = agent.validate_batch(rgb64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        out = agent.validate_batch(rgb64)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 2 * 1e3
    res["agent_validate_batch64"] = {"ms_per_batch64": ms, "mp_per_s": 64 * H * W / 1e6 / (ms * 1e-3), "bpp": out["bpp"],
                                     "psnr": out["psnr"], "h2d_bytes": rgb64.numel() * 4, "d2h_bytes": 24,
                                     "note": "BASELINE configs[2] shape (batch 64 of 512x768), wall clock, host RGB in, bpp + PSNR out"}
    del agent, rgb64
    torch.cuda.empty_cache()
    # configs[4]: one 2048x2048 image cut into 8 independent tiles of 2048x256 (parallel.tiles_of), 5-level learned
    # lifting + conditioned2ZT.  One rank's share on 8 GPUs = one tile; on one GPU the 8 tiles run as a batch of 8.
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.parallel import tiles_of
    cfg5 = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                          entropy_layer="conditioned2ZTsepSubbands", dwtlevels=5)
    torch.manual_seed(1337)
    model = LiftingBasedDWTNetWrapper(cfg5).to(dev).eval()
    img = om.preprocess(torch.rand(1, 3, 2048, 2048)).to(dev)
    boxes = tiles_of(2048, 2048, 8)
    tiles = torch.cat([img[:, :, y0:y1, x0:x1] for (y0, y1, x0, x1) in boxes], dim=0).contiguous()
    c5 = {"tile": [boxes[0][1] - boxes[0][0], boxes[0][3] - boxes[0][2]], "tiles": len(boxes), "levels": 5}
    with torch.no_grad():
        for key, xin in (("ms_one_tile", tiles[:1]), ("ms_eight_tiles_one_gpu", tiles)):
            for _ in range(2):
                model(xin)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(3):
                xhat, si_xe, si_xo = model(xin)
            ev1.record()
            torch.cuda.synchronize()
            c5[key] = ev0.elapsed_time(ev1) / 3
    bits = float(si_xe.double().sum() + sum(s.double().sum() for s in si_xo))
    c5["bpp_full_image"] = bits / (2048 * 2048)
    # one tile is launch-bound (about a thousand small launches): the same call replayed as one CUDA graph
    try:
        from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils.cuda_graph import GraphedForward
        one = tiles[:1].contiguous()
        graphed = GraphedForward(model, one)
        for _ in range(2):
            graphed(one)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            gx, gsi_xe, gsi_xo = graphed(one)
        ev1.record()
        torch.cuda.synchronize()
        c5["ms_one_tile_cuda_graph"] = ev0.elapsed_time(ev1) / 5
        with torch.no_grad():
            ex, esi_xe, esi_xo = model(one)
        c5["cuda_graph_matches_eager"] = bool(torch.equal(gx, ex) and torch.equal(gsi_xe, esi_xe)
                                              and all(torch.equal(a, b) for a, b in zip(gsi_xo, esi_xo)))
        c5["mp_per_s_8_gpus_one_tile_each_cuda_graph"] = 2048 * 2048 / 1e6 / (c5["ms_one_tile_cuda_graph"] * 1e-3)
        del graphed
    except Exception as e:  # noqa: BLE001 -- a capture failure is reported, the eager numbers above stand
        c5["cuda_graph_error"] = f"{type(e).__name__}: {e}"[:300]
    c5["mp_per_s_one_gpu"] = 2048 * 2048 / 1e6 / (c5["ms_eight_tiles_one_gpu"] * 1e-3)
    c5["mp_per_s_8_gpus_one_tile_each"] = 2048 * 2048 / 1e6 / (c5["ms_one_tile"] * 1e-3)
    c5["note"] = ("8-GPU figure = image pixels / time of one rank's tile measured on this GPU (tiles are independent, "
                  "no exchange; bpp = sum of the tiles' bits / full-image pixels)")
    res["config5_tiles"] = c5
    del model
    res["note"] = "random-init weights: bpp is a by-product, not a quality claim"
    return res


def colour_probe(dev, pk):
    """Agent-side pointwise kernels on batch 64 of 512x768 RGB (302 MB per tensor > L2): RGB->YCbCr with
    Y-0.5 (24 B per pixel) and Y+0.5 / YCbCr->RGB / -0.5 / clamp / squared error (36 B per pixel with the reconstruction
    written, 24 B without)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    x = torch.rand(4 * B, 3, H, W, device=dev)
    y = ops.rgb_to_ycbcr_shift(x)
    res = {}
    px = 4 * B * H * W
    for name, fn, bpp in (("rgb_to_ycbcr_shift", lambda: ops.rgb_to_ycbcr_shift(x), 24.0),
                          ("ycbcr_to_rgb_sse", lambda: ops.ycbcr_to_rgb_sse(y, x, want_xhat=True), 36.0),
                          ("ycbcr_to_rgb_sse_no_xhat", lambda: ops.ycbcr_to_rgb_sse(y, x, want_xhat=False), 24.0)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        gbs = bpp * px / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "bytes_per_pixel": bpp}
    return res


def coder_probe(dev):
    """Parallel entropy coder: one colour plane of batch 16 of 512x768 through ``onlyEZWT``
    (4 levels): subbands -> interleaved rANS streams -> subbands, round trip checked bit for bit."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import onlyEZWT
    from oracle import model as om
    cfg = om.default_cfg(entropy_layer="onlyEZWT", dwtlevels=LEVELS)
    torch.manual_seed(1337)
    em = onlyEZWT(cfg).to(dev).eval()
    torch.manual_seed(3)
    xe = torch.randn(B, 1, H >> LEVELS, W >> LEVELS, device=dev) * 4
    xo = [torch.randn(B, 3, H >> (l + 1), W >> (l + 1), device=dev) * (1.5 + l) for l in range(LEVELS)]

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3, out

    with torch.no_grad():
        si_xe, sis, _, _ = em(xe, xo)
        est = float(si_xe.double().sum() + sum(t.double().sum() for t in sis))
        ms_enc, (streams, xe_q, qs) = timed(lambda: em.compress(xe, xo))
        ms_dec, (dxe, dqs) = timed(lambda: em.decompress(streams))
    exact = bool(torch.equal(dxe, xe_q) and all(torch.equal(a, b) for a, b in zip(dqs, qs)))
    nbytes = sum(t.nbytes() for t in streams)
    px = B * H * W
    return {"model": "onlyEZWT, one colour plane, batch 16 of 512x768, random-init weights", "round_trip_exact": exact,
            "compress_ms": ms_enc, "decompress_ms": ms_dec, "mp_per_s_compress": px / 1e6 / (ms_enc * 1e-3),
            "mp_per_s_decompress": px / 1e6 / (ms_dec * 1e-3), "coded_bits_per_coefficient": nbytes * 8 / px,
            "estimated_bits_per_coefficient": est / px, "streams": int(sum(t.counts.numel() for t in streams)),
            "note": "wall clock incl. the context CNNs of both sides and the host-side prefix sums; estimate counts up to 30 bits "
                    "for symbols the coder's 16-bit probabilities cap at ~17"}


def cpu_port_step(x_img, sds, cfg):
    """One image (3 planes) through the oracle port: 4-level forward + inverse."""
    from oracle import lifting as olift
    with torch.no_grad():
        for c in range(3):
            yl, yh = olift.transform_forward(x_img[:, c:c + 1], sds[c], "m.autoencoder.", cfg)
            olift.transform_inverse(yl, yh, sds[c], "m.autoencoder.", cfg)


def cpu_models():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import \
        LiftingBasedNeuralWaveletv4
    from oracle import model as om
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=LEVELS)
    torch.manual_seed(1337)
    sds = []
    for _ in range(3):
        n = LiftingBasedNeuralWaveletv4(cfg)
        sds.append({"m.autoencoder." + k: v.detach() for k, v in n.state_dict().items()})
    return sds, cfg


def cpu_baseline(budget_s=12.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sds, cfg = cpu_models()
    x = synthetic_input(1337)[0:1]
    cpu_port_step(x, sds, cfg)     # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        cpu_port_step(x, sds, cfg)
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 8:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": (H * W / 1e6) / dt, "unit": "MP/s", "cores": cores, "kind": "port",
            "sample": f"{n} x one 512x768 image (3 planes, 4-level forward+inverse) of the batch-16 workload, "
                      f"torch {torch.__version__} CPU, {cores} threads; oracle port == unmodified reference bit-for-bit in the dev container"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sds, cfg = cpu_models()
    x = synthetic_input(1337)[0:1]
    for _ in range(args.warmup):
        cpu_port_step(x, sds, cfg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_step(x, sds, cfg)
    dt = (time.perf_counter() - t0) / args.steps
    v = (H * W / 1e6) / dt
    line = {
        "impl": "reference", "metric": "megapixels/sec encode+decode (learned lifting DWT forward+inverse)", "value": v,
        "unit": "MP/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W, "levels": LEVELS,
                   "step": "bounded sample: one 512x768 image (3 planes) per step"},
        "cpu_baseline": {"value": v, "unit": "MP/s", "cores": cores, "kind": "port",
                         "sample": "one 512x768 image (3 planes, 4-level forward+inverse) per step; oracle port of the "
                                   "reference's torch-CPU path (the Python reference cannot travel to the GPU box)"},
        "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_JSON_OUT = None


def emit(line):
    """The one JSON line, on the process's real stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's "NCCL version ..." banner when
    # NCCL_DEBUG is set in the environment) is sent to stderr by pointing fd 1 at fd 2 for the rest of the run
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
