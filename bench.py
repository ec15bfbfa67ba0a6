#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native learned-lifting + tree-entropy hot path.

Workload = BASELINE.json configs[2]: 4-level learned lifting DWT (predict/update CNNs) + ``SubbandAutoEncoderBerk`` scaling
network (the reference's ``liftingDWT.json`` default) + ``conditioned2ZTsepSubbands`` inter/intra-subband tree-based
entropy model, full encode+decode (``model.eval(); model(x)``: transform -> quantise + rate estimate -> inverse transform),
bpp + PSNR, batch 64 of 512x768 synthetic RGB images per GPU, fp32.  Metric: megapixels/s (image pixels B*H*W per step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--blocks all|none|a,b,...]

* own arm: one process per GPU (torchrun for N>1), weak scaling -- every rank codes its own batch, no data-path collective;
  barrier + device-event timing, max over ranks.  ``value`` = inputs resident in HBM; ``e2e`` = the same step through
  ``LiftingBasedDWTAgent.validate_batch`` on a pinned HOST RGB batch (H2D of the batch and D2H of the R-D scalars inside the
  timed region).
* ``--impl reference``: the reference's CPU path for the same workload = the oracle port (oracle/, torch-on-CPU restatement
  checked bit-for-bit against the unmodified reference in the dev container; the Python reference itself cannot travel to
  the GPU box), all host threads, each step a bounded sample (one 512x768 image) of the workload.
* bpp-matched: in the ``cpu_baseline`` leg the oracle's output for image 0 of the batch is compared with the GPU's (bpp
  within 0.1 %, symbol flips counted) -- the oracle is the checker there, never the thing measured as the product.
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

B, H, W, LEVELS = 64, 512, 768, 4
METRIC = "megapixels/sec encode+decode (bpp-matched)"
WORKLOAD = ("configs[2]: learned lifting DWT (4 levels) + SubbandAutoEncoderBerk + conditioned2ZTsepSubbands tree-based entropy "
            "model, full encode+decode bpp+PSNR, batch 64 of 512x768 RGB, 3 colour planes")
# SURVEY.md 8(d), per plane pixel: learned lifting 144 532 FLOP each way, Berk scaling network 260 822 each way,
# conditioned2ZT entropy model 430 482 -> full eval forward 1 241 192
FLOP_PER_PLANE_PX = 1241192.0
PKG = "imagecompressionlearnedliftingandlearnedtreebasedmodels_b200"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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
        load = sm[len(sm) // 2:] if sm else []     # under load = the upper half (the sampler also sees the idle edges)
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ helpers
def product():
    import importlib
    return importlib.import_module(PKG)


def cfg_of(name, **kw):
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C
    return C.baseline_config(name, **kw)


def make_agent(cfg, dev, weights="keyed"):
    """LiftingBasedDWTAgent on ``dev`` with seeded construction (1337) + synthetic-weights v2 (non-degenerate symbols)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.agents import LiftingBasedDWTAgent
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    torch.manual_seed(1337)
    agent = LiftingBasedDWTAgent(cfg, device=dev)
    if weights == "keyed":
        S.load_keyed_weights(agent.model)        # in place: the optimizer keeps its parameter objects
    return agent


def ev_time(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ``ncu --set full``
    extract (profiles/r02_ncu_traffic.json: {kernel_key: {"bytes": ..., "source": ...}}); None when no capture exists."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.isfile(p):
        return None, None
    try:
        with open(p) as f:
            d = json.load(f)
        e = d.get(kernel_key)
        return (float(e["bytes"]), e.get("source")) if e else (None, None)
    except Exception:  # noqa: BLE001
        return None, None


# ------------------------------------------------------------------------------------------------ own arm
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
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    _lib.load()
    cfg = cfg_of("cfg3")
    agent = make_agent(cfg, dev)
    x_host = S.synthetic_rgb(B, H, W, 1337 + rank).pin_memory()
    x_dev = x_host.to(dev)

    def step_resident():
        return agent.validate_batch_async(x_dev)

    def step_e2e():
        return agent.validate_batch(x_host)          # H2D of the pinned batch ... D2H of [sse, bits_xe, bits_xo]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        out = None
        for _ in range(warmup):
            out = fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count()
        ev0.record()
        for _ in range(steps):
            out = fn()
        ev1.record()
        timed.launches = ops.launch_count() - n0
        barrier()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, vals = timed(step_resident, args.steps, args.warmup)
    launches = timed.launches
    clocks = sampler.stop() if sampler else None
    ms_e2e, scal = timed(step_e2e, args.steps, max(1, args.warmup))
    rd = agent._rd_scalars(vals.cpu(), x_dev.numel())

    mp_step = B * H * W / 1e6
    ms_step = ms_total / args.steps
    value = world * mp_step / (ms_step * 1e-3)
    e2e_value = world * mp_step / (ms_e2e / args.steps * 1e-3)

    blocks = parse_blocks(args.blocks)
    multi = {}
    # blocks that exercise the N-rank paths run on every rank (collectives inside)
    if "config5" in blocks:
        multi["config5_tiles"] = config5_block(dev, rank, world, dist)
    if "train_step" in blocks:
        multi["train_step"] = train_block(dev, rank, world, dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    plane_px = 3 * B * H * W
    line = {
        "metric": METRIC, "value": value, "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W, "levels": LEVELS,
                   "netType": cfg.netType, "autoencoder": cfg.autoencoder, "entropy_layer": cfg.entropy_layer,
                   "weights": "seed-1337 construction + synthetic-weights v2 (keyed_weights: non-degenerate symbols, "
                              "sigma heads off the likelihood floor); the authors' checkpoints are not available offline",
                   "arithmetic": "lifting conv2/conv3 + scaling network on tcgen05 as 3xTF32 (fp32-level accuracy, they feed the "
                                 "quantiser); context CNNs BF16 operands / FP32 accumulate (they only move bpp)",
                   "l2": "per-step working set (302 MB RGB in, 302 MB YCbCr, multi-GB channels-last intermediates) exceeds the 126 MB L2",
                   "parallelism": f"image-parallel x{world}, no data-path collective",
                   "third_party_parity": "compressai / pytorch_wavelets arithmetic is pinned to the oracle's restatement only "
                                         "(un-vendored upstream, not installable offline): third-party parity unpinned"},
        "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 24,
                "ms_per_step": ms_e2e / args.steps,
                "api": "LiftingBasedDWTAgent.validate_batch(pinned host RGB batch) -> python floats (rd_loss, mse, psnr, rate1, rate2, bpp)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "result": {"bpp": rd["bpp"], "psnr": rd["psnr"], "rate_xe_bpp": rd["rate1"], "rate_xo_bpp": rd["rate2"],
                   "e2e_bpp": scal["bpp"],
                   "note": "synthetic weights: bpp / PSNR are by-products of the measured work, not a quality claim"},
        "step_flops": {"useful_tflops": FLOP_PER_PLANE_PX * plane_px / (ms_step * 1e-3) / 1e12,
                       "note": "SURVEY.md 8(d): 1 241 192 conv FLOP per plane pixel for the whole eval forward (lifting 2 x 144 532, "
                               "scaling network 2 x 260 822, conditioned2ZT 430 482) / step time"},
    }
    line["roofline"] = roofline_block(dev, pk, pk_kind, agent, ms_step)
    line["roofline_lifting"] = roofline_lifting_block(dev, pk, pk_kind)
    if "component_ms" in blocks:
        line["component_ms"] = component_block(dev, agent)
    line.update(multi)
    if world == 1:
        for name, fn in (("lifting_transform", lambda: lifting_block(dev, pk)), ("dwt97", lambda: dwt97_probe(dev, pk)),
                         ("context_cnn", lambda: context_probe(dev, pk)), ("codec_forward", lambda: codec_probe(dev)),
                         ("agent_pointwise", lambda: colour_probe(dev, pk)), ("entropy_coder", lambda: coder_probe(dev)),
                         ("cfg1", lambda: cfg1_block(dev)), ("gpu_torch_baseline", lambda: gpu_torch_baseline(dev))):
            if name in blocks:
                try:
                    line[name] = fn()
                except Exception as e:  # noqa: BLE001 -- a side block must not cost the headline line
                    line[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
                torch.cuda.empty_cache()
        if "cpu_baseline" in blocks:
            del x_dev
            torch.cuda.empty_cache()
            line["cpu_baseline"], line["bpp_match"] = cpu_baseline_and_check(agent, x_host[0:1], budget_s=20.0)
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


ALL_BLOCKS = ["component_ms", "lifting_transform", "dwt97", "context_cnn", "codec_forward", "agent_pointwise", "entropy_coder",
              "cfg1", "config5", "train_step", "gpu_torch_baseline", "cpu_baseline"]


def parse_blocks(spec):
    if spec in (None, "", "all"):
        return set(ALL_BLOCKS)
    if spec == "none":
        return set()
    return set(s.strip() for s in spec.split(",") if s.strip())


# ------------------------------------------------------------------------------------------------ roofline of the dominant kernel
def roofline_block(dev, pk, pk_kind, agent, ms_step):
    """Dominant kernel family of the headline step = the 3xTF32 conv + GDN kernels of SubbandAutoEncoderBerk (~45 % of the step;
    the tensor-core lifting kernel, ~37 %, is reported beside it as ``roofline_lifting``).  One launch of its largest instance
    -- ``ll::igemm_tf32_gdn_pair_kernel``: 3x3 conv 96 -> 192 fused with its GDN on the level-0 subbands of 8 images
    (8 x 256 x 384 pixels) -- timed alone with CUDA events on the launching stream.
    achieved = useful FLOPs of the launch (2 x (96 x 9 + 192) x 192 per pixel: conv + GDN norm; the 3x split and padding are
    NOT counted) / its duration; peak = the dense TF32 tensor-pipe rate measured in this run by ``ll_tf32_peak_probe``."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    nb, h, w = 8, H // 2, W // 2
    C, N = 96, 192
    a = torch.randn(nb, h, w, 2 * C, device=dev)
    wp = ops.pack_tf32_weight(torch.randn(N, C, 3, 3, device=dev) * 0.05)
    gp = ops.pack_tf32_weight((torch.rand(N, N, device=dev) * 0.01 + 0.1 * torch.eye(N, device=dev)).reshape(N, N, 1, 1).contiguous())
    bias, beta = torch.zeros(N, device=dev), torch.ones(N, device=dev)
    ms = ev_time(lambda: ops.igemm_tf32_gdn(a, wp, bias, gp, beta, N), 10)
    flop = 2.0 * (C * 9 + N) * N * nb * h * w
    tf32_peak = ops.tf32_peak_tflops()
    achieved = flop / (ms * 1e-3) / 1e12
    traffic, src = ncu_traffic("igemm_tf32_gdn_pair_kernel conv 96->192 + GDN 8x256x384")
    rf = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
          "traffic": traffic, "traffic_source": src,
          "kernel": "ll::igemm_tf32_gdn_pair_kernel (CTA pairs, 3xTF32 implicit-GEMM 3x3 conv 96->192 + GDN, 8x256x384 px)",
          "ms_per_launch": ms, "issued_tflops": 3 * achieved, "issued_frac": 3 * achieved / tf32_peak,
          "peak_kind": "measured in this run: dense tcgen05 kind::tf32 M128xN256xK8 from shared memory on all SMs (ll_tf32_peak_probe); "
                       f"for context the {pk_kind} cuBLAS bf16 burst peak is {pk['bf16_tflops']} TFLOP/s",
          "algorithmic_bytes": (2 * C + 2 * N) * 4.0 * nb * h * w,
          "note": "achieved = algorithmic FLOPs of ONE launch / its average duration over 10 back-to-back launches (CUDA events on "
                  "the launching stream); issued = x3 for the 3xTF32 split that gives fp32-level accuracy (the path feeds the "
                  "quantiser); algorithmic_bytes = [hi|lo] input read once + [hi|lo] output written once"}
    del a
    return rf


# per 52-column row step the 3xFP16 lifting kernel issues 40 tcgen05.mma (M128 x N64 x K16, kind::f16): 15 each for conv2 / conv3
# (5 vertical taps x 3 terms of the split) and 10 for conv4 (5 taps x 2: the three terms sit in different accumulator rows)
LIFT_ISSUED_FLOP_PER_PX = 40 * 2.0 * 128 * 64 * 16 / 52
LIFT_MAC_PER_PX = 13603          # conv1 400 (SIMT) + conv2 6400 + conv3 6400 + conv4 400 + 3-tap pre-filter


def roofline_lifting_block(dev, pk, pk_kind):
    """Second kernel family of the headline step: ``ll::lift_step_tc_kernel`` (3xFP16 variant, the default) timed alone --
    level-0 row step on a (16,256,768) view: 2 x 13 603 MAC per view pixel, 97 % of them (conv2, conv3, conv4) as a 3xFP16
    split on tcgen05 ``kind::f16`` -- against the measured dense 16-bit tensor peak (cuBLAS bf16, MEASURED_PEAKS.json)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import \
        LiftingBasedNeuralWaveletv4
    cfg = cfg_of("cfg2")
    torch.manual_seed(1337)
    net = LiftingBasedNeuralWaveletv4(cfg).to(dev).eval()
    blobs = net.waveletForward[0]._blobs()
    nb = 16
    vsrc = torch.rand(nb, H // 2, W, device=dev) - 0.5
    vdin = torch.rand(nb, H // 2, W, device=dev) - 0.5
    vout = torch.empty_like(vsrc)
    k_ms = ev_time(lambda: ops.lift_step([(vsrc, vdin, vout)], blobs[0], 1.0, 0.1, False), 20)
    peak = float(pk["bf16_tflops"])
    k_flop = 2.0 * LIFT_MAC_PER_PX * vsrc.numel()
    traffic, src = ncu_traffic("lift_step_tc_kernel level-0 row step (16,256,768)")
    ach = k_flop / (k_ms * 1e-3) / 1e12
    issued = LIFT_ISSUED_FLOP_PER_PX * vsrc.numel() / (k_ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": src,
            "kernel": "ll::lift_step_tc_kernel<3xFP16> (level-0 row step, (16,256,768) view)",
            "ms_per_launch": k_ms, "issued_tflops": issued, "issued_frac": issued / peak,
            "peak_kind": f"{pk_kind} dense 16-bit tensor peak (cuBLAS bf16 burst) -- the kernel's MMAs are kind::f16",
            "algorithmic_bytes": 12.0 * vsrc.numel(),
            "note": "useful conv FLOPs of one launch / its duration; issued = the 40 M128xN64xK16 MMAs per 52-column row step "
                    "(3-term split, weights-as-M: 80 of 128 rows used by conv2 / conv3, 15 by conv4, 52 of 64 columns kept). "
                    "The kernel is bound by the MMA rate at this shape (62 cycles per N = 64 kind::f16 MMA with an MN-major B "
                    "operand, measured) and by the SIMT roles that share the SM, not by the tensor pipe's peak"}


def component_block(dev, agent):
    """Where the headline step's time goes: one colour plane's network on 16 images of 512x768, stage by stage."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    sub = agent.model.model0
    ae = sub.autoencoder
    x = torch.rand(16, 1, H, W, device=dev) - 0.5
    res = {}
    with torch.no_grad():
        yl, yh = ae.transform(x)
        res["lifting_fwd"] = ev_time(lambda: ae.transform(x), 3, 1)
        res["lifting_inv"] = ev_time(lambda: ae.inverse_transform(yl, yh), 3, 1)
        enc = lambda: (ae.Yl_ae.encode(yl), [ae.Yh_ae[i].encode(yh[i]) for i in range(LEVELS)])
        oxe, oxo = enc()
        res["scaling_net_encode"] = ev_time(enc, 3, 1)
        res["scaling_net_decode"] = ev_time(lambda: (ae.Yl_ae.decode(oxe), [ae.Yh_ae[i].decode(oxo[i]) for i in range(LEVELS)]), 3, 1)
        res["entropy_model"] = ev_time(lambda: sub.entropymodel(oxe, oxo), 3, 1)
    res["unit"] = "ms per colour plane at batch 16 of 512x768 (the headline step = 3 planes x 4 such batches + colour / reductions)"
    return res


# ------------------------------------------------------------------------------------------------ configs[1]: transform only
def lifting_block(dev, pk):
    """BASELINE configs[1] (round 1's headline): 4-level learned lifting forward + inverse alone, batch 16 of 512x768, three
    colour planes on three streams; plus the tensor-core lifting kernel timed alone (level-0 row step on a (16,256,768)
    view) against the measured TF32 peak."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import \
        LiftingBasedNeuralWaveletv4
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    cfg = cfg_of("cfg2")
    torch.manual_seed(1337)
    nets = [LiftingBasedNeuralWaveletv4(cfg).to(dev).eval() for _ in range(3)]
    nb = 16
    x = ops.rgb_to_ycbcr_shift(S.synthetic_rgb(nb, H, W, 7).to(dev))
    planes = [x[:, c:c + 1].contiguous() for c in range(3)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]

    def step():
        cur = torch.cuda.current_stream()
        outs = [None] * 3
        for c, net in enumerate(nets):
            st = streams[c]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                yl, yh = net.transform(planes[c])
                outs[c] = net.inverse_transform(yl, yh)
                outs[c].record_stream(cur)
        for st in streams:
            cur.wait_stream(st)
        return outs

    with torch.no_grad():
        ms = ev_time(step, 5, 3)
        rec = torch.cat(step(), dim=1)
        for net in nets:
            net.lift_precision = "fp32"
        ms32 = ev_time(step, 2, 1)
        for net in nets:
            net.lift_precision = "tc16"
        blobs = nets[0].waveletForward[0]._blobs()
        vsrc = torch.rand(nb, H // 2, W, device=dev) - 0.5
        vdin = torch.rand(nb, H // 2, W, device=dev) - 0.5
        vout = torch.empty_like(vsrc)
        k_ms = ev_time(lambda: ops.lift_step([(vsrc, vdin, vout)], blobs[0], 1.0, 0.1, False), 20)
    fma_peak = ops.fma_peak_tflops()
    peak16 = float(pk["bf16_tflops"])
    k_flop = 2.0 * LIFT_MAC_PER_PX * vsrc.numel()
    k_issued = LIFT_ISSUED_FLOP_PER_PX * vsrc.numel()
    flops = 2 * 144532.0 * 3 * nb * H * W
    return {"workload": "configs[1]: learned lifting DWT 4-level forward+inverse, batch 16 of 512x768, 3 colour planes",
            "mp_per_s": nb * H * W / 1e6 / (ms * 1e-3), "ms_per_step": ms,
            "perfect_reconstruction_max_abs_err": float((rec - x).abs().max().item()),
            "hbm_gbs_algorithmic": 21.25 * 3 * nb * H * W / (ms * 1e-3) / 1e9,
            "step_useful_tflops": flops / (ms * 1e-3) / 1e12,
            "lift_step_tc_kernel": {"view": [nb, H // 2, W], "ms": k_ms, "arithmetic": "3xFP16 split on tcgen05 kind::f16 (conv2, conv3, conv4)",
                                    "useful_tflops": k_flop / (k_ms * 1e-3) / 1e12,
                                    "useful_frac_of_16bit_tensor_peak": k_flop / (k_ms * 1e-3) / 1e12 / peak16,
                                    "issued_frac_of_16bit_tensor_peak": k_issued / (k_ms * 1e-3) / 1e12 / peak16,
                                    "tensor_peak_16bit": peak16},
            "fp32_fma_path": {"ms_per_step": ms32, "useful_frac_of_ffma2_peak": flops / (ms32 * 1e-3) / 1e12 / fma_peak,
                              "ffma2_peak_measured": fma_peak}}


def dwt97_probe(dev, pk):
    """HBM-bound fixed-filter kernels (K1): 4-level CDF 9/7 forward and inverse on 16x3 planes of 512x768 (algorithmic
    10.625 B per plane pixel per direction) and the level-0 launch alone at batch 16 / 64."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, ops
    nb = 16
    x = torch.rand(nb, 3, H, W, device=dev) - 0.5
    yl, yh = ops.dwt97_forward(x, LEVELS)
    res = {}
    for name, fn in (("fwd", lambda: ops.dwt97_forward(x, LEVELS)), ("inv", lambda: ops.dwt97_inverse(yl, yh))):
        ms = ev_time(fn, 20)
        gbs = 8.0 * sum(4.0 ** -l for l in range(LEVELS)) * 3 * nb * H * W / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"]}
    res["note"] = "4-level call; 150 MB in+out per level-0 launch (> L2 126 MB); back-to-back launches, CUDA events"
    lib = _lib.load()
    sp = torch.cuda.current_stream().cuda_stream
    for b in (nb, 4 * nb):
        n = 3 * b
        xi = torch.rand(n, H, W, device=dev) - 0.5
        ll = torch.empty(n, H // 2, W // 2, device=dev)
        hh = torch.empty(n, 3, H // 2, W // 2, device=dev)
        xr = torch.empty_like(xi)
        f = lambda: _lib.check(lib.ll_dwt97_fwd_level(xi.data_ptr(), H * W, ll.data_ptr(), H * W // 4, hh.data_ptr(), 3 * H * W // 4, n, H, W, sp))
        g = lambda: _lib.check(lib.ll_dwt97_inv_level(ll.data_ptr(), H * W // 4, hh.data_ptr(), 3 * H * W // 4, xr.data_ptr(), H * W, n, H, W, sp))
        for name, fn in (("fwd", f), ("inv", g)):
            ms = ev_time(fn, 20)
            gbs = 8.0 * n * H * W / (ms * 1e-3) / 1e9
            res[f"level0_{name}_batch{b}"] = {"ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"],
                                              "kernel": f"ll::dwt97_{name}_fast_kernel"}
        res[f"level0_batch{b}_perfect_reconstruction_max_abs_err"] = float((xr - xi).abs().max().item())
        del xi, ll, hh, xr
    return res


def context_probe(dev, pk):
    """Tensor-core side of the entropy models (K3): the 243->243 3x3 plc conv as a BF16 tcgen05 implicit GEMM on 8 planes of
    256x384, and the eval forward of every entropy layer on the subbands of 16 planes of 512x768."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models import LiftingBasedDWT_net as M
    nb, h, w = 8, H // 2, W // 2
    x = torch.zeros(nb, h, w, 256, dtype=torch.bfloat16, device=dev)
    x[..., :243] = torch.randn(nb, h, w, 243, device=dev).to(torch.bfloat16)
    wt = torch.randn(243, 243, 3, 3, device=dev) * 0.02
    bias = torch.randn(243, device=dev)
    wp = ops.pack_igemm_weight(wt, npad=256, kpad=256)
    out = torch.empty(nb, h, w, 256, dtype=torch.bfloat16, device=dev)
    ms = ev_time(lambda: ops.igemm_conv(x, wp, bias, 243, out_nhwc=out), 10)
    useful = 2.0 * nb * h * w * 243 * 243 * 9 / (ms * 1e-3) / 1e12
    issued = 2.0 * nb * h * w * 256 * 256 * 9 / (ms * 1e-3) / 1e12
    res = {"plc_igemm": {"ms": ms, "tflops_useful": useful, "tflops_issued": issued, "peak": pk["bf16_tflops"],
                         "frac_useful": useful / pk["bf16_tflops"], "frac_issued": issued / pk["bf16_tflops"],
                         "bound": "tensor", "dtype": "bf16 operands, fp32 accumulate (TMEM)",
                         "shape": "M=8x256x384 px, N=243 (pad 256), K=9x243 (pad 9x256); in+out 806 MB > L2"}}
    del x, out
    cfg = cfg_of("cfg3")
    nb = 16
    xe = torch.randn(nb, 1, H >> LEVELS, W >> LEVELS, device=dev) * 4
    xo = [torch.randn(nb, 3, H >> (l + 1), W >> (l + 1), device=dev) * 4 for l in range(LEVELS)]
    layers = {}
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils.cuda_graph import GraphedForward
    for name, cls, flop in (("conditioned2ZTsepSubbands", M.DWTConditioned2EntropyLayerZTsepSubbands, 430482),
                            ("onlyEZWT", M.onlyEZWT, 354023), ("ZTBlock", M.DWTConditioned2EntropyLayerZTBlock, 47315),
                            ("factorized", M.DWTFactorizedEntropyLayer, 132)):
        torch.manual_seed(1337)
        layer = cls(cfg).to(dev).eval()
        with torch.no_grad():
            ms = ev_time(lambda: layer(xe, xo), 3)
        layers[name] = {"ms_per_plane_batch16": ms, "mp_per_s_3_planes": nb * H * W / 1e6 / (3 * ms * 1e-3),
                        "flops_per_plane_px": flop, "useful_tflops": flop * nb * H * W / (ms * 1e-3) / 1e12}
        try:
            graphed = GraphedForward(layer, xe, xo)
            layers[name]["ms_per_plane_batch16_cuda_graph"] = ev_time(lambda: graphed(xe, xo), 3)
            del graphed
        except Exception as e:  # noqa: BLE001
            layers[name]["cuda_graph_error"] = f"{type(e).__name__}: {e}"[:300]
        del layer
    res["entropy_layers"] = layers
    return res


def codec_probe(dev):
    """The codec forward with the lighter pointwise scaling network (``autoencoder: SubbandAutoEncoder``), same entropy model,
    batch 16 and 64 of 512x768 through the agent (host RGB in, bpp + PSNR out)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    cfg = cfg_of("cfg3", autoencoder="SubbandAutoEncoder")
    agent = make_agent(cfg, dev)
    res = {}
    for nb in (16, 64):
        rgb = S.synthetic_rgb(nb, H, W, 11).pin_memory()
        out = agent.validate_batch(rgb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            out = agent.validate_batch(rgb)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 2 * 1e3
        res[f"agent_validate_batch{nb}"] = {"ms": ms, "mp_per_s": nb * H * W / 1e6 / (ms * 1e-3), "bpp": out["bpp"], "psnr": out["psnr"],
                                            "h2d_bytes": rgb.numel() * 4, "d2h_bytes": 24}
        del rgb
    res["note"] = "SubbandAutoEncoder (pointwise 1->32->32->32->1) instead of SubbandAutoEncoderBerk; wall clock, host RGB in, python floats out"
    return res


def colour_probe(dev, pk):
    """Agent-side pointwise kernels on batch 64 of 512x768 RGB (302 MB per tensor > L2)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    x = torch.rand(64, 3, H, W, device=dev)
    y = ops.rgb_to_ycbcr_shift(x)
    res = {}
    px = 64 * H * W
    for name, fn, bpp in (("rgb_to_ycbcr_shift", lambda: ops.rgb_to_ycbcr_shift(x), 24.0),
                          ("ycbcr_to_rgb_sse", lambda: ops.ycbcr_to_rgb_sse(y, x, want_xhat=True), 36.0),
                          ("ycbcr_to_rgb_sse_no_xhat", lambda: ops.ycbcr_to_rgb_sse(y, x, want_xhat=False), 24.0)):
        ms = ev_time(fn, 10)
        gbs = bpp * px / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "bytes_per_pixel": bpp}
    return res


def coder_probe(dev):
    """Parallel entropy coder: one colour plane of batch 16 of 512x768 through ``onlyEZWT`` (4 levels):
    subbands -> interleaved rANS streams -> subbands, round trip checked bit for bit."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import onlyEZWT
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    cfg = cfg_of("cfg3", entropy_layer="onlyEZWT")
    torch.manual_seed(1337)
    em = onlyEZWT(cfg)
    em.load_state_dict(S.keyed_weights(em.state_dict()), strict=True)
    em = em.to(dev).eval()
    nb = 16
    torch.manual_seed(3)
    xe = torch.randn(nb, 1, H >> LEVELS, W >> LEVELS, device=dev) * 4
    xo = [torch.randn(nb, 3, H >> (l + 1), W >> (l + 1), device=dev) * (1.5 + l) for l in range(LEVELS)]

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
    px = nb * H * W
    return {"model": "onlyEZWT, one colour plane, batch 16 of 512x768, synthetic-weights v2", "round_trip_exact": exact,
            "compress_ms": ms_enc, "decompress_ms": ms_dec, "mp_per_s_compress": px / 1e6 / (ms_enc * 1e-3),
            "mp_per_s_decompress": px / 1e6 / (ms_dec * 1e-3), "coded_bits_per_coefficient": nbytes * 8 / px,
            "estimated_bits_per_coefficient": est / px, "streams": int(sum(t.counts.numel() for t in streams)),
            "note": "wall clock incl. the context CNNs of both sides and the host-side prefix sums"}


def cfg1_block(dev):
    """BASELINE configs[0] on the GPU (its CPU time is the reference arm's business): CDF 9/7 fixed-filter 4-level DWT +
    conditioned2ZT, rate estimate on ONE 256x256 RGB image -- launch-bound, so eager and replayed as one CUDA graph."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    cfg = cfg_of("cfg1")
    agent = make_agent(cfg, dev)
    rgb = S.synthetic_rgb(1, 256, 256, 5).to(dev)
    res = {}
    t = lambda: agent.validate_batch_async(rgb)
    res["ms_eager"] = ev_time(t, 5)
    out = agent._rd_scalars(t().cpu(), rgb.numel())
    agent.cuda_graph = True
    try:
        res["ms_cuda_graph"] = ev_time(t, 10)
        out_g = agent._rd_scalars(t().cpu(), rgb.numel())
        res["cuda_graph_matches_eager"] = bool(out_g["bpp"] == out["bpp"])
    except Exception as e:  # noqa: BLE001
        res["cuda_graph_error"] = f"{type(e).__name__}: {e}"[:300]
    res.update({"bpp": out["bpp"], "psnr": out["psnr"], "mp_per_s_eager": 256 * 256 / 1e6 / (res["ms_eager"] * 1e-3),
                "workload": "configs[0]: CDF 9/7 4-level + conditioned2ZT, one 256x256 RGB image"})
    if "ms_cuda_graph" in res:
        res["mp_per_s_cuda_graph"] = 256 * 256 / 1e6 / (res["ms_cuda_graph"] * 1e-3)
    return res


# ------------------------------------------------------------------------------------------------ blocks that use all ranks
def config5_block(dev, rank, world, dist):
    """BASELINE configs[4]: one 2048x2048 image cut into 8 independent tiles of 2048x256 (``parallel.tiles_of``), 5-level
    learned lifting + Berk scaling network + conditioned2ZT.  Rank r codes tiles r::world (one tile per rank on 8 GPUs) as
    one CUDA graph per tile batch; bits and squared error are summed over ranks (the only exchange: 2 scalars); time = the
    slowest rank's device time.  On rank 0 the 8-tile batch on one GPU gives the reference bpp the sharded run must equal."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import parallel
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    cfg = cfg_of("cfg5")
    agent = make_agent(cfg, dev)
    img = S.synthetic_rgb(1, 2048, 2048, 4242).to(dev)
    boxes = parallel.tiles_of(2048, 2048, 8)
    mine = parallel.shard_indices(len(boxes), rank, world)
    tiles = torch.cat([img[:, :, boxes[i][0]:boxes[i][1], boxes[i][2]:boxes[i][3]] for i in mine], dim=0).contiguous()
    agent.cuda_graph = len(mine) <= 2          # launch-bound when a rank holds one or two tiles
    fn = lambda: agent.validate_batch_async(tiles)
    for _ in range(3):
        vals = fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        vals = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / n, device=dev)
    sse, bxe, bxo = parallel.sum_over_ranks([float(v) for v in vals.cpu()], device=dev)
    px = 2048 * 2048
    res = {"tile": [boxes[0][1] - boxes[0][0], boxes[0][3] - boxes[0][2]], "tiles": len(boxes), "levels": 5, "ranks": world,
           "tiles_per_rank": len(mine), "ms_per_image": ms, "mp_per_s": px / 1e6 / (ms * 1e-3),
           "bpp_full_image": (bxe + bxo) / px, "mse": sse / (3 * px), "cuda_graph": bool(agent.cuda_graph),
           "measured": f"on {world} real rank(s), max over ranks of the device time (no projection)"}
    if rank == 0 and world > 1:          # the unsharded answer, for the equality check (outside the timed region)
        agent.cuda_graph = False
        all_tiles = torch.cat([img[:, :, b[0]:b[1], b[2]:b[3]] for b in boxes], dim=0).contiguous()
        ref = agent.validate_batch_async(all_tiles).cpu()
        res["bpp_one_gpu_8_tile_batch"] = float(ref[1] + ref[2]) / px
        res["bpp_equal_to_one_gpu"] = bool(abs(res["bpp_one_gpu_8_tile_batch"] - res["bpp_full_image"]) <= 1e-9 * res["bpp_full_image"])
    del agent
    torch.cuda.empty_cache()
    return res


def train_block(dev, rank, world, dist):
    """BASELINE configs[3]: rate-distortion training step, batch 8 of 256x256 crops per GPU, the default model (learned lifting
    + Berk scaling network + conditioned2ZT, 16.7 M parameters = 66.9 MB of fp32 gradients), through
    ``LiftingBasedDWTAgent.train_batch``: forward on the fused kernels, recompute backward, Adam; under torchrun the gradients
    are averaged with bucketed NCCL all-reduces issued from autograd hooks while backward is still running."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import parallel
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import synthetic as S
    cfg = cfg_of("cfg4", learning_rate=1e-5)
    agent = make_agent(cfg, dev)
    x = S.synthetic_rgb(8, 256, 256, 99 + rank).to(dev)
    for _ in range(2):
        agent.train_batch(x)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        out = agent.train_batch(x)
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / n, device=dev)
    nparam = sum(p.numel() for p in agent.model.parameters() if p.requires_grad)
    res = {"ranks": world, "ms_per_step": ms, "images_per_s": 8 * world / (ms * 1e-3), "rd_loss": float(out[0]),
           "params": nparam, "grad_bytes": nparam * 4, "allreduce_collectives_per_step": agent.last_allreduce_collectives,
           "allreduce": "bucketed (16 MB) NCCL all-reduce from post-accumulate-grad hooks, overlapped with backward" if world > 1 else "none (1 rank)",
           "scaling": "weak"}
    # all-reduce cost alone, for the overlap claim: the same buckets reduced back to back with nothing else running
    if world > 1 and agent._buckets is not None:
        import torch.distributed as d
        bk = agent._buckets
        torch.cuda.synchronize()
        d.barrier()
        e0.record()
        for b in bk.buckets:
            d.all_reduce(b["flat"])
        e1.record()
        torch.cuda.synchronize()
        res["allreduce_alone_ms"] = parallel.max_over_ranks(e0.elapsed_time(e1), device=dev)
    del agent
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------ baselines (the oracle is the checker)
def oracle_state(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def gpu_torch_baseline(dev):
    """The reference MATH through stock PyTorch + cuDNN on the same B200 (what the reference would do here: its modules are
    plain torch ops on ``.cuda()`` tensors, agents/base.py:20-28): the oracle port's modules run on cuda:0, cudnn.benchmark on,
    TF32 off (``torch.backends.cudnn.allow_tf32 = False`` -- the reference never enables TF32 for matmul and symbols must be
    bit-exact), on a bounded sample of the headline workload (4 images of 512x768, cfg3).  A baseline like ``cpu_baseline``:
    measured beside the product, never part of it."""
    from oracle import model as om
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        LiftingBasedDWTNetWrapper
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C, synthetic as S
    cfg = cfg_of("cfg3")
    torch.manual_seed(1337)
    model = S.load_keyed_weights(LiftingBasedDWTNetWrapper(cfg))
    sd = {k: v.to(dev) for k, v in model.state_dict().items()}
    ocfg = om.default_cfg(**C.BASELINE_CONFIGS["cfg3"])
    nb = 4
    x = om.preprocess(S.synthetic_rgb(nb, H, W, 1337)).to(dev)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    try:
        with torch.no_grad():
            ms = ev_time(lambda: om.wrapper_forward(x, sd, ocfg), 3, 2)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return {"value": nb * H * W / 1e6 / (ms * 1e-3), "unit": "MP/s", "ms": ms, "kind": "port on cuda (stock torch 2.11 + cuDNN, TF32 off, cudnn.benchmark)",
            "sample": f"{nb} images of 512x768 (configs[2] model), device-resident input, eval forward",
            "note": "the headline value / this = the kernel win over cuDNN eager on the same GPU"}


def cpu_baseline_and_check(agent, rgb1, budget_s=20.0):
    """(cpu_baseline, bpp_match): the oracle port on the host cores on ONE image of the headline workload (timed), and its
    output as the checker of the GPU path on the same image: bpp within 0.1 %, symbols bit-exact (flips counted)."""
    from oracle import model as om
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = oracle_state(agent.model)
    ocfg = om.default_cfg(**C.BASELINE_CONFIGS["cfg3"])
    x = om.preprocess(rgb1)
    with torch.no_grad():
        t0 = time.perf_counter()
        oxhat, osi_xe, osi_xo, outs = om.wrapper_forward(x, sd, ocfg, full=True)      # first run doubles as the warm-up
        first = time.perf_counter() - t0
        n, t0 = 0, time.perf_counter()
        while n < 4 and (time.perf_counter() - t0) + first * (n + 1) / max(n, 1) < budget_s:
            om.wrapper_forward(x, sd, ocfg)
            n += 1
        dt = (time.perf_counter() - t0) / n if n else first
    base = {"value": (H * W / 1e6) / dt, "unit": "MP/s", "cores": cores, "kind": "port",
            "sample": f"{max(n, 1)} x one 512x768 image (3 planes: lifting + Berk scaling network + conditioned2ZT + inverse) of the "
                      f"batch-64 workload, torch {torch.__version__} CPU, {cores} threads; oracle port == unmodified reference "
                      "bit-for-bit in the dev container"}
    # the GPU path on the same image
    dev = agent.device
    agent.model.eval()
    with torch.no_grad():
        xd = x.to(dev)
        xhat, si_xe, si_xo = agent.model(xd)
        flips = nsym = 0
        for c, sub in enumerate(agent.model.planes()):
            oxe, oxo = sub.autoencoder.encode(xd[:, c:c + 1].contiguous())
            _, _, xe_q, xo_q = sub.entropymodel(oxe, oxo)
            for g, o in zip([xe_q] + list(xo_q), [outs[c][3]] + list(outs[c][4])):
                flips += int((g.cpu() != o).sum())
                nsym += o.numel()
        # reconstruction on IDENTICAL symbols (a rounding-boundary flip changes the decoder's input by a whole step)
        rec = torch.cat([sub.autoencoder.decode(outs[c][3].to(dev), [t.to(dev) for t in outs[c][4]])
                         for c, sub in enumerate(agent.model.planes())], dim=1)
    bits = float(si_xe.double().sum() + sum(s.double().sum() for s in si_xo))
    obits = float(osi_xe.double().sum() + sum(s.double().sum() for s in osi_xo))
    rel = (rec.cpu() - oxhat).abs().max().item() / oxhat.abs().max().item()
    rel_e2e = (xhat.cpu() - oxhat).abs().max().item() / oxhat.abs().max().item()
    chk = {"image": "image 0 of rank 0's batch", "bpp_gpu": bits / (H * W), "bpp_oracle": obits / (H * W),
           "bpp_rel_diff": abs(bits - obits) / obits, "bpp_within_0p1_percent": bool(abs(bits - obits) <= 1e-3 * obits),
           "symbols": nsym, "symbol_flips": flips,
           "symbol_flips_note": "fp32 rounding-boundary ties (tests/test_gpu_fullsize.py audits each against the float64 oracle: "
                                "|frac - 0.5| <= 1e-5); the reference itself flips such symbols between thread counts",
           "reconstruction_rel_err_same_symbols": rel, "reconstruction_within_1e-4": bool(rel < 1e-4),
           "reconstruction_rel_err_end_to_end": rel_e2e}
    return base, chk


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the headline path (oracle port: the Python reference cannot travel to the GPU
    box), all host threads, each step = one 512x768 image of the configs[2] workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import model as om
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        LiftingBasedDWTNetWrapper
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C, synthetic as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = cfg_of("cfg3")
    torch.manual_seed(1337)
    model = S.load_keyed_weights(LiftingBasedDWTNetWrapper(cfg))
    sd = oracle_state(model)
    ocfg = om.default_cfg(**C.BASELINE_CONFIGS["cfg3"])
    x = om.preprocess(S.synthetic_rgb(1, H, W, 1337))
    with torch.no_grad():
        for _ in range(args.warmup):
            om.wrapper_forward(x, sd, ocfg)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            xhat, si_xe, si_xo = om.wrapper_forward(x, sd, ocfg)
        dt = (time.perf_counter() - t0) / args.steps
    v = (H * W / 1e6) / dt
    bits = float(si_xe.double().sum() + sum(s.double().sum() for s in si_xo))
    line = {
        "impl": "reference", "metric": METRIC, "value": v,
        "unit": "MP/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "height": H, "width": W, "levels": LEVELS,
                   "netType": cfg.netType, "autoencoder": cfg.autoencoder, "entropy_layer": cfg.entropy_layer,
                   "step": "bounded sample: one 512x768 image (3 planes) of the batch-64 workload per step"},
        "cpu_baseline": {"value": v, "unit": "MP/s", "cores": cores, "kind": "port",
                         "sample": "one 512x768 image (3 planes: lifting + Berk scaling network + conditioned2ZT + inverse) per step; "
                                   "oracle port of the reference's torch-CPU path (the Python reference cannot travel to the GPU box)"},
        "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "result": {"bpp": bits / (H * W)},
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
    ap.add_argument("--blocks", default="all", help="side blocks: all | none | comma list of " + ",".join(ALL_BLOCKS))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
