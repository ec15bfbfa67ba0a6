"""CPU tests: the C-ABI library loads and exports every symbol ``include/ll_api.h`` declares
(no compute calls without a GPU), the ctypes table mirrors the header, the product refuses CPU
tensors, and the host emulation of the kernel bodies (tests/emul -- test tool, never a product
path) agrees with the oracle, which checks tile / ring / halo / wrap indexing in the GPU-less
container."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from oracle import lifting as olift, model as om, thirdparty as tp

from common import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ll_api.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ll_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in ll_api.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms)      # the ctypes table covers the header exactly
    assert _lib.load().ll_version() >= 100        # host-only call
    # process-wide mode switches and debug hooks are not part of the product ABI
    for gone in ("ll_lift_set_mode", "ll_lift_get_mode", "ll_lift_set_debug_buffer", "ll_tc_tf32_probe", "ll_fma_peak_probe"):
        assert not hasattr(lib, gone), gone


def test_probe_library_exports_its_header():
    """include/ll_probe.h (measurement / unit probes) lives in its own library."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(build.PROBE_LIB_PATH)
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ll_probe.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(ll_[a-z0-9_]+)\s*\(", src)))
    assert syms == sorted(k for k in _lib.PROBE_SIGNATURES if k != "ll_last_error")
    for s_ in syms:
        assert hasattr(lib, s_), s_


def test_library_is_sm100a_native():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    with pytest.raises(RuntimeError):
        ops.dwt97_forward(torch.zeros(1, 1, 16, 16), 1)
    with pytest.raises(RuntimeError):
        ops.quantize(torch.zeros(4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "imagecompressionlearnedliftingandlearnedtreebasedmodels_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


# ---------------------------------------------------------------------------- host emulation
@pytest.fixture(scope="module")
def emul():
    so = os.path.join(ROOT, "tests", "emul", "libll_emul.so")
    src = os.path.join(ROOT, "tests", "emul", "lift_emul.cpp")
    if not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.ll_emul_lift_level_scratch_floats.restype = ctypes.c_size_t
    return lib


F32P = ctypes.POINTER(ctypes.c_float)
I64 = ctypes.c_int64


def P(t):
    return ctypes.cast(t.data_ptr(), F32P)


def lifting_weights():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers import lifting_dwt_nets as ldn
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=1)
    torch.manual_seed(1337)
    net = ldn.LiftingBasedNeuralWaveletv4(cfg)
    sd = om.keyed_weights({"m.autoencoder." + k: v for k, v in net.state_dict().items()})
    return sd


@pytest.mark.parametrize("B,h,w,scale,linear,ncta", [(2, 24, 40, 0, 0, 3), (1, 20, 116, 1, 1, 5), (1, 36, 128, 0, 0, 148)])
def test_emulated_lifting_level_matches_oracle(emul, B, h, w, scale, linear, ncta):
    sd = lifting_weights()
    pfx = "m.autoencoder.waveletForward.0."
    cfg = om.default_cfg(scale=scale, linearity_flag=0 if linear else 1)
    steps = [("convBlock.0.weight", "P.0."), ("convBlock.1.weight", "U.0."), ("convBlock.2.weight", "P.1."),
             ("convBlock.3.weight", "U.1.")]
    blobs = []
    for pre, blk in steps:
        b = torch.zeros(13656)
        args = [sd[pfx + pre].contiguous()] + [sd[pfx + blk + f"conv{i}.{wb}"].contiguous() for i in (1, 2, 3, 4)
                                               for wb in ("weight", "bias")]
        emul.ll_emul_pack_lift_step(*[P(a) for a in args], P(b))
        blobs.append(b)
    arr = (F32P * 4)(*[P(b) for b in blobs])
    torch.manual_seed(B * h + w)
    x = torch.rand(B, 1, h, w) - 0.5
    with torch.no_grad():
        LL, LH, HL, HH = olift.one_level_forward(x, sd, pfx, cfg)
        rec = olift.one_level_inverse(LL, LH, HL, HH, sd, pfx, cfg)
    ref = torch.cat([LH, HL, HH], 1).contiguous()
    nan = float("nan")
    ll = torch.full((B, 1, h // 2, w // 2), nan)
    yh = torch.full((B, 3, h // 2, w // 2), nan)
    scratch = torch.full((emul.ll_emul_lift_level_scratch_floats(B, h, w),), nan)
    nh, nl = sd[pfx + "nh"].contiguous(), sd[pfx + "nl"].contiguous()
    rc = emul.ll_emul_lift_level_fwd(P(x), I64(h * w), P(ll), I64(h * w // 4), P(yh), I64(3 * h * w // 4), P(scratch),
                                     B, h, w, arr, ctypes.c_float(0.1), linear, scale, P(nh), P(nl), ncta)
    assert rc == 0
    assert rel_err(ll, LL) < 2e-6 and rel_err(yh, ref) < 2e-6      # also proves no NaN-poisoned slot was read
    xr = torch.full((B, 1, h, w), nan)
    rc = emul.ll_emul_lift_level_inv(P(LL.contiguous()), I64(h * w // 4), P(ref), I64(3 * h * w // 4), P(xr), I64(h * w),
                                     P(scratch), B, h, w, arr, ctypes.c_float(0.1), linear, scale, P(nh), P(nl), ncta)
    assert rc == 0
    assert rel_err(xr, rec) < 2e-6


@pytest.mark.parametrize("N,h,w", [(2, 32, 48), (1, 40, 136), (3, 4, 6), (1, 2, 2), (1, 8, 12), (2, 6, 10), (1, 64, 260)])
def test_emulated_dwt97_matches_oracle(emul, N, h, w):
    torch.manual_seed(N + h + w)
    x = torch.rand(N, 1, h, w) - 0.5
    yl, yh = tp.dwt97_forward(x, 1)
    nan = float("nan")
    ll = torch.full((N, h // 2, w // 2), nan)
    y = torch.full((N, 3, h // 2, w // 2), nan)
    emul.ll_emul_dwt97_fwd_level(P(x), I64(h * w), P(ll), I64(h * w // 4), P(y), I64(3 * h * w // 4), N, h, w)
    assert (ll - yl[:, 0]).abs().max().item() < 1e-6 and (y - yh[0][:, 0]).abs().max().item() < 1e-6
    xr = torch.full((N, h, w), nan)
    ylc, yhc = yl[:, 0].contiguous(), yh[0][:, 0].contiguous()
    emul.ll_emul_dwt97_inv_level(P(ylc), I64(h * w // 4), P(yhc), I64(3 * h * w // 4), P(xr), I64(h * w), N, h, w)
    assert (xr - tp.dwt97_inverse(yl, yh)[:, 0]).abs().max().item() < 1e-6


def test_oracle_known_answers():
    """KATs of SURVEY.md section 4 on the oracle itself."""
    # zeroed CNN output => plain CDF 9/7 lifting with zero extension; K^2 ratios vs the filter bank
    sd = lifting_weights()
    pfx = "m.autoencoder.waveletForward.0."
    for k in list(sd):
        if "conv4" in k:
            sd[k] = torch.zeros_like(sd[k])
    taps = [[0.0, olift.LIFTING_COEFF[0], olift.LIFTING_COEFF[0]], [olift.LIFTING_COEFF[1], olift.LIFTING_COEFF[1], 0.0],
            [0.0, olift.LIFTING_COEFF[2], olift.LIFTING_COEFF[2]], [olift.LIFTING_COEFF[3], olift.LIFTING_COEFF[3], 0.0]]
    for k in range(4):
        sd[pfx + f"convBlock.{k}.weight"] = torch.tensor(taps[k]).view(1, 1, 3, 1)
    cfg = om.default_cfg()
    torch.manual_seed(2)
    x = torch.rand(1, 1, 64, 64) - 0.5
    LL, LH, HL, HH = olift.one_level_forward(x, sd, pfx, cfg)
    yl, yh = tp.dwt97_forward(x, 1)
    K = olift.LIFTING_COEFF[5]
    s = slice(8, -8)
    assert (yl[0, 0, s, s] - K * K * LL[0, 0, s, s]).abs().max() < 1e-5
    assert (yh[0][0, 0, 0, s, s] + LH[0, 0, s, s]).abs().max() < 1e-5
    assert (yh[0][0, 0, 1, s, s] + HL[0, 0, s, s]).abs().max() < 1e-5
    assert (yh[0][0, 0, 2, s, s] - HH[0, 0, s, s] / (K * K)).abs().max() < 1e-5
    # the filter bank restatement: conv form == closed-form periodic sums; perfect reconstruction
    ll_d, yh_d = tp.dwt97_forward_direct(x, 1)
    assert abs(yl.double().numpy() - ll_d).max() < 1e-6
    assert abs(yh[0].double().numpy() - yh_d[0]).max() < 1e-6
    assert (tp.dwt97_inverse(yl, yh) - x).abs().max() < 1e-5
    # Gaussian likelihood vs scipy
    from scipy.stats import norm
    y = torch.tensor([0.0, 1.0, -3.0, 7.0])
    sg = torch.tensor([0.05, 1.0, 2.0, 3.0])
    mu = torch.tensor([0.2, -0.5, 0.0, 1.0])
    lik = tp.gaussian_likelihood(y, sg, mu)
    s2 = torch.clamp(sg, min=0.11).double().numpy()
    v = (y - mu).abs().double().numpy()
    want = norm.cdf((0.5 - v) / s2) - norm.cdf((-0.5 - v) / s2)
    assert abs(lik.double().numpy() - want).max() < 1e-6


def test_agent_mirrors_are_importable_and_refuse_to_run_unconfigured():
    """``LiftingBasedDWTAgent(config)`` / ``CompressionAgent(config)`` keep the reference's names; without
    a loader (or, for CompressionAgent, a model -- the reference leaves it None) they raise instead of doing CPU work."""
    import pytest
    from oracle import model as om
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.agents import CompressionAgent, LiftingBasedDWTAgent
    cfg = om.default_cfg(dwtlevels=2, autoencoder="SubbandAutoEncoder")
    agent = LiftingBasedDWTAgent(cfg, device="cpu")
    assert agent.clrch == 1 and agent.lambda_ == cfg.lambda_ and len(list(agent.model.parameters())) > 0
    with pytest.raises(RuntimeError):
        agent.validate()
    with pytest.raises(RuntimeError):
        CompressionAgent(cfg, device="cpu").train_one_epoch()
    import torch
    with pytest.raises(RuntimeError):
        agent.preprocess(torch.rand(1, 3, 8, 8))          # CPU tensor: the kernels have no CPU fallback


# ---------------------------------------------------------------------------------------------- config surface
def test_config_surface_matches_reference_keys(tmp_path):
    """utils/config.py + liftingDWT.json (reference: utils/config.py:50-103, liftingDWT.json:1-53)."""
    import json
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C
    cfg, d = C.get_config_from_json(C.DEFAULT_JSON)
    assert isinstance(d, dict) and cfg.agent == "LiftingBasedDWTAgent" and cfg.dwtlevels == 4 and cfg["lambda_"] == 11700
    with pytest.raises(AttributeError):
        cfg.no_such_key
    ref_json = "/root/reference/liftingDWT.json"
    if os.path.isfile(ref_json):            # dev container only: same keys, same values except the machine-specific paths
        ref = json.load(open(ref_json))
        assert list(ref) == list(d)
        for k, v in ref.items():
            if not (k.startswith("train_data_") or k in ("test_data", "valid_data")):
                assert d[k] == v, k
    cfg.exp_name = "unit"
    out = C.process_config(cfg, root=str(tmp_path), quiet=True)
    for k in ("summary_dir", "checkpoint_dir", "out_dir", "log_dir"):
        assert os.path.isdir(out[k]) and out[k].endswith(os.sep)
    bad = tmp_path / "bad.json"
    bad.write_text("{not json")
    with pytest.raises(ValueError):
        C.get_config_from_json(str(bad))
    c3 = C.baseline_config("cfg3")
    assert c3.netType == "LiftingBasedNeuralWaveletv4" and c3.autoencoder == "SubbandAutoEncoderBerk"


def test_product_synthetic_weights_equal_the_oracle_recipe():
    """utils/synthetic.keyed_weights (bench / smoke) == oracle.model.keyed_weights (what the goldens were made with)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        LiftingBasedDWTNetWrapper
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C, synthetic as S
    for name, lv in (("cfg3", 2), ("cfg1", 2)):
        cfg = C.baseline_config(name, dwtlevels=lv)
        torch.manual_seed(1337)
        m = LiftingBasedDWTNetWrapper(cfg)
        a, b = S.keyed_weights(m.state_dict()), om.keyed_weights(m.state_dict())
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)
        m.load_state_dict(a, strict=True)
    v1a, v1b = S.amplify_v1(m.state_dict()), om.amplify_v1(m.state_dict())
    assert all(torch.equal(v1a[k], v1b[k]) for k in v1a)


class _StubModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.ones(3))


def test_base_agent_checkpoint_and_run_semantics(tmp_path):
    """agents/base.py mirrors the reference's BaseAgent (agents/base.py:64-168): checkpoint dict keys, best copy,
    resume of model + loggers but not optimizer, run() saving a checkpoint on an exception and re-raising."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.agents.base import BaseAgent
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.loggers import RDLogger
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils.config import EasyDict

    class Agent(BaseAgent):
        def __init__(self, config):
            super().__init__(config, device="cpu")
            self.model = _StubModel()
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=self.lr)
            self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer)
            self.train_logger, self.trnit_logger, self.valid_logger, self.test_logger = (RDLogger() for _ in range(4))
            self.losses = [3.0, 2.0, 2.5]
            self.boom = False

        def train_one_epoch(self):
            if self.boom:
                raise ValueError("boom")
            self.train_logger(1.0, 0.1, 0.5, 0.4)
            self.current_iteration += 1

        def validate(self):
            return self.losses[self.current_epoch]

    cfg = EasyDict(mode="train", max_epoch=3, validate_every=1, checkpoint_dir=str(tmp_path) + os.sep, seed=1,
                   learning_rate=1e-3, gpu_device=0)
    a = Agent(cfg)
    a.run()
    a.finalize()
    ck = torch.load(str(tmp_path / "checkpoint.pth.tar"), weights_only=False)
    assert set(ck) == {"epoch", "iteration", "best_valid_loss", "state_dict", "optimizer", "scheduler", "train_logger",
                       "trnit_logger", "valid_logger", "test_logger"}
    assert a.best_valid_loss == 2.0 and os.path.isfile(str(tmp_path / "model_best.pth.tar"))
    assert torch.load(str(tmp_path / "model_best.pth.tar"), weights_only=False)["best_valid_loss"] == 2.0
    b = Agent(cfg)
    with torch.no_grad():
        a.model.w.mul_(2)
    a.save_checkpoint()
    assert b.load_checkpoint("checkpoint.pth.tar") and torch.equal(b.model.w, a.model.w)
    assert b.current_iteration == a.current_iteration and b.train_logger.state_dict() == a.train_logger.state_dict()
    assert b.load_checkpoint("missing.pth.tar") is False          # skipped silently, like the reference
    os.remove(str(tmp_path / "checkpoint.pth.tar"))
    b.boom = True
    b.current_epoch = 0
    with pytest.raises(ValueError):
        b.run()
    assert os.path.isfile(str(tmp_path / "checkpoint.pth.tar"))   # exception -> checkpoint -> re-raise
    cfg2 = EasyDict(cfg, mode="nonsense")
    with pytest.raises(NameError):
        Agent(cfg2).run()


def test_rdlogger_state_dict_layout():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.loggers import RDLogger
    lg = RDLogger()
    lg(2.0, 0.01, 0.5, 0.25)
    lg(4.0, 0.03, 1.5, 0.0)
    assert set(lg.state_dict()) == {"loss", "mse", "rate", "rate2", "it", "ep"} and lg.state_dict()["it"] == 2
    loss, mse, rate, rate2 = lg.display(lr=1e-4, typ="tr")
    assert (loss, rate, rate2) == (3.0, 1.0, 0.25) and abs(mse - 0.02) < 1e-12 and lg.current_epoch == 1
