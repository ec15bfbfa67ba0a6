"""GPU tests (``-m gpu``) of the agent-side pointwise kernels and of the ``LiftingBasedDWTAgent`` mirror:
colour transform, clamp and squared error against the oracle's restatement of compressai.transforms (BT.709) and of
agents/liftingDWT_agent.py:164-186; the agent's validate scalars against the oracle pipeline on the same weights."""
import math

import pytest
import torch

from oracle import model as om, thirdparty as tp

from common import keyed_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    return ops


@pytest.mark.parametrize("shape", [(2, 3, 64, 96), (1, 3, 7, 9), (3, 3, 1, 5), (1, 3, 512, 768)])
def test_colour_kernels_match_oracle(shape):
    ops = _ops()
    torch.manual_seed(sum(shape))
    x = torch.rand(*shape)
    ref = om.preprocess(x)                                   # RGB2YCbCr, Y - 0.5 (agents/liftingDWT_agent.py:170-171)
    got = ops.rgb_to_ycbcr_shift(x.to(DEV)).cpu()
    assert (got - ref).abs().max().item() <= 2e-7            # same IEEE operations: at most an ulp
    # post-processing of a perturbed "reconstruction": + (0.5,0,0), YCbCr2RGB, - 0.5, clamp, squared error (:174-186)
    yhat = ref + 0.3 * torch.randn(*shape)                   # large enough to hit the clamp
    shift = torch.tensor([[[0.5]], [[0.0]], [[0.0]]])
    oxhat = (tp.ycbcr2rgb(yhat + shift) - 0.5).clamp(-0.5, 0.5)
    osse = ((x - 0.5 - oxhat).double() ** 2).sum(dim=(1, 2, 3))
    xhat, sse = ops.ycbcr_to_rgb_sse(yhat.to(DEV), x.to(DEV))
    assert (xhat.cpu() - oxhat).abs().max().item() <= 4e-7
    assert ((xhat.cpu() == 0.5) | (xhat.cpu() == -0.5)).any()       # the clamp was exercised
    assert torch.allclose(sse.cpu(), osse, rtol=2e-5, atol=1e-9)
    _, sse2 = ops.ycbcr_to_rgb_sse(yhat.to(DEV), x.to(DEV), want_xhat=False)
    assert torch.allclose(sse2.cpu(), osse, rtol=2e-5, atol=1e-9)
    xh3, none = ops.ycbcr_to_rgb_sse(yhat.to(DEV))
    assert none is None and torch.equal(xh3, xhat)
    # exact inverse pair up to rounding
    back, _ = ops.ycbcr_to_rgb_sse(ops.rgb_to_ycbcr_shift(x.to(DEV)))
    assert (back.cpu() - (x - 0.5)).abs().max().item() <= 1e-6


def test_colour_kernel_errors():
    ops = _ops()
    with pytest.raises(RuntimeError):
        ops.rgb_to_ycbcr_shift(torch.rand(1, 3, 4, 4))                       # CPU tensor: no fallback
    with pytest.raises(ValueError):
        ops.rgb_to_ycbcr_shift(torch.rand(1, 1, 4, 4, device=DEV))
    with pytest.raises(ValueError):
        ops.ycbcr_to_rgb_sse(torch.rand(1, 3, 4, 4, device=DEV), torch.rand(1, 3, 4, 5, device=DEV))
    assert ops.rgb_to_ycbcr_shift(torch.rand(0, 3, 4, 4, device=DEV)).shape == (0, 3, 4, 4)


def test_agent_validate_batch_matches_oracle_pipeline():
    """``LiftingBasedDWTAgent.validate_batch`` (pre-processing, codec forward, post-processing, R-D scalars) against the
    oracle: preprocess -> wrapper_forward -> ycbcr2rgb / clamp -> TrainRDLoss.forward3, same synthetic weights."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.agents import CompressionAgent, LiftingBasedDWTAgent
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                         entropy_layer="conditioned2ZTsepSubbands", dwtlevels=2, ctx_precision="fp32")
    cfg.learning_rate, cfg.lambda_ = 1e-4, 0.01
    torch.manual_seed(1337)
    agent = LiftingBasedDWTAgent(cfg, device=DEV)
    sd = keyed_state(agent.model.cpu())
    agent.model.to(DEV)
    torch.manual_seed(4)
    x = torch.rand(2, 3, 64, 96)
    got = agent.validate_batch(x)
    with torch.no_grad():
        y = om.preprocess(x)
        oyhat, osi_xe, osi_xo = om.wrapper_forward(y, sd, cfg)
        shift = torch.tensor([[[0.5]], [[0.0]], [[0.0]]])
        oxhat = (tp.ycbcr2rgb(oyhat + shift) - 0.5).clamp(-0.5, 0.5)
        loss, mse, r1, r2 = om.rd_loss(x - 0.5, oxhat, osi_xe, osi_xo, cfg.lambda_)
    assert abs(got["bpp"] - float(r1 + r2)) <= 1e-3 * float(r1 + r2)            # bpp within 0.1 %
    assert abs(got["mse"] - float(mse)) <= 1e-3 * float(mse)
    assert abs(got["psnr"] - 10.0 * math.log10(1.0 / float(mse))) <= 0.01
    assert abs(got["rd_loss"] - float(loss)) <= 1e-3 * float(loss)
    # loaders: validate() averages the batches; the CompressionAgent skeleton refuses to run without a model
    agent.data_loader = [x, x]
    lr0 = agent.optimizer.param_groups[0]["lr"]
    steps0 = agent.scheduler.last_epoch
    valid_rd_loss = agent.validate()                      # the reference returns the mean validation loss (:201)
    assert abs(valid_rd_loss - got["rd_loss"]) <= 1e-9 * max(1.0, got["rd_loss"])
    assert abs(agent.last_validation["bpp"] - got["bpp"]) <= 1e-9 * max(1.0, got["bpp"])
    # validate() has no side effect on the learning-rate schedule (the reference steps it in train_one_epoch, :110-111)
    assert agent.scheduler.last_epoch == steps0 and agent.optimizer.param_groups[0]["lr"] == lr0
    with pytest.raises(RuntimeError):
        CompressionAgent(cfg, device=DEV).validate()
    # optional ``cuda_graph`` key: the same batch through a captured graph gives the same scalars, also on replay
    agent.cuda_graph = True
    for _ in range(2):
        again = agent.validate_batch(x)
        assert all(again[k] == got[k] for k in got), (again, got)


def test_agent_train_batch_steps():
    """One optimisation step of ``train_one_epoch`` (:78-125): finite losses, parameters move."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.agents import LiftingBasedDWTAgent
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                         entropy_layer="conditioned2ZTsepSubbands", dwtlevels=2)
    cfg.learning_rate, cfg.lambda_ = 1e-4, 0.01
    torch.manual_seed(1337)
    agent = LiftingBasedDWTAgent(cfg, device=DEV)
    keyed_state(agent.model.cpu())
    agent.model.to(DEV)
    before = [p.detach().clone() for p in agent.model.parameters()]
    torch.manual_seed(6)
    rd, mse, r1, r2 = agent.train_batch(torch.rand(2, 3, 32, 32))
    assert all(torch.isfinite(v).all() for v in (rd, mse, r1, r2))
    moved = sum(int(not torch.equal(a, b.detach())) for a, b in zip(before, agent.model.parameters()))
    assert moved > 0


def test_agent_epoch_semantics_follow_the_reference():
    """train_one_epoch (:75-111): zero_grad + step on every batch with the loss scaled by 1/grad_acc_iters, the
    distortion-only -> R+lambda*D switch at loss_prnt_iters, ReduceLROnPlateau stepped once per epoch on the mean
    training loss; BaseAgent.train: validate every epoch, checkpoint, best copy."""
    import os
    import tempfile
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.agents import LiftingBasedDWTAgent
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.losses.rate_dist import TrainDLoss, TrainRDLoss
    cfg = om.default_cfg(netType="CDF97", entropy_layer="factorized", dwtlevels=1)
    tmp = tempfile.mkdtemp()
    cfg.update(learning_rate=1e-4, lambda_=0.01, mode="train", max_epoch=2, validate_every=1, grad_acc_iters=2,
               training_loss_switch=0, loss_switch_thr=1e9, loss_prnt_iters=2, checkpoint_dir=tmp + os.sep)
    torch.manual_seed(1337)
    torch.manual_seed(3)
    batches = [torch.rand(1, 3, 16, 16) for _ in range(3)]
    agent = LiftingBasedDWTAgent(cfg, data_loader=batches, device=DEV)
    assert isinstance(agent.train_loss, TrainDLoss)
    steps = []
    orig_step = agent.optimizer.step
    agent.optimizer.step = lambda *a, **k: (steps.append(1), orig_step(*a, **k))[1]
    agent.run()
    agent.finalize()
    assert len(steps) == 6                                  # every batch steps (grad_acc_iters only scales the loss)
    assert isinstance(agent.train_loss, TrainRDLoss) and agent.training_loss_switch == 1     # switched (:103-109)
    assert agent.scheduler.last_epoch == 2                  # one scheduler step per epoch
    assert agent.current_iteration == 6 and agent.current_epoch == 2
    assert os.path.isfile(os.path.join(tmp, "checkpoint.pth.tar")) and os.path.isfile(os.path.join(tmp, "model_best.pth.tar"))
    ck = torch.load(os.path.join(tmp, "checkpoint.pth.tar"), weights_only=False)
    assert list(ck["state_dict"]) == list(agent.model.state_dict())
    # resume: a fresh agent picks the checkpoint up (model + counters + loggers)
    cfg2 = om.default_cfg(netType="CDF97", entropy_layer="factorized", dwtlevels=1)
    cfg2.update(cfg, resume_training=True, checkpoint_file="checkpoint.pth.tar")
    again = LiftingBasedDWTAgent(cfg2, data_loader=batches, device=DEV)
    assert again.current_iteration == 6
    assert all(torch.equal(a, b) for a, b in zip(again.model.state_dict().values(), agent.model.state_dict().values()))
