"""Transform layers of the codec (reference: graphs/layers/lifting_dwt_nets.py).

``LiftingBasedNeuralWaveletv4`` (:646-827), ``DWTPytorchWaveletsLayer`` (:212-277),
``SubbandAutoEncoder`` (:82-125), ``SubbandAutoEncoderBerk`` (:126-165) with the same
constructors, sub-module names and parameter registration order as the reference, so
``state_dict`` keys (2 136 for the 4-level learned model) and seeded initialisation match.
"""
import torch
import torch.nn as nn

from ... import _autograd, ops
from ...compat import GDN, DWTForward, DWTInverse
from ._packing import PackCache
from .P_block_v2 import P_block_v2
from .wavelet_forward_v2 import wavelet_forward_v2
from .wavelet_inverse_v2 import wavelet_inverse_v2

lifting_coeff = [-1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971, 0.869864451624781,
                 1.149604398860241]  # bior4.4


def get_cdf97_filters(oned_or_twod='2D'):
    """(lifting_dwt_nets.py:414-430) the in-repo pin of the 9/7 taps."""
    h_ana_lp = torch.tensor([0.0, 0.037828455507264, -0.023849465019557, -0.110624404418437, 0.377402855612831, 0.852698679008894, 0.377402855612831, -0.110624404418437, -0.023849465019557, 0.037828455507264], dtype=torch.double)
    h_ana_hp = torch.tensor([0.0, -0.064538882628697, 0.040689417609164, 0.418092273221617, -0.788485616405583, 0.418092273221617, 0.040689417609164, -0.064538882628697, 0.0, 0.0], dtype=torch.double)
    h_syn_lp = torch.tensor([0.0, -0.064538882628697, -0.040689417609164, 0.418092273221617, 0.788485616405583, 0.418092273221617, -0.040689417609164, -0.064538882628697, 0.0, 0.0], dtype=torch.double)
    h_syn_hp = torch.tensor([0.0, -0.037828455507264, -0.023849465019557, 0.110624404418437, 0.377402855612831, -0.852698679008894, 0.377402855612831, 0.110624404418437, -0.023849465019557, -0.037828455507264], dtype=torch.double)
    if oned_or_twod == '1D':
        return h_ana_lp, h_ana_hp, h_syn_lp, h_syn_hp
    outer = lambda a, b: a.view(-1, 1) * b.view(1, -1)
    return (outer(h_ana_lp, h_ana_lp), outer(h_ana_lp, h_ana_hp), outer(h_ana_hp, h_ana_lp), outer(h_ana_hp, h_ana_hp),
            outer(h_syn_lp, h_syn_lp), outer(h_syn_lp, h_syn_hp), outer(h_syn_hp, h_syn_lp), outer(h_syn_hp, h_syn_hp))


class SubbandAutoEncoder(nn.Module):
    """Pointwise "scaling network": per channel 1 -> 32 -> 32 -> 32 -> 1 with tanh
    (lifting_dwt_nets.py:82-125), as one fused CUDA kernel (csrc/subband_ae.cu)."""

    def __init__(self, in_ch):
        super().__init__()
        K, P, iC, H = 1, 0, in_ch, 32
        self.in_ch = in_ch
        self.ae_down = nn.Sequential(
            nn.Conv2d(iC * 1, iC * H, kernel_size=K, stride=1, padding=P, groups=iC), nn.Tanh(),
            nn.Conv2d(iC * H, iC * H, kernel_size=K, stride=1, padding=P, groups=iC), nn.Tanh(),
            nn.Conv2d(iC * H, iC * H, kernel_size=K, stride=1, padding=P, groups=iC), nn.Tanh(),
            nn.Conv2d(iC * H, iC * 1, kernel_size=K, stride=1, padding=P, groups=iC))
        self.ae_up = nn.Sequential(
            nn.ConvTranspose2d(iC * 1, iC * H, kernel_size=K, stride=1, padding=P, groups=iC, output_padding=0), nn.Tanh(),
            nn.ConvTranspose2d(iC * H, iC * H, kernel_size=K, stride=1, padding=P, groups=iC, output_padding=0), nn.Tanh(),
            nn.ConvTranspose2d(iC * H, iC * H, kernel_size=K, stride=1, padding=P, groups=iC, output_padding=0), nn.Tanh(),
            nn.ConvTranspose2d(iC * H, iC * 1, kernel_size=K, stride=1, padding=P, groups=iC, output_padding=0))
        self._down_cache, self._up_cache = PackCache(), PackCache()

    def _blob(self, seq, cache, transposed):
        layers = [(seq[k].weight, seq[k].bias) for k in (0, 2, 4, 6)]
        flat = [t for wb in layers for t in wb]
        return cache.get(flat, lambda: ops.pack_ae1(layers, self.in_ch, transposed))

    def encode(self, x):
        return _autograd.run_module(lambda x: ops.ae1_apply(x, self._blob(self.ae_down, self._down_cache, False)), self.ae_down, x)

    def encode_and_round(self, x):
        """(y, round(y)): the quantiser's rounding fused into the same pass."""
        return ops.ae1_apply(x, self._blob(self.ae_down, self._down_cache, False), want_round=True)

    def decode(self, y_hat):
        return _autograd.run_module(lambda y: ops.ae1_apply(y, self._blob(self.ae_up, self._up_cache, True)), self.ae_up, y_hat)


class SubbandAutoEncoderBerk(nn.Module):
    """3x3 conv + GDN / inverse-GDN scaling network (lifting_dwt_nets.py:126-165).  It feeds the quantiser, so it runs
    at fp32-level accuracy: convs 1-3 and the three GDN norms as 3xTF32 tcgen05 implicit GEMMs (csrc/igemm_conv.cu: every
    conv fused with its GDN, the first one in the kernel's head mode), the last conv on the exact FP32 kernel.  With autograd on (training), forward and backward run through torch
    fp32 ops (TF32 off) -- the recompute path of ``_autograd`` -- there is no inference-time backend switch."""

    def __init__(self, in_ch):
        super().__init__()
        K, P, iC, H = 3, 1, in_ch, 64
        self.ae_down = nn.Sequential(
            nn.Conv2d(iC * 1, iC * H // 2, kernel_size=K, stride=1, padding=P), GDN(iC * H // 2),
            nn.Conv2d(iC * H // 2, iC * H, kernel_size=K, stride=1, padding=P), GDN(iC * H),
            nn.Conv2d(iC * H, iC * H // 2, kernel_size=K, stride=1, padding=P), GDN(iC * H // 2),
            nn.Conv2d(iC * H // 2, iC * 1, kernel_size=K, stride=1, padding=P))
        self.ae_up = nn.Sequential(
            nn.ConvTranspose2d(iC * 1, iC * H // 2, kernel_size=K, stride=1, padding=P), GDN(iC * H // 2, inverse=True),
            nn.ConvTranspose2d(iC * H // 2, iC * H, kernel_size=K, stride=1, padding=P), GDN(iC * H, inverse=True),
            nn.ConvTranspose2d(iC * H, iC * H // 2, kernel_size=K, stride=1, padding=P), GDN(iC * H // 2, inverse=True),
            nn.ConvTranspose2d(iC * H // 2, iC * 1, kernel_size=K, stride=1, padding=P))
        self._down_cache, self._up_cache = PackCache(), PackCache()

    @staticmethod
    def _exact(fn, x):
        if not x.is_cuda:
            raise RuntimeError("SubbandAutoEncoderBerk: CUDA tensors only (no CPU fallback)")
        old_c, old_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            return fn(x)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_c, old_m

    # ---- tensor-core path: convs 2-3 and the three GDN norms as 3xTF32 implicit GEMMs (fp32-level accuracy) ----
    AE_BATCH_CHUNK = 32    # images per launch group: bounds the channels-last fp32 intermediates (4.8 GB per layer output at 256x384)
    fuse_gdn = True        # convs 2-3 run fused with their GDN (ll_igemm_tf32_gdn); False = the two-kernel chain (A/B checks)

    def _pack(self, seq, transposed, cache):
        convs = [seq[0], seq[2], seq[4], seq[6]]
        gdns = [seq[1], seq[3], seq[5]]
        srcs = [c.weight for c in convs] + [g.beta for g in gdns] + [g.gamma for g in gdns]

        def build():
            eq = (lambda w: w.detach().transpose(0, 1).flip(2, 3).contiguous()) if transposed else (lambda w: w.detach())
            pk = {"w0": eq(convs[0].weight), "w3": eq(convs[3].weight),
                  "wp": [ops.pack_tf32_weight(convs[1].weight, transposed), ops.pack_tf32_weight(convs[2].weight, transposed)],
                  "gdn": []}
            for g in gdns:
                C = g.beta.numel()
                beta = g.beta_reparam(g.beta.detach()).contiguous()
                gamma = g.gamma_reparam(g.gamma.detach()).reshape(C, C, 1, 1).contiguous()
                pk["gdn"].append((ops.pack_tf32_weight(gamma), beta))
            return pk

        return cache.get(srcs, build)

    def _run_tc(self, seq, x, transposed, cache):
        pk = self._pack(seq, transposed, cache)
        convs = [seq[0], seq[2], seq[4], seq[6]]
        inv = seq[1].inverse
        outs = []
        for b0 in range(0, x.shape[0], self.AE_BATCH_CHUNK):
            xb = x[b0:b0 + self.AE_BATCH_CHUNK].contiguous()
            c0 = pk["w0"].shape[0]
            if self.fuse_gdn and c0 in ops.GDN_FUSED_WIDTHS and xb.shape[1] <= 3:
                # first conv (K = 9 iC, exact FP32 FMA) + its GDN in one kernel
                z = ops.conv3_gdn_head(xb, pk["w0"], convs[0].bias, pk["gdn"][0][0], pk["gdn"][0][1], inverse=inv)
            else:
                c1 = ops.conv2d(xb, pk["w0"], convs[0].bias)                               # exact fp32 SIMT (K = 9 * iC)
                y, s = ops.nchw_to_nhwc_split(c1, squares=True)
                del c1
                _, z = ops.igemm_tf32(s, pk["gdn"][0][0], pk["gdn"][0][1], y.shape[3], epi=2, inverse=inv, y=y)
                del y, s
            for k in (0, 1):
                cout = convs[1 + k].weight.shape[1 if transposed else 0]
                if self.fuse_gdn and cout in ops.GDN_FUSED_WIDTHS:
                    # conv + GDN in one kernel: the conv output, its square and the norm stay in tensor memory
                    z = ops.igemm_tf32_gdn(z, pk["wp"][k], convs[1 + k].bias, pk["gdn"][1 + k][0], pk["gdn"][1 + k][1], cout, inverse=inv)
                else:
                    y, s = ops.igemm_tf32(z, pk["wp"][k], convs[1 + k].bias, cout, epi=1)
                    _, z = ops.igemm_tf32(s, pk["gdn"][1 + k][0], pk["gdn"][1 + k][1], cout, epi=2, inverse=inv, y=y)
                    del y, s
            outs.append(ops.nhwc_split_conv3(z, pk["w3"], convs[3].bias))     # exact fp32, straight from the chain's layout
            del z
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def _use_tc(self, x):
        return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))

    def encode(self, x):
        if self._use_tc(x):
            return self._run_tc(self.ae_down, x, False, self._down_cache)
        return self._exact(self.ae_down, x)

    def decode(self, y_hat):
        if self._use_tc(y_hat):
            return self._run_tc(self.ae_up, y_hat, True, self._up_cache)
        return self._exact(self.ae_up, y_hat)


class DWTPytorchWaveletsLayer(nn.Module):
    """``netType: "CDF97"``: J-level periodised 9/7 filter bank + pointwise subband
    auto-encoders (lifting_dwt_nets.py:212-277)."""

    def __init__(self, config):
        super().__init__()
        self.dwtlevels = config.dwtlevels
        self.clrch = config.clrch
        self.conv_grps = config.clrch
        self.xfm = DWTForward(J=self.dwtlevels, mode='periodization', wave='bior4.4')
        self.ifm = DWTInverse(mode='periodization', wave='bior4.4')
        self.Yl_ae = SubbandAutoEncoder(in_ch=1 * config.clrch)
        self.Yh_ae = nn.ModuleList()
        for i in range(0, self.dwtlevels):
            self.Yh_ae.append(SubbandAutoEncoder(in_ch=3 * config.clrch))

    def encode(self, x):
        Yl, Yh = self.xfm(x)
        out_xe = self.Yl_ae.encode(Yl)
        out_xo_list = []
        for i in range(0, self.dwtlevels):
            B, C, Three, H, W = Yh[i].shape
            out_xo_list.append(self.Yh_ae[i].encode(Yh[i].view(B, C * 3, H, W)))
        return out_xe, out_xo_list

    def decode(self, out_xe, out_xo_list):
        Yl = self.Yl_ae.decode(out_xe)
        Yh = []
        for i in range(0, self.dwtlevels):
            Yh_ae = self.Yh_ae[i].decode(out_xo_list[i])
            B, C, H, W = out_xo_list[i].shape
            Yh.append(Yh_ae.view(B, C // 3, 3, H, W))
        return self.ifm((Yl, Yh))


class LiftingBasedNeuralWaveletv4(nn.Module):
    """Multi-level learned lifting transform (lifting_dwt_nets.py:646-827)."""

    def __init__(self, config):
        super().__init__()
        if config.clrch != 1:
            raise ValueError("LiftingBasedNeuralWaveletv4 only works with clrch == 1 "
                             "(the reference forces (1,1,3,1) pre-filter weights, lifting_dwt_nets.py:785-819)")
        self.waveletLevel = config.dwtlevels
        self.liftingLevel = config.num_lifting_perlayer
        self.blockprop = config.block_property
        self.clrch = config.clrch
        self.linearityflag = config.linearity_flag
        self.conv_filter_size = config.filtersize
        self.postprocessflag = config.postprocess
        self.res_connection_weight = config.res_connection_weight
        self.P_blocks = nn.ModuleList()
        self.U_blocks = nn.ModuleList()
        self.waveletForward = nn.ModuleList()
        self.waveletInverse = nn.ModuleList()
        self.Yh_ae = nn.ModuleList()
        self.config = config
        # new optional key (default keeps reference configs valid): "tc" = conv2/conv3 of every lifting step
        # on tcgen05 with the 3xTF32 split (fp32-level accuracy), "fp32" = all layers on the FP32 FMA pipe
        self._lift_precision = config.get("lift_precision", ops.DEFAULT_LIFT_PRECISION) if hasattr(config, "get") else getattr(config, "lift_precision", ops.DEFAULT_LIFT_PRECISION)
        ops.lift_precision_code(self._lift_precision)
        self.depth_scale = config.depth_scale * 8
        self.preProcessingList = self.preProcessBlock(config.clrch, config.filtersize)
        if config.autoencoder == "SubbandAutoEncoder":
            self.Yl_ae = SubbandAutoEncoder(in_ch=1 * config.clrch)
            for i in range(0, self.waveletLevel):
                self.Yh_ae.append(SubbandAutoEncoder(in_ch=3 * config.clrch))
        elif config.autoencoder == "SubbandAutoEncoderBerk":
            self.Yl_ae = SubbandAutoEncoderBerk(in_ch=1 * config.clrch)
            for i in range(0, self.waveletLevel):
                self.Yh_ae.append(SubbandAutoEncoderBerk(in_ch=3 * config.clrch))

        if self.blockprop == 'same':
            numberOfBlocks = self.liftingLevel
        elif self.blockprop == 'different':
            numberOfBlocks = self.liftingLevel * 2 * self.waveletLevel
        else:
            raise ValueError(f"block_property {self.blockprop!r}")
        self.nh = nn.Parameter(nn.init.constant_(torch.empty(1, 1, 1, 1), 0.0), requires_grad=True)
        self.nl = nn.Parameter(nn.init.constant_(torch.empty(1, 1, 1, 1), 0.0), requires_grad=True)

        for _ in range(numberOfBlocks):
            self.P_blocks.append(P_block_v2(self.linearityflag, self.clrch, self.conv_filter_size, self.depth_scale))
            self.U_blocks.append(P_block_v2(self.linearityflag, self.clrch, self.conv_filter_size, self.depth_scale))

        L, n = self.waveletLevel, self.liftingLevel
        for lvl in range(L):
            if self.blockprop == 'same':
                Pf, Uf, Pi, Ui = self.P_blocks, self.U_blocks, self.P_blocks, self.U_blocks
            else:
                # forward level l: blocks [l n, (l+1) n); inverse: slice start fixed at L n (:712-722)
                Pf, Uf = self.P_blocks[lvl * n:(lvl + 1) * n], self.U_blocks[lvl * n:(lvl + 1) * n]
                Pi, Ui = self.P_blocks[L * n:L * n + (lvl + 1) * n], self.U_blocks[L * n:L * n + (lvl + 1) * n]
            self.waveletForward.append(wavelet_forward_v2(Pf, Uf, self.res_connection_weight, n, self.preProcessingList,
                                                          self.config, self.nh, self.nl))
            self.waveletInverse.append(wavelet_inverse_v2(Pi, Ui, self.res_connection_weight, n, self.preProcessingList,
                                                          self.config, self.nh, self.nl))

    @property
    def lift_precision(self):
        return self._lift_precision

    @lift_precision.setter
    def lift_precision(self, mode):
        """Per-module arithmetic of the lifting kernels (handed to every launch; nothing is process-wide)."""
        ops.lift_precision_code(mode)
        self._lift_precision = mode
        for m in list(self.waveletForward) + list(self.waveletInverse):
            m.lift_precision = mode

    def transform(self, input):
        """The lifting levels alone: x -> (LL, [Yh_l (B,3,h,w)])."""
        Yh = []
        ll = input
        for lvl in range(self.waveletLevel):
            ll, yh = self.waveletForward[lvl].level(ll)
            Yh.append(yh)
        return ll, Yh

    def inverse_transform(self, Yl, Yh):
        ll = Yl
        for lvl in range(self.waveletLevel - 1, -1, -1):
            ll = self.waveletInverse[lvl].level(ll, Yh[lvl])
        return ll

    def encode(self, input):
        """(lifting_dwt_nets.py:724-746)."""
        Yl, Yh = self.transform(input)
        out_xe = self.Yl_ae.encode(Yl)
        out_xo_list = [self.Yh_ae[i].encode(Yh[i]) for i in range(self.waveletLevel)]
        return out_xe, out_xo_list

    def decode(self, out_xe, out_xo_list):
        """(lifting_dwt_nets.py:748-782)."""
        Yl = self.Yl_ae.decode(out_xe)
        Yh = [self.Yh_ae[i].decode(out_xo_list[i]) for i in range(self.waveletLevel)]
        return self.inverse_transform(Yl, Yh)

    def preProcessBlock(self, csize, conv_filter_size):
        """3-tap pre-filters initialised to the CDF 9/7 lifting coefficients (:784-827)."""
        taps = [(0.0, lifting_coeff[0], lifting_coeff[0]), (lifting_coeff[1], lifting_coeff[1], 0.0),
                (0.0, lifting_coeff[2], lifting_coeff[2]), (lifting_coeff[3], lifting_coeff[3], 0.0)]
        params = [torch.tensor(([t[0]], [t[1]], [t[2]])).view(1, 1, 3, 1) for t in taps]
        convList = nn.ModuleList()
        convs = []
        for _ in range(4):
            convs.append(nn.Conv2d(1 * csize, 1 * csize, kernel_size=(3, 1), stride=1, padding=(3 // 2, 0), bias=False))
        for conv, p in zip(convs, params):
            conv.weight = torch.nn.Parameter(p, requires_grad=True)
            convList.append(conv)
        return convList
