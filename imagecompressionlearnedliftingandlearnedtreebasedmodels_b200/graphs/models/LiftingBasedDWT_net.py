"""Codec model and the tree-based subband entropy layers (reference:
graphs/models/LiftingBasedDWT_net.py).

``LiftingBasedDWTNetWrapper`` (:35-99), ``LiftingBasedDWTNet`` (:100-180),
``DWTFactorizedEntropyLayer`` (:182-231), ``DWTConditioned2EntropyLayerZTsepSubbands``
(:233-372), ``DWTConditioned2EntropyLayerZTBlock`` (:558-757), ``onlyEZWT`` (:759-840): same
constructors, sub-module names and registration order (checkpoints load with strict=True,
seeded construction reproduces the reference's initial weights); ``forward`` bodies run on the
sm_100a kernels.  The serial per-coefficient coder (``test`` / ``compress_ar`` /
``decompress_ar``, :374-556) is out of scope.
"""
import math

import torch
from torch import nn

from ... import ops
from ... import _autograd, _torch_ref, compat
from ...compat import EntropyBottleneck, GaussianConditional
from ..layers._packing import PackCache
from ..layers.lifting_dwt_nets import DWTPytorchWaveletsLayer, LiftingBasedNeuralWaveletv4
from ..layers.masked_conv2d import MaskedConv2d

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64
# images per launch group inside the context models: bounds the 243/486-channel intermediates
# (level 0 of a 512x768 plane needs ~0.6 GB per image in fp32)
CTX_BATCH_CHUNK = 16
CTX_TC_CHAIN = True     # coarsest-level causal chains: the two dense masked 3x3 layers on the BF16 tensor path
# A/B switches of the tensor-core context path (measurement scripts only; both on in the product):
CTX_GEMM_HEAD = True     # plc head / masked csc as 1-tap igemm layers over ops.ctx_im2col (else the fp32 SIMT ctx_conv_nhwc)
CTX_FUSED_TAIL = True    # cgp layers 2-4 + rate in one launch (else igemm_conv + cgp_tail_rate)


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def _conv(seq_item, x, lrelu=False, upsample2=False, **kw):
    """nn.Conv2d / MaskedConv2d through the direct-conv kernel (differentiable: backward recomputes the conv
    with torch ops, see _autograd.py; the channel-remapped ``out=`` form is inference-only)."""
    if isinstance(seq_item, MaskedConv2d):
        seq_item.apply_mask()
    w, b, g = seq_item.weight, seq_item.bias, seq_item.groups
    if _autograd.needs_grad([x, w, b]):
        if kw:
            raise RuntimeError("the channel-remapped conv output is not differentiable; use the plain form")
        return _autograd.run(lambda x, w, b: ops.conv2d(x, w, b, groups=g, lrelu=lrelu, upsample2=upsample2),
                             lambda x, w, b: _torch_ref.conv2d(x, w, b, g, lrelu, upsample2), [x, w, b])
    return ops.conv2d(x, w, b, groups=g, lrelu=lrelu, upsample2=upsample2, **kw)


def _chain(seq, x):
    """Conv / LeakyReLU alternation of an nn.Sequential with the activation fused into the conv."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        act = i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU)
        x = _conv(m, x, lrelu=act)
        i += 2 if act else 1
    return x


def _dep_net(seq, x):
    """A ZTBlock dependency CNN (:618-680): [Conv3x3, LReLU, Conv3x3, LReLU, Conv1x1, LReLU, Conv1x1, LReLU, Conv1x1].
    Inference: the two 3x3 convs on the direct-conv kernel, the pointwise tail (32->32->32->1) in one fused pass;
    with autograd on, the generic differentiable chain."""
    mods = list(seq)
    tail = len(mods) == 9 and all(isinstance(mods[i], nn.Conv2d) and mods[i].kernel_size == (1, 1) for i in (4, 6, 8)) \
        and mods[4].in_channels == 32 and mods[4].out_channels == 32 and mods[6].out_channels == 32 and mods[8].out_channels == 1
    if not tail or _autograd.needs_grad([x] + list(seq.parameters())):
        return _chain(seq, x)
    x = _conv(mods[0], x, lrelu=True)
    x = _conv(mods[2], x, lrelu=True)
    return ops.pw_mlp3(x, mods[4], mods[6], mods[8])


class LiftingBasedDWTNetWrapper(nn.Module):
    """Per-colour-plane dispatch (:35-99): ``clrch == 1`` builds three independent nets."""

    def __init__(self, config):
        super().__init__()
        self.clrch = config.clrch
        if self.clrch == 3:
            self.model = LiftingBasedDWTNet(config)
        elif self.clrch == 1:
            self.model0 = LiftingBasedDWTNet(config)
            self.model1 = LiftingBasedDWTNet(config)
            self.model2 = LiftingBasedDWTNet(config)
        # (the Berk scaling network streams multi-GB channels-last intermediates per plane: three of them in flight
        # only fight over HBM and the allocator, measured 250 -> 279 ms; the lighter configurations gain 5 %)
        ae = config.get("autoencoder", "") if hasattr(config, "get") else getattr(config, "autoencoder", "")
        self.plane_streams = not (config.netType == "LiftingBasedNeuralWaveletv4" and ae == "SubbandAutoEncoderBerk")
        self._streams = None

    def planes(self):
        return [self.model] if self.clrch == 3 else [self.model0, self.model1, self.model2]

    def set_bit_accumulator(self, acc):
        """Route sum(self-information) of every subband into ``acc`` (float64[1] on the device)."""
        for m in self.planes():
            m.entropymodel.bit_acc = acc

    def _forward_planes(self, x):
        """The three colour planes are independent networks (:52-54): in inference each one runs on its own CUDA
        stream, so the many small launches of the deep levels of one plane overlap the kernels of another."""
        if torch.is_grad_enabled() or not x.is_cuda or not self.plane_streams:
            return [m(x[:, c:c + 1, :, :]) for c, m in enumerate(self.planes())]
        if self._streams is None or self._streams[0].device != x.device:
            self._streams = [torch.cuda.Stream(device=x.device) for _ in range(3)]
        cur = torch.cuda.current_stream(x.device)
        outs = [None] * 3
        for c, m in enumerate(self.planes()):
            st = self._streams[c]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                outs[c] = m(x[:, c:c + 1, :, :].contiguous())
        for st in self._streams:
            cur.wait_stream(st)
        for o in outs:
            for t in [o[0], o[1]] + list(o[2]):
                t.record_stream(cur)
        return outs

    def forward(self, x):
        if self.clrch == 3:
            return self.model.forward(x)
        outs = self._forward_planes(x)
        xhat = torch.cat([o[0] for o in outs], dim=1)
        si_xe = torch.cat([o[1] for o in outs], dim=1)
        si_xo = []
        for o in outs:
            si_xo.extend(o[2])
        return xhat, si_xe, si_xo

    def display(self, x):
        pass

    def aux_loss(self):
        return sum(m.aux_loss() for m in self.planes())

    def compress(self, x):
        """(:76-99) x (B,3,H,W) -> (xhat, len_xe, len_xo): bits per pixel actually spent on the LL band and on the detail
        subbands, summed over the colour planes (per-stream length tables included); the bitstreams themselves are kept
        in ``self.last_bitstreams`` (one list per plane) for :meth:`decompress`."""
        if self.clrch != 1:
            xhat, streams = self.model.compress(x)
            per_plane = [streams]
        else:
            outs = [m.compress(x[:, c:c + 1].contiguous()) for c, m in enumerate(self.planes())]
            xhat = torch.cat([o[0] for o in outs], dim=1)
            per_plane = [o[1] for o in outs]
        self.last_bitstreams = per_plane
        B, _, H, W = x.shape
        len_xe = sum(st[0].nbytes() for st in per_plane) * 8 / (B * H * W)
        len_xo = sum(sum(t.nbytes() for t in st[1:]) for st in per_plane) * 8 / (B * H * W)
        return xhat, len_xe, len_xo

    def decompress(self, per_plane=None):
        per_plane = self.last_bitstreams if per_plane is None else per_plane
        if self.clrch != 1:
            return self.model.decompress(per_plane[0])
        return torch.cat([m.decompress(st) for m, st in zip(self.planes(), per_plane)], dim=1)


class LiftingBasedDWTNet(nn.Module):
    """Transform + entropy model of one plane (:100-180)."""

    def __init__(self, config):
        super().__init__()
        self.clrch = config.clrch
        if config.netType == "CDF97":
            self.autoencoder = DWTPytorchWaveletsLayer(config)
        elif config.netType == "LiftingBasedNeuralWaveletv4":
            self.autoencoder = LiftingBasedNeuralWaveletv4(config)
        else:
            raise ValueError(f"netType {config.netType!r} is outside the lifting hot path "
                             "(supported: 'CDF97', 'LiftingBasedNeuralWaveletv4')")
        self.entropy_layer = config.entropy_layer
        if self.entropy_layer == "factorized":
            self.entropymodel = DWTFactorizedEntropyLayer(config)
        elif self.entropy_layer == "onlyEZWT":
            self.entropymodel = onlyEZWT(config)
        elif self.entropy_layer == "conditioned2ZTsepSubbands":
            self.entropymodel = DWTConditioned2EntropyLayerZTsepSubbands(config)
        elif self.entropy_layer == "DWTConditioned2EntropyLayerZTBlock":
            self.entropymodel = DWTConditioned2EntropyLayerZTBlock(config)

    def compress(self, x):
        """x (B,C,H,W) -> (xhat, streams): real bitstreams for the entropy layers that decode a subband at a time
        (``factorized``, ``onlyEZWT``; interleaved rANS, ``coding.py``).  The reference's ``compress`` (:136-152) calls
        ``entropymodel.test``, which only exists for ``conditioned2ZTsepSubbands`` (its serial per-coefficient coder,
        :374-556) -- that one stays out of scope."""
        if not hasattr(self.entropymodel, "compress"):
            raise NotImplementedError(f"compress(): no parallel coder for entropy_layer={self.entropy_layer!r} "
                                      "(the serial per-coefficient coder of the reference is out of scope)")
        with torch.no_grad():
            out_xe, out_xo_list = self.autoencoder.encode(x)
            streams, xe_q, qs = self.entropymodel.compress(out_xe, out_xo_list)
            xhat = self.autoencoder.decode(xe_q, qs)
        return xhat, streams

    def decompress(self, streams):
        """Inverse of :meth:`compress`: bitstreams -> reconstruction (B,C,H,W)."""
        if not hasattr(self.entropymodel, "decompress"):
            raise NotImplementedError(f"decompress(): no parallel coder for entropy_layer={self.entropy_layer!r}")
        with torch.no_grad():
            xe_q, qs = self.entropymodel.decompress(streams)
            return self.autoencoder.decode(xe_q, qs)

    def forward(self, x):
        """x (B,C,H,W) -> (xhat, si_xe, si_xo_list) (:154-170)."""
        if self.entropy_layer not in ("factorized", "conditioned2ZTsepSubbands",
                                      "DWTConditioned2EntropyLayerZTBlock", "onlyEZWT"):
            raise ValueError
        out_xe, out_xo_list = self.autoencoder.encode(x)
        si_xe, si_xo_list, xe_qnt, xo_list_qnt = self.entropymodel(out_xe, out_xo_list)
        xhat = self.autoencoder.decode(xe_qnt, xo_list_qnt)
        return xhat, si_xe, si_xo_list

    def display(self, x):
        pass

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))


def _ctx_precision(config):
    """New optional config key (default keeps the reference's config files valid): "bf16" | "fp32"."""
    v = config.get("ctx_precision", "bf16") if hasattr(config, "get") else getattr(config, "ctx_precision", "bf16")
    if v not in ("bf16", "fp32"):
        raise ValueError(f"ctx_precision must be 'bf16' or 'fp32', got {v!r}")
    return v


def _sos_ses(config):
    se, so = 1 * config.clrch, 3 * config.clrch
    ses, sos = [], []
    for _ in range(config.dwtlevels):
        sos.append(so)
        ses.append(se)
        so = se * 3
        se = se * 1
    return ses, sos


class DWTFactorizedEntropyLayer(nn.Module):
    """Fully factorized model (:182-231)."""

    def __init__(self, config):
        super().__init__()
        self.num_lifting_layers = config.dwtlevels
        assert self.num_lifting_layers > 0
        self.clrch = config.clrch
        self.se, self.so = 1, 3
        se, so = self.se * self.clrch, self.so * self.clrch
        self.ent_out_xo_list = nn.ModuleList()
        self.scl_out_xo_list = nn.ParameterList()
        self.scb_out_xo_list = nn.ParameterList()
        for i in range(0, self.num_lifting_layers, 1):
            self.ent_out_xo_list.append(EntropyBottleneck(channels=so))
            self.scl_out_xo_list.append(nn.Parameter(nn.init.constant_(torch.empty(1, so, 1, 1), i + 1.0)))
            self.scb_out_xo_list.append(nn.Parameter(nn.init.constant_(torch.empty(1, so, 1, 1), 1.0)))
            so = se * self.so
            se = se * self.se
        self.ent_out_xe = EntropyBottleneck(channels=se / self.se)
        self.scl_out_xe = nn.Parameter(nn.init.constant_(torch.empty(1, int(se / self.se), 1, 1), 5.0))
        self.scb_out_xe = nn.Parameter(nn.init.constant_(torch.empty(1, int(se / self.se), 1, 1), 1.0 / 5.0))
        self.bit_acc = None

    def forward(self, out_xe, out_xo_list):
        qs, sis = [], []
        for i in range(self.num_lifting_layers):
            q, bits = self.ent_out_xo_list[i].rate(out_xo_list[i], self.training, self.bit_acc)
            sis.append(bits)
            qs.append(q)
        xe_q, si_xe = self.ent_out_xe.rate(out_xe, self.training, self.bit_acc)
        return si_xe, sis, xe_q, qs

    @torch.no_grad()
    def compress(self, out_xe, out_xo_list):
        """Real bitstreams (interleaved rANS, ``coding.py``): returns (streams, xe_q, qs); ``streams[0]`` codes the LL
        band, ``streams[1 + l]`` level l (finest first)."""
        from ... import coding
        was = self.training
        self.eval()
        si_xe, sis, xe_q, qs = self.forward(out_xe, out_xo_list)
        self.train(was)
        streams = [coding.encode_factorized(self.ent_out_xe, xe_q)]
        streams += [coding.encode_factorized(self.ent_out_xo_list[i], qs[i]) for i in range(self.num_lifting_layers)]
        return streams, xe_q, qs

    @torch.no_grad()
    def decompress(self, streams):
        from ... import coding
        xe_q = coding.decode_factorized(self.ent_out_xe, streams[0])
        qs = [coding.decode_factorized(self.ent_out_xo_list[i], streams[1 + i]) for i in range(self.num_lifting_layers)]
        return xe_q, qs


class DWTConditioned2EntropyLayerZTsepSubbands(nn.Module):
    """Causal spatial context + parent-subband (zero-tree) context (:233-372)."""

    def __init__(self, config):
        super().__init__()
        self.num_lifting_layers = config.dwtlevels
        self.scale_table = get_scale_table()
        self.config = config
        assert self.num_lifting_layers > 0
        self.clrch = config.clrch
        self.se, self.so = 1, 3
        self.ses, self.sos = _sos_ses(config)
        self.plc_list = nn.ModuleList()
        self.csc_list = nn.ModuleList()
        self.cgp_out_xo_list = nn.ModuleList()
        self.ent_out_xo_list = nn.ModuleList()
        self.scl_out_xo_list = nn.ParameterList()
        self.scb_out_xo_list = nn.ParameterList()
        for i in range(0, self.num_lifting_layers - 1, 1):
            inn_ch1 = self.sos[i + 1]
            out_ch1 = inn_ch1 * 81
            self.plc_list.append(nn.Sequential(nn.Conv2d(inn_ch1, out_ch1, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(),
                                               nn.Conv2d(out_ch1, out_ch1, kernel_size=3, stride=1, padding=1)))
            inn_ch2 = self.sos[i]
            out_ch2 = inn_ch2 * 81
            self.csc_list.append(MaskedConv2d(mask_type='A', in_channels=inn_ch2, out_channels=out_ch2, kernel_size=5,
                                              stride=1, padding=2, groups=inn_ch2))
            inn_ch = out_ch1 + out_ch2
            out_ch = self.sos[i] * 2
            self.cgp_out_xo_list.append(nn.Sequential(
                nn.Conv2d(inn_ch, inn_ch, kernel_size=1, stride=1, padding=0, groups=inn_ch1), nn.LeakyReLU(inplace=True),
                nn.Conv2d(inn_ch, inn_ch // 3, kernel_size=1, stride=1, padding=0, groups=inn_ch1), nn.LeakyReLU(inplace=True),
                nn.Conv2d(inn_ch // 3, inn_ch // 9, kernel_size=1, stride=1, padding=0, groups=inn_ch1), nn.LeakyReLU(inplace=True),
                nn.Conv2d(inn_ch // 9, out_ch, kernel_size=1, stride=1, padding=0, groups=inn_ch1)))
            self.ent_out_xo_list.append(GaussianConditional(scale_table=None, scale_bound=0.11))
        i = self.num_lifting_layers - 1
        self.csc_list.append(self._causal_chain(self.sos[i]))
        self.ent_out_xo_list.append(GaussianConditional(scale_table=None, scale_bound=0.11))
        self.csc_xe = self._causal_chain(self.ses[i])
        self.ent_out_xe = GaussianConditional(scale_table=None, scale_bound=0.11)
        self.bit_acc = None
        # "bf16": dense context CNNs (plc conv2, cgp layers 1-2) on the tcgen05 tensor cores;
        # "fp32": everything on the exact-fp32 SIMT kernels.  Only (sigma, mu) -- i.e. bpp -- differ.
        self.ctx_precision = _ctx_precision(config)
        self._tc_cache = [PackCache() for _ in range(self.num_lifting_layers)]
        self._chain_cache = {"xe": PackCache(), "xo": PackCache()}

    @staticmethod
    def _causal_chain(inn):
        o = inn * 81
        mk = lambda t, ci, co: MaskedConv2d(mask_type=t, in_channels=ci, out_channels=co, kernel_size=3, stride=1,
                                            padding=1, groups=inn)
        return nn.Sequential(mk('A', inn, o), nn.LeakyReLU(inplace=True), mk('B', o, o), nn.LeakyReLU(inplace=True),
                             mk('B', o, o // 3), nn.LeakyReLU(inplace=True), mk('B', o // 3, o // 9),
                             nn.LeakyReLU(inplace=True), mk('B', o // 9, inn * 2))

    def _chain_bits_input(self, key, seq, q):
        """(sigma, mu) map of a coarsest-level causal chain (:298-317: masked 3x3 convs inn -> 81 inn -> 81 inn -> 27 inn ->
        9 inn -> 2 inn, groups = inn).  Inference with ``ctx_precision: "bf16"``: the two dense layers (81 -> 81 and 81 -> 27
        per group, 96 % of the chain's MACs) run as grouped 9-tap tcgen05 implicit GEMMs on channels-last bf16 -- the mask
        is zeros in the packed weights -- and the thin layers around them stay on the exact-fp32 direct-conv kernel.  Like
        every other context CNN on this path the result only moves (sigma, mu), i.e. bpp."""
        convs = [m for m in seq if isinstance(m, MaskedConv2d)]
        G = convs[0].groups if convs else 0
        fits = CTX_TC_CHAIN and self.ctx_precision == "bf16" and len(convs) == 5 and 1 <= G <= 3 and \
            all(c.groups == G and tuple(c.kernel_size) == (3, 3) for c in convs) and \
            convs[0].out_channels // G <= 128 and convs[2].out_channels // G <= 32 and \
            not _autograd.needs_grad([q] + [p for c in convs for p in c.parameters()])
        if not fits:
            return _chain(seq, q)
        n1, n3 = convs[0].out_channels // G, convs[2].out_channels // G
        for c in convs:
            c.apply_mask()

        def build():
            w2, w3 = convs[1].weight.detach(), convs[2].weight.detach()
            l2 = [ops.pack_igemm_weight(w2[n1 * g:n1 * (g + 1)].contiguous(), npad=128, kpad=128) for g in range(G)]
            l3 = [ops.pack_igemm_weight(w3[n3 * g:n3 * (g + 1)].contiguous(), npad=32, kpad=128) for g in range(G)]
            return dict(l2=torch.stack(l2).contiguous(), l3=torch.stack(l3).contiguous())

        pk = self._chain_cache[key].get([convs[1].weight, convs[2].weight], build)
        koff = [[128 * g, 128 * g + 64] for g in range(G)]
        B, _, h, w = q.shape
        outs = []
        for b0 in range(0, B, 4 * CTX_BATCH_CHUNK):
            qb = q[b0:b0 + 4 * CTX_BATCH_CHUNK]
            n = qb.shape[0]
            t1 = torch.zeros(n, 128 * G, h, w, dtype=torch.float32, device=q.device)      # 81 live channels per 128-slot
            ops.conv2d(qb, convs[0].weight, convs[0].bias, groups=G, lrelu=True, out=t1, co_group=n1, co_stride=128, co_off=0)
            a = torch.empty(n, h, w, 128 * G, dtype=torch.bfloat16, device=q.device)
            ops.nchw_to_nhwc_bf16(t1, a, 0)
            del t1
            t2 = torch.empty(n, h, w, 128 * G, dtype=torch.bfloat16, device=q.device)
            ops.igemm_conv(a, pk["l2"], convs[1].bias, n1, lrelu=True, out_nhwc=t2, nhwc_coff=0, nhwc_gstride=128, koff=koff)
            del a
            t3 = ops.igemm_conv(t2, pk["l3"], convs[2].bias, n3, lrelu=True, koff=koff)        # fp32 NCHW (n, 27 G, h, w)
            del t2
            outs.append(_conv(convs[4], _conv(convs[3], t3, lrelu=True)))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    # ---- tensor-core path (default): BF16 tcgen05 implicit GEMMs, FP32 accumulation -------------
    def _tc_pack(self, i):
        """Packed BF16 weights of level i: plc conv2 (9 taps, 256x256), cgp layer 1 (per group
        K = [128-channel window of plc | 128-channel csc slot], N = 192) and layer 2 (K = 192, N = 64)."""
        plc, cgp = self.plc_list[i], self.cgp_out_xo_list[i]
        srcs = [plc[2].weight, cgp[0].weight, cgp[2].weight, plc[0].weight, self.csc_list[i].weight]

        def build():
            C = self.sos[i]
            nper = plc[0].out_channels // C                      # 81
            n1, n2 = cgp[0].out_channels // C, cgp[2].out_channels // C   # 162, 54
            starts = [(nper * g) // 64 * 64 for g in range(C)]
            if C > 3 or nper > 128 or any(nper * g + nper > starts[g] + 128 for g in range(C)) or \
                    C * nper > 256 or n1 > 192 or n2 > 64:
                raise NotImplementedError("tensor-core context path is laid out for clrch=1 (3 subbands x 81 channels)")
            dev = plc[2].weight.device
            wp_plc = ops.pack_igemm_weight(plc[2].weight, npad=256, kpad=256)
            w1 = cgp[0].weight.detach()[:, :, 0, 0]              # (C*n1, 2*nper)
            w2 = cgp[2].weight.detach()[:, :, 0, 0]              # (C*n2, n1)
            l1, l2, k1 = [], [], []
            for g in range(C):
                wg = torch.zeros(n1, 256, 1, 1, device=dev)
                off = nper * g - starts[g]
                wg[:, off:off + nper, 0, 0] = w1[n1 * g:n1 * (g + 1), :nper]
                wg[:, 128:128 + nper, 0, 0] = w1[n1 * g:n1 * (g + 1), nper:]
                l1.append(ops.pack_igemm_weight(wg, npad=192, kpad=256))
                l2.append(ops.pack_igemm_weight(w2[n2 * g:n2 * (g + 1)].reshape(n2, n1, 1, 1).contiguous(), npad=64, kpad=192))
                k1.append([starts[g], starts[g] + 64, 256 + 128 * g, 256 + 128 * g + 64])
            k2 = [[192 * g, 192 * g + 64, 192 * g + 128] for g in range(C)]
            pk = dict(plc=wp_plc, l1=torch.stack(l1).contiguous(), l2=torch.stack(l2).contiguous(), k1=k1, k2=k2,
                      nper=nper, n1=n1, n2=n2)
            # plc head (3x3 on the upsampled parent) and masked csc as 1-tap GEMMs over the split im2col rows of
            # ops.ctx_im2col: weights [W_hi | W_hi | W_lo] against [x_hi | x_lo | x_hi]
            csc = self.csc_list[i]
            if C == 3 and plc[0].in_channels == 3 and tuple(plc[0].kernel_size) == (3, 3) and csc.groups == 3 and \
                    tuple(csc.kernel_size) == (5, 5) and csc.in_channels == 3:
                csc.apply_mask()
                wh = torch.zeros(plc[0].out_channels, 128, 1, 1, device=dev)
                wh[:, :81, 0, 0] = ops.split_bf16_weight(plc[0].weight.detach().reshape(plc[0].out_channels, 27))
                pk["head"] = ops.pack_igemm_weight(wh, npad=256, kpad=128)
                wc = []
                for g in range(3):
                    wg = torch.zeros(nper, 64, 1, 1, device=dev)
                    wg[:, :36, 0, 0] = ops.split_bf16_weight(csc.weight.detach()[nper * g:nper * (g + 1), 0].reshape(nper, 25)[:, :12])
                    wc.append(ops.pack_igemm_weight(wg, npad=128, kpad=64))
                pk["csc"] = torch.stack(wc).contiguous()
            return pk

        return self._tc_cache[i].get(srcs, build)

    def _tc_fits(self, i):
        """The tensor-core layout holds 3 subbands x <= 128 channels (clrch = 1); other reference-valid shapes
        (clrch = 3: 9 subbands per level) take the exact-fp32 ``_level`` path."""
        C = self.sos[i]
        nper = self.plc_list[i][0].out_channels // C
        n1, n2 = self.cgp_out_xo_list[i][0].out_channels // C, self.cgp_out_xo_list[i][2].out_channels // C
        starts = [(nper * g) // 64 * 64 for g in range(C)]
        return C <= 3 and nper <= 128 and all(nper * g + nper <= starts[g] + 128 for g in range(C)) and \
            C * nper <= 256 and n1 <= 192 and n2 <= 64

    def _level_bits_tc(self, i, x, q, con, noise, acc):
        """Self-information (B,3,h,w) of conditioned level i on the tensor cores: head -> plc igemm and
        csc land in one NHWC bf16 tensor; cgp layers 1-2 are grouped 1x1 igemms; layers 3-4 are fused
        with the Gaussian rate (:352-365)."""
        B, C, h, w = x.shape
        pk = self._tc_pack(i)
        plc, cgp, csc = self.plc_list[i], self.cgp_out_xo_list[i], self.csc_list[i]
        csc.apply_mask()
        bits = torch.empty(B, C, h, w, dtype=torch.float32, device=x.device)
        for b0 in range(0, B, CTX_BATCH_CHUNK):
            b1 = min(B, b0 + CTX_BATCH_CHUNK)
            n = b1 - b0
            g_in = torch.empty(n, h, w, 256 + 128 * C, dtype=torch.bfloat16, device=x.device)
            if "head" in pk and CTX_GEMM_HEAD:
                # both small-Cin convs on the tensor cores: one im2col pass, then 1-tap igemm layers
                a = ops.ctx_im2col(con[b0:b1], q[b0:b1])
                t = torch.empty(n, h, w, 256, dtype=torch.bfloat16, device=x.device)
                ops.igemm_conv(a, pk["head"], plc[0].bias, plc[0].out_channels, lrelu=True, out_nhwc=t, koff=[[0, 64]])
                ops.igemm_conv(a, pk["csc"], csc.bias, pk["nper"], out_nhwc=g_in, nhwc_coff=256, nhwc_gstride=128,
                               koff=[[128], [192], [256]])
                del a
            else:
                t = ops.ctx_conv_nhwc(con[b0:b1], plc[0].weight, plc[0].bias, upsample2=True, lrelu=True, region=256)
                ops.ctx_conv_nhwc(q[b0:b1], csc.weight, csc.bias, groups=csc.groups, live_taps=12, out=g_in, coff=256,
                                  co_group=pk["nper"], co_gstride=128, region=128 * C)
            ops.igemm_conv(t, pk["plc"], plc[2].bias, plc[2].out_channels, out_nhwc=g_in, nhwc_coff=0)
            del t
            h1 = torch.empty(n, h, w, 192 * C, dtype=torch.bfloat16, device=x.device)
            ops.igemm_conv(g_in, pk["l1"], cgp[0].bias, pk["n1"], lrelu=True, out_nhwc=h1, nhwc_gstride=192, koff=pk["k1"])
            del g_in
            nz = noise[b0:b1] if noise is not None else None
            if CTX_FUSED_TAIL and pk["n2"] <= 64 and cgp[6].weight.shape[1] <= 20:
                # layers 2-4 + rate in one launch: the 54-channel map stays in tensor memory / registers
                bits[b0:b1] = ops.igemm_cgp_tail(h1, pk["l2"], cgp[2].bias, pk["n2"], pk["k2"], cgp[4].weight, cgp[4].bias,
                                                 cgp[6].weight, cgp[6].bias, x[b0:b1], nz, acc=acc)
                del h1
            else:
                h2 = ops.igemm_conv(h1, pk["l2"], cgp[2].bias, pk["n2"], lrelu=True, koff=pk["k2"])
                del h1
                bits[b0:b1] = ops.cgp_tail_rate(h2, cgp[4].weight, cgp[4].bias, cgp[6].weight, cgp[6].bias, x[b0:b1], nz, acc=acc)
                del h2
        return bits

    def _level_grad(self, i, x, q, con):
        """Differentiable form of ``_level`` (training): same kernels in the forward pass, the concat is a
        torch op instead of a write pattern, every conv recomputes through torch in backward."""
        plc, cgp, csc = self.plc_list[i], self.cgp_out_xo_list[i], self.csc_list[i]
        c = _conv(csc, q)
        t = _conv(plc[0], con, lrelu=True, upsample2=True)
        pl = _conv(plc[2], t)
        p0, p1, p2 = pl.chunk(3, dim=1)
        c0, c1, c2 = c.chunk(3, dim=1)
        return _chain(cgp, torch.cat((p0, c0, p1, c1, p2, c2), dim=1))

    def _level(self, i, x, q, con):
        """sigma/mu maps (B,6,h,w) of conditioned level i from the quantised child ``q`` and the
        half-resolution quantised parent ``con`` (:352-362) -- exact-fp32 SIMT path
        (``ctx_precision = "fp32"``)."""
        B, C, h, w = x.shape
        ms = torch.empty(B, 2 * C, h, w, dtype=torch.float32, device=x.device)
        plc, cgp, csc = self.plc_list[i], self.cgp_out_xo_list[i], self.csc_list[i]
        # the reference chunks plc / csc 3-way whatever clrch is (:357-359); with clrch = 1 a chunk = one subband's 81 channels
        nper = plc[2].out_channels // 3
        if csc.out_channels != plc[2].out_channels:
            raise ValueError("conditioned2ZT: plc and csc widths differ (parent and child levels need equal channel counts)")
        for b0 in range(0, B, CTX_BATCH_CHUNK):
            b1 = min(B, b0 + CTX_BATCH_CHUNK)
            cat = torch.empty(b1 - b0, 2 * nper * 3, h, w, dtype=torch.float32, device=x.device)
            # (plc0, csc0, plc1, csc1, plc2, csc2) channel order of :357-359 as write patterns
            csc(q[b0:b1], out=cat, co_group=nper, co_stride=2 * nper, co_off=nper)
            t = ops.conv2d(con[b0:b1], plc[0].weight, plc[0].bias, lrelu=True, upsample2=True)
            ops.conv2d(t, plc[2].weight, plc[2].bias, out=cat, co_group=nper, co_stride=2 * nper, co_off=0)
            del t
            ms[b0:b1] = _chain(cgp, cat)
            del cat
        return ms

    def forward(self, out_xe, out_xo_list):
        L = self.num_lifting_layers
        mode = "noise" if self.training else "dequantize"
        acc = self.bit_acc
        xe_q = self.ent_out_xe.quantize(out_xe, mode)
        si_xe = self.ent_out_xe.bits(out_xe, self._chain_bits_input("xe", self.csc_xe, xe_q), self.training, acc=acc)
        qs, sis = [], []
        i = L - 1
        q = self.ent_out_xo_list[i].quantize(out_xo_list[i], mode)
        sis.append(self.ent_out_xo_list[i].bits(out_xo_list[i], self._chain_bits_input("xo", self.csc_list[i], q), self.training,
                                                acc=acc))
        qs.append(q)
        con = q
        for i in range(L - 2, -1, -1):
            q = self.ent_out_xo_list[i].quantize(out_xo_list[i], mode)
            lvl_params = list(self.plc_list[i].parameters()) + list(self.cgp_out_xo_list[i].parameters()) + \
                list(self.csc_list[i].parameters())
            if _autograd.needs_grad([out_xo_list[i], q, con] + lvl_params):
                # a gradient is needed through this level: by its own parameters, or -- frozen entropy model, trainable
                # transform -- by the subbands alone (the rate term must reach the lifting CNNs either way)
                ms = self._level_grad(i, out_xo_list[i], q, con)
                sis.append(self.ent_out_xo_list[i].bits(out_xo_list[i], ms, self.training, acc=acc))
            elif self.ctx_precision == "bf16" and self._tc_fits(i):
                noise = compat.draw_noise(out_xo_list[i]) if self.training else None
                sis.append(self._level_bits_tc(i, out_xo_list[i].contiguous(), q, con, noise, acc))
            else:
                ms = self._level(i, out_xo_list[i], q, con)
                sis.append(self.ent_out_xo_list[i].bits(out_xo_list[i], ms, self.training, acc=acc))
            qs.append(q)
            con = q
        qs.reverse()
        sis.reverse()
        return si_xe, sis, xe_q, qs


    def test(self, out_xe, out_xo_list):
        raise NotImplementedError("serial coder (test/compress_ar/decompress_ar) is out of scope")


class DWTConditioned2EntropyLayerZTBlock(nn.Module):
    """Parent + 2x2 polyphase block conditioning, phases ee -> eo -> oe -> oo (:558-757)."""

    def __init__(self, config):
        super().__init__()
        self.num_lifting_layers = config.dwtlevels
        self.dwtLevels = config.dwtlevels
        assert self.num_lifting_layers > 0
        self.clrch = config.clrch
        self.multiplier = 8
        self.se = 1
        self.so = 3
        self.ses, self.sos = _sos_ses(config)
        names = [f"dep_{k}_list_{s}" for s in ("mu", "sigma") for k in (1, 2, 3, 4)]
        for n in names:
            setattr(self, n, nn.ModuleList())
        self.cgp_out_xo_list = nn.ModuleList()
        self.ent_out_xo_list = nn.ModuleList()
        self.scl_out_xo_list = nn.ParameterList()
        self.scb_out_xo_list = nn.ParameterList()
        hid = 32

        def net(cin):
            return nn.Sequential(
                nn.Conv2d(cin, hid, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(inplace=True),
                nn.Conv2d(hid, hid, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(inplace=True),
                nn.Conv2d(hid, hid, kernel_size=1, stride=1, padding=0), nn.LeakyReLU(inplace=True),
                nn.Conv2d(hid, hid, kernel_size=1, stride=1, padding=0), nn.LeakyReLU(inplace=True),
                nn.Conv2d(hid, 1, kernel_size=1, stride=1, padding=0))

        for i in range(0, self.num_lifting_layers - 1, 1):
            for j in range(3):
                self.ent_out_xo_list.append(GaussianConditional(scale_table=None, scale_bound=0.11))
                self.scl_out_xo_list.append(nn.Parameter(nn.init.constant_(torch.empty(1, self.sos[i], 1, 1), i * 1.0 + 1.0)))
                self.scb_out_xo_list.append(nn.Parameter(nn.init.constant_(torch.empty(1, self.sos[i], 1, 1), 1.0)))
                # construction order of the reference: mu nets 1..4, then sigma nets 1..4
                for s in ("mu", "sigma"):
                    for k in (1, 2, 3, 4):
                        getattr(self, f"dep_{k}_list_{s}").append(net(k))
        self.ent_out_xo_list.append(GaussianConditional(scale_table=None, scale_bound=0.11))
        self.gaussian_conditional = GaussianConditional(None)
        self.ent_out_xe = EntropyBottleneck(channels=1)
        self.ent_out_xo = EntropyBottleneck(channels=3)
        self.bit_acc = None

    _SLOTS = [(0, 0), (0, 1), (1, 0), (1, 1)]     # phases ee, eo, oe, oo

    def _phase_ms(self, n, k, dep):
        """(sigma, mu) of phase k (1..4) of conditioned subband n from its dependencies ``dep`` (B,k,h,w) -> (B,2,h,w)."""
        mu = _dep_net(getattr(self, f"dep_{k}_list_mu")[n], dep)
        sg = _dep_net(getattr(self, f"dep_{k}_list_sigma")[n], dep)
        return torch.cat((sg, mu), dim=1)

    def forward(self, out_xe, out_xo_list, keep_ms=None):
        """``keep_ms``: optional list that receives, per conditioned level (coarse to fine) and subband, the full-resolution
        (sigma, mu) tensor (B,2,H,W)."""
        L = self.dwtLevels
        acc = self.bit_acc
        mode = "noise" if self.training else "dequantize"
        xe_q, si_xe = self.ent_out_xe.rate(out_xe, self.training, acc)
        qs, sis = [], []
        q, si = self.ent_out_xo.rate(out_xo_list[L - 1], self.training, acc)
        qs.append(q)
        sis.append(si)
        con = q
        for i in range(0, L - 1):
            lvl = L - i - 2
            si_j, q_j, ms_j = [], [], []
            for j in range(3):
                gc = self.ent_out_xo_list[(L - i - 1) * 3 - j - 1]
                xin = out_xo_list[lvl][:, j:j + 1].contiguous()
                B, _, H, W = xin.shape
                ms = torch.empty(B, 2, H, W, dtype=torch.float32, device=xin.device)   # ch0 sigma, ch1 mu
                qq = gc.quantize(xin, mode)
                ee, eo, oe = qq[:, :, 0::2, 0::2], qq[:, :, 0::2, 1::2], qq[:, :, 1::2, 0::2]
                d1 = con[:, j:j + 1]
                n = j + i * 3
                deps = [d1, torch.cat((d1, ee), 1), torch.cat((d1, ee, eo), 1), torch.cat((d1, ee, eo, oe), 1)]
                for k, (dep, (ry, rx)) in enumerate(zip(deps, self._SLOTS), start=1):
                    ms[:, 1:2, ry::2, rx::2] = _dep_net(getattr(self, f"dep_{k}_list_mu")[n], dep)
                    ms[:, 0:1, ry::2, rx::2] = _dep_net(getattr(self, f"dep_{k}_list_sigma")[n], dep)
                si_j.append(gc.bits(xin, ms, self.training, acc=acc))
                q_j.append(qq)
                ms_j.append(ms)
            sis.append(torch.cat(si_j, dim=1))
            con = torch.cat(q_j, dim=1)
            qs.append(con)
            if keep_ms is not None:
                keep_ms.append(ms_j)
        qs.reverse()
        sis.reverse()
        return si_xe, sis, xe_q, qs

    @torch.no_grad()
    def compress(self, out_xe, out_xo_list):
        """Real bitstreams (interleaved rANS, ``coding.py``).  LL and the coarsest level under their factorized models;
        every conditioned subband as four streams, one per 2x2 phase (ee, eo, oe, oo), under the Gaussian of its phase
        discretised on the integer grid -- the layer decodes and conditions on plain round(x) (:719-724,754), while its
        rate estimate is taken at round(x - mu) + mu (:747-749), so the coded size follows the estimate only
        approximately.  Returns (streams, xe_q, qs); streams = [LL, coarsest, then per level (coarse to fine), subband
        and phase]."""
        from ... import coding
        was = self.training
        self.eval()
        mss = []
        si_xe, sis, xe_q, qs = self.forward(out_xe, out_xo_list, keep_ms=mss)
        self.train(was)
        L = self.dwtLevels
        streams = [coding.encode_factorized(self.ent_out_xe, xe_q), coding.encode_factorized(self.ent_out_xo, qs[L - 1])]
        for i in range(0, L - 1):
            q = qs[L - i - 2]
            for j in range(3):
                for ry, rx in self._SLOTS:
                    streams.append(coding.encode_gaussian(q[:, j:j + 1, ry::2, rx::2].contiguous(),
                                                          mss[i][j][:, :, ry::2, rx::2].contiguous(), integer_grid=True))
        return streams, xe_q, qs

    @torch.no_grad()
    def decompress(self, streams):
        """Coarse to fine; inside a subband the four phases in order, each conditioned on the parent and on the phases
        decoded before it."""
        from ... import coding
        L = self.dwtLevels
        xe_q = coding.decode_factorized(self.ent_out_xe, streams[0])
        con = coding.decode_factorized(self.ent_out_xo, streams[1])
        qs = [con]
        pos = 2
        for i in range(0, L - 1):
            q_j = []
            for j in range(3):
                d1 = con[:, j:j + 1]
                B, _, h, w = d1.shape
                qq = torch.empty(B, 1, 2 * h, 2 * w, dtype=torch.float32, device=d1.device)
                dep = d1
                for k, (ry, rx) in enumerate(self._SLOTS, start=1):
                    ph = coding.decode_gaussian(streams[pos], self._phase_ms(j + i * 3, k, dep.contiguous()))
                    pos += 1
                    qq[:, :, ry::2, rx::2] = ph
                    dep = torch.cat((dep, ph), 1)
                q_j.append(qq)
            con = torch.cat(q_j, dim=1)
            qs.append(con)
        qs.reverse()
        return xe_q, qs


class onlyEZWT(nn.Module):
    """Coarsest level + LL factorized, finer levels conditioned on the parent subband (:759-840)."""

    def __init__(self, config):
        super().__init__()
        self.num_lifting_layers = config.dwtlevels
        self.scale_table = get_scale_table()
        self.config = config
        assert self.num_lifting_layers > 0
        self.clrch = config.clrch
        self.se, self.so = 1, 3
        self.ses, self.sos = _sos_ses(config)
        self.plc_list = nn.ModuleList()
        self.ent_out_xo_list = nn.ModuleList()
        for i in range(0, self.num_lifting_layers - 1, 1):
            inn_ch1 = self.sos[i + 1]
            out_ch1 = inn_ch1 * 81
            self.plc_list.append(nn.Sequential(
                nn.Conv2d(inn_ch1, out_ch1, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(),
                nn.Conv2d(out_ch1, out_ch1, kernel_size=3, stride=1, padding=1), nn.LeakyReLU(),
                nn.Conv2d(out_ch1, 6, kernel_size=1, stride=1, padding=0)))
            self.ent_out_xo_list.append(GaussianConditional(scale_table=None, scale_bound=0.11))
        self.ent_out_xe = EntropyBottleneck(channels=1)
        self.ent_out_xo = EntropyBottleneck(channels=3)
        self.bit_acc = None
        self.ctx_precision = _ctx_precision(config)      # "fp32": FP32 FMA kernels only; anything else: 3xTF32 chain
        self._tc_cache = [PackCache() for _ in range(self.num_lifting_layers - 1)]

    def forward(self, out_xe, out_xo_list, keep_ms=None):
        """``keep_ms``: optional list that receives the (sigma, mu) tensor of every conditioned level (finest first)."""
        L = self.num_lifting_layers
        acc = self.bit_acc
        xe_q, si_xe = self.ent_out_xe.rate(out_xe, self.training, acc)
        qs, sis, mss = [], [], []
        q, si = self.ent_out_xo.rate(out_xo_list[L - 1], self.training, acc)
        qs.append(q)
        sis.append(si)
        con = q
        for i in range(L - 2, -1, -1):
            x = out_xo_list[i]
            ms = self._ms(i, con)
            bits, q = self.ent_out_xo_list[i].bits(x, ms, self.training, want_y=True, acc=acc)
            sis.append(bits)
            qs.append(q)
            mss.append(ms)
            con = q
        qs.reverse()
        sis.reverse()
        if keep_ms is not None:
            keep_ms.extend(reversed(mss))
        return si_xe, sis, xe_q, qs

    def _ms(self, i, con):
        """(sigma, mu) of level i from the dequantised parent level ``con`` (B,3,h/2,w/2) -> (B,6,h,w) (:826-831)."""
        plc = self.plc_list[i]
        B = con.shape[0]
        if _autograd.needs_grad([con] + list(plc.parameters())):
            return _chain(plc[2:], _conv(plc[0], con, lrelu=True, upsample2=True))   # differentiable, whole batch
        ms = torch.empty(B, 6, 2 * con.shape[2], 2 * con.shape[3], dtype=torch.float32, device=con.device)
        # fp32-level accuracy on purpose: this layer's mu is part of the *dequantised* output (round(x - mu) + mu, :832)
        # and so of the reconstruction (1e-4 tolerance); BF16 operands would move it by ~1e-3.  (cond2ZT returns plain
        # round(x): there mu only feeds the rate.)  Default: the 243 -> 243 conv (88 % of the layer's FLOPs) on the
        # 3xTF32 tensor-core chain of SubbandAutoEncoderBerk; ``ctx_precision: "fp32"`` keeps everything on the FP32 FMA
        # kernels.
        tc = self.ctx_precision != "fp32" and plc[2].weight.shape[0] <= 256
        if tc:
            pk = self._tc_cache[i].get([plc[0].weight, plc[0].bias, plc[2].weight, plc[2].bias], lambda: self._tc_pack(plc))
        for b0 in range(0, B, CTX_BATCH_CHUNK):
            b1 = min(B, b0 + CTX_BATCH_CHUNK)
            if tc:
                t = ops.conv2d(con[b0:b1], pk["w0"], pk["b0"], lrelu=True, upsample2=True)     # (b,256,H,W), pad channels 0
                _, sz = ops.nchw_to_nhwc_split(t, squares=False, want_y=False)
                del t
                y, _ = ops.igemm_tf32(sz, pk["wp2"], pk["b2"], pk["cpad"], epi=3)
                del sz
                ops.nhwc_lrelu_conv1(y, plc[4].weight, plc[4].bias, plc[4].weight.shape[1], out=ms[b0:b1])
                del y
                continue
            t = ops.conv2d(con[b0:b1], plc[0].weight, plc[0].bias, lrelu=True, upsample2=True)
            t = ops.conv2d(t, plc[2].weight, plc[2].bias, lrelu=True)
            ops.conv2d(t, plc[4].weight, plc[4].bias, out=ms[b0:b1])
            del t
        return ms

    @staticmethod
    def _tc_pack(plc):
        """Weights of the first two convs zero-padded to a multiple of 32 channels for the 3xTF32 implicit GEMM."""
        c = plc[2].weight.shape[0]
        cpad = (c + 31) // 32 * 32
        w0 = plc[0].weight.detach()
        w0p = torch.zeros(cpad, *w0.shape[1:], dtype=torch.float32, device=w0.device)
        w0p[:c] = w0
        b0p = torch.zeros(cpad, dtype=torch.float32, device=w0.device)
        b0p[:c] = plc[0].bias.detach()
        w2p = torch.zeros(cpad, cpad, 3, 3, dtype=torch.float32, device=w0.device)
        w2p[:c, :c] = plc[2].weight.detach()
        b2p = torch.zeros(cpad, dtype=torch.float32, device=w0.device)
        b2p[:c] = plc[2].bias.detach()
        return {"w0": w0p, "b0": b0p, "wp2": ops.pack_tf32_weight(w2p), "b2": b2p, "cpad": cpad}

    @torch.no_grad()
    def compress(self, out_xe, out_xo_list):
        """Real bitstreams (interleaved rANS, ``coding.py``): LL and the coarsest level under their factorized models,
        every finer level under the Gaussian conditioned on its decoded parent.  Returns (streams, xe_q, qs) with
        ``streams[0]`` = LL, ``streams[1 + l]`` = level l (finest first)."""
        from ... import coding
        L = self.num_lifting_layers
        was = self.training
        self.eval()
        mss = []
        si_xe, sis, xe_q, qs = self.forward(out_xe, out_xo_list, keep_ms=mss)
        self.train(was)
        streams = [None] * (L + 1)
        streams[0] = coding.encode_factorized(self.ent_out_xe, xe_q)
        streams[L] = coding.encode_factorized(self.ent_out_xo, qs[L - 1])
        for i in range(L - 2, -1, -1):
            streams[1 + i] = coding.encode_gaussian(qs[i], mss[i])
        return streams, xe_q, qs

    @torch.no_grad()
    def decompress(self, streams):
        """Coarse to fine: a level's (sigma, mu) come from the level decoded just before it."""
        from ... import coding
        L = self.num_lifting_layers
        xe_q = coding.decode_factorized(self.ent_out_xe, streams[0])
        qs = [None] * L
        qs[L - 1] = coding.decode_factorized(self.ent_out_xo, streams[L])
        for i in range(L - 2, -1, -1):
            qs[i] = coding.decode_gaussian(streams[1 + i], self._ms(i, qs[i + 1]))
        return xe_q, qs
