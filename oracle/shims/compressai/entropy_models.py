"""Module-shaped compressai 1.2.1 entropy models over oracle.thirdparty maths.
Registration order of parameters/buffers follows compressai so that seeded
construction and ``state_dict`` keys match what the reference would produce."""
import torch
import torch.nn as nn

from compressai.ops.bound_ops import LowerBound
from oracle import thirdparty as tp


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def quantize(self, inputs, mode, means=None):
        return tp.quantize(inputs, mode, means)

    @staticmethod
    def dequantize(inputs, means=None, dtype=torch.float):
        out = inputs.type(dtype)
        if means is not None:
            out = out + means
        return out


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self.tail_mass = float(tail_mass)
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    def _standardized_cumulative(self, inputs):
        return tp.standardized_cumulative(inputs)

    def _likelihood(self, inputs, scales, means=None):
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, "noise" if training else "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood


class EntropyBottleneck(EntropyModel):
    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        for name, value in tp.eb_init_params(self.channels, self.filters, self.init_scale).items():
            self.register_parameter(name, nn.Parameter(value))
        self.register_buffer("target", tp.eb_target(self.tail_mass))

    def _params(self):
        return dict(self.named_parameters(recurse=False))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach()

    def loss(self):
        return tp.eb_loss(self._params())

    def forward(self, x, training=None):
        if training is None:
            training = self.training
        return tp.entropy_bottleneck_forward(self._params(), x, training,
                                             1e-9 if self.use_likelihood_bound else 0.0)
