"""``MaskedConv2d`` -- PixelCNN-style causal conv (reference: graphs/layers/masked_conv2d.py:5-21).

Same parameters and ``mask`` buffer; like the reference, ``forward`` multiplies the mask into
``weight.data`` in place before convolving.  The convolution is the sm_100a direct-conv kernel.
"""
from torch import nn

from ... import _autograd, _torch_ref, ops


class MaskedConv2d(nn.Conv2d):
    def __init__(self, mask_type, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert mask_type in ('A', 'B')
        self.register_buffer('mask', self.weight.data.clone())
        _, _, kH, kW = self.weight.size()
        self.mask.fill_(1)
        if kW > 1:
            self.mask[:, :, kH // 2, kW // 2 + (mask_type == 'B'):] = 0
        elif kW == 1 and mask_type == 'A':
            self.mask[:, :, kH // 2, kW // 2 + (mask_type == 'B'):] = 0
        if kH > 1:
            self.mask[:, :, kH // 2 + 1:] = 0

    def apply_mask(self):
        self.weight.data *= self.mask

    def forward(self, x, lrelu=False, **remap):
        self.apply_mask()
        if _autograd.needs_grad([x, self.weight, self.bias]) and not remap:
            g = self.groups
            return _autograd.run(lambda x, w, b: ops.conv2d(x, w, b, groups=g, lrelu=lrelu),
                                 lambda x, w, b: _torch_ref.conv2d(x, w, b, g, lrelu, False), [x, self.weight, self.bias])
        return ops.conv2d(x, self.weight, self.bias, groups=self.groups, lrelu=lrelu, **remap)
