"""Bitstreams for the entropy layers that decode a whole subband at a time.

The reference's only real coder is the serial per-coefficient loop of ``conditioned2ZTsepSubbands``
(LiftingBasedDWT_net.py:374-556); ``factorized`` (:182-231) and ``onlyEZWT`` (:759-840) have no ``test`` method there at
all, although their contexts only depend on levels that are already decoded.  Here their quantised subbands are coded
with the interleaved rANS kernels of ``csrc/rans.cu`` under the model's own distributions, so the byte count is the
estimated rate plus the flush of the streams, and ``decode(encode(q)) == q`` bit for bit.
"""
import struct
from dataclasses import dataclass

import torch

from . import ops

_MAGIC = b"LLR1"


@dataclass
class SubbandBitstream:
    """One coded (B,C,H,W) tensor: ``B*S`` interleaved rANS streams, stream ``b*S+s`` = ``counts[b*S+s]`` words."""
    mode: int
    shape: tuple
    S: int
    counts: torch.Tensor      # int32 (B*S,)
    words: torch.Tensor       # int16 storage of the uint16 words

    def nbytes(self):
        """Size of :meth:`to_bytes`: payload + the per-stream length table + a 28-byte header."""
        return 2 * self.words.numel() + 4 * self.counts.numel() + 28

    def nbytes_per_image(self):
        """Payload + length-table bytes of every image (the streams of an image are separable)."""
        B = self.shape[0]
        return (2 * self.counts.view(B, self.S).to(torch.int64).sum(dim=1) + 4 * self.S).cpu()

    def to_bytes(self):
        B, C, H, W = self.shape
        head = _MAGIC + struct.pack("<6i", self.mode, B, C, H, W, self.S)
        return head + self.counts.cpu().numpy().tobytes() + self.words.cpu().numpy().tobytes()

    @classmethod
    def from_bytes(cls, data, device):
        if data[:4] != _MAGIC:
            raise ValueError("not a subband bitstream")
        mode, B, C, H, W, S = struct.unpack("<6i", data[4:28])
        n = B * S
        counts = torch.frombuffer(bytearray(data[28:28 + 4 * n]), dtype=torch.int32).to(device)
        words = torch.frombuffer(bytearray(data[28 + 4 * n:]), dtype=torch.int16).to(device)
        if int(counts.sum()) != words.numel():
            raise ValueError("truncated subband bitstream")
        return cls(mode, (B, C, H, W), S, counts, words)


def encode_factorized(eb, q, streams=None):
    """``q`` = dequantised output of ``EntropyBottleneck`` (round(x - median) + median)."""
    words, counts, S = ops.rans_encode(ops.RANS_EB, q, eb._blob(), streams)
    return SubbandBitstream(ops.RANS_EB, tuple(q.shape), S, counts, words)


def decode_factorized(eb, bs):
    return ops.rans_decode(ops.RANS_EB, bs.words, bs.counts, eb._blob(), bs.shape, bs.S)


def encode_gaussian(q, ms, streams=None, integer_grid=False):
    """``ms`` (B,2C,H,W): sigma, mu.  ``q`` = dequantised output of ``GaussianConditional`` (round(x - mu) + mu), or,
    with ``integer_grid``, the plain round(x) the ZTBlock layer decodes (Gaussian discretised on the integers)."""
    mode = ops.RANS_GAUSS_GRID if integer_grid else ops.RANS_GAUSS
    words, counts, S = ops.rans_encode(mode, q, ms, streams)
    return SubbandBitstream(mode, tuple(q.shape), S, counts, words)


def decode_gaussian(bs, ms):
    return ops.rans_decode(bs.mode, bs.words, bs.counts, ms, bs.shape, bs.S)
