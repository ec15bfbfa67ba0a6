"""Shifted-window operand addressing probe (csrc/probe/halo_probe.cu): which (halo row pitch, base_offset) combinations give
the exact 3x3-tap A operand out of ONE TMA-loaded halo tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.ops import ptr, stream_ptr
lib = _lib.load_probe()
torch.manual_seed(0)
a = torch.randint(-8, 9, (18, 16, 32), device="cuda:0").float()
b = torch.randint(-8, 9, (32, 32), device="cuda:0").float()
for pitch in (10, 12, 16):
    for bo in (0, 1):
        res = []
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                d = torch.full((128, 32), float("nan"), device="cuda:0")
                _lib.check_probe(lib.ll_halo_probe(ptr(a), ptr(b), ptr(d), pitch, dy, dx, bo, stream_ptr()))
                torch.cuda.synchronize()
                win = a[1 + dy:17 + dy, 1 + dx:9 + dx, :].reshape(128, 32)
                ref = win @ b.t()
                res.append(int((d != ref).sum()))
        print(f"pitch {pitch} base_offset_mode {bo}: mismatching outputs per tap (of 4096) {res}", flush=True)
