import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from test_gpu_tc_probe import _probe, DEV
torch.manual_seed(0)
kb = 4
A = torch.randn(128, 8*kb, device=DEV); B = torch.randn(8*kb, 64, device=DEV)
d, _ = _probe(A, B, 5)            # reference readout through 32x32b
raw, _ = _probe(A, B, 5 | 16)     # first 4096 floats = raw 16x256b.x4 registers
raw = raw.flatten()[:4096].view(4, 2, 32, 16).cpu(); d = d.cpu()
ok = True
for w in range(4):
    for h in range(2):
        for t in range(32):
            for j in range(4):
                for half in range(2):
                    for e in range(2):
                        lane = 32 * w + 16 * h + t // 4 + 8 * half
                        col = 3 + 8 * j + 2 * (t % 4) + e
                        got = raw[w, h, t, 4 * j + 2 * half + e].item()
                        exp = d[lane, col].item()
                        if got != exp:
                            if ok: print("first mismatch", w, h, t, j, half, e, got, exp)
                            ok = False
print("16x256b.x4 mapping as assumed:", ok)
