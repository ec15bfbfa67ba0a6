import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from test_gpu_tc_probe import _probe, DEV
# usage: gpu_probe_dbg2.py <column offset>: checks the assumed register mapping of tcgen05.ld.16x32bx2.x32
coff = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0)
kb = 4
A = torch.randn(128, 8*kb, device=DEV); B = torch.randn(8*kb, 64, device=DEV)
d, _ = _probe(A, B, 5)            # reference readout through 32x32b
raw, _ = _probe(A, B, 5 | 32 | (coff << 8))
raw = raw.flatten().view(4, 2, 32, 32).cpu(); d = d.cpu()
ok = True; checked = 0
for w in range(4):
    for h in range(2):
        for t in range(32):
            for j in range(32):
                lane = 32 * w + 16 * h + (t % 16)
                col = coff + 32 * (t // 16) + j
                if col >= 64: continue
                checked += 1
                if raw[w, h, t, j].item() != d[lane, col].item():
                    if ok: print("first mismatch", w, h, t, j, raw[w, h, t, j].item(), d[lane, col].item())
                    ok = False
print(f"16x32bx2.x32 mapping at column offset {coff}: {ok} ({checked} values)")
