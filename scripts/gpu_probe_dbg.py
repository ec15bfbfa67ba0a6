import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from test_gpu_tc_probe import _probe, DEV
torch.manual_seed(0)
for kb in (20,):
    A = torch.randn(128, 8*kb, device=DEV); B = torch.randn(8*kb, 64, device=DEV)
    ref = A.double() @ B.double(); sc = ref.abs().max().item()
    for mode in (4, 5, 13):
        d, c = _probe(A, B, mode)
        e = (d.double()-ref).abs().max().item()/sc
        d2, c2 = _probe(A, B, mode, reps=20)
        nm = kb * (3 if mode & 1 else 1)
        print(f"kb={kb} mode={mode} relerr={e:.3e} cycles(1 chain)={c} cycles/MMA at 20 chains={(c2*20)/(20*nm):.1f}")
