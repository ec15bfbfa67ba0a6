import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import SubbandAutoEncoderBerk
torch.manual_seed(0)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    for ic, shape in ((3, (8, 3, 256, 384)), (1, (16, 1, 32, 48))):
        ae = SubbandAutoEncoderBerk(ic).to("cuda:0").eval()
        x = torch.randn(*shape, device="cuda:0")
        y = ae.encode(x)
        print(f"in_ch={ic} {shape}: encode {t(lambda: ae.encode(x)):.2f} ms, decode {t(lambda: ae.decode(y)):.2f} ms")
