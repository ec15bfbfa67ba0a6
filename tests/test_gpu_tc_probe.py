"""GPU test of the tensor-core building block of the learned-lifting kernel (``-m gpu``): A resident in
tensor memory, B in MN-major SWIZZLE_128B shared-memory atoms, tcgen05.mma kind::tf32, against a
float64 matmul.  Tolerances: plain TF32 ~ 2^-11 per operand; the 3xTF32 split must be as good as an
fp32 FMA chain (a few 1e-7 of the output scale)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _probe(A, B, split, reps=1):
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib
    lib = _lib.load_probe()
    kb = A.shape[1] // 8
    D = torch.empty(128, 64, device=DEV)
    cyc = torch.zeros(1, dtype=torch.int64, device=DEV)
    _lib.check_probe(lib.ll_tc_tf32_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), kb, split, reps, cyc.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return D, int(cyc.item())


@pytest.mark.parametrize("kblocks", [1, 2, 10, 20])
def test_tf32_ts_mma_layouts_and_split_accuracy(kblocks):
    torch.manual_seed(kblocks)
    A = torch.randn(128, 8 * kblocks, device=DEV)
    B = torch.randn(8 * kblocks, 64, device=DEV)
    ref = A.double() @ B.double()
    scale = ref.abs().max().item()
    d1, _ = _probe(A, B, 4)
    assert (d1.double() - ref).abs().max().item() <= 4e-3 * scale          # plain TF32: layouts are right
    d3, _ = _probe(A, B, 5)
    err3 = (d3.double() - ref).abs().max().item() / scale
    ref32 = A @ B                                                          # fp32 (TF32 disabled by default for matmul)
    err32 = (ref32.double() - ref).abs().max().item() / scale
    assert err3 <= 2e-6, (err3, err32)                                     # 3xTF32: fp32-level
    print(f"kblocks={kblocks}: tf32 err {(d1.double() - ref).abs().max().item() / scale:.2e}, 3xTF32 err {err3:.2e}, fp32 err {err32:.2e}")


def test_tf32_mma_issue_rate():
    A = torch.randn(128, 160, device=DEV)
    B = torch.randn(160, 64, device=DEV)
    _, c1 = _probe(A, B, 4, reps=50)
    _, c3 = _probe(A, B, 5, reps=50)
    print(f"20 k-blocks x 50 chains: {c1 / 20:.1f} (1 MMA/block) / {c3 / 60:.1f} (3xTF32) cycles per M128xN64xK8 MMA")
    assert c1 > 0 and c3 > 0


def test_measured_peaks_are_plausible():
    """The FFMA2 and dense TF32 probes (libll_probe.so) that bench.py uses as roofline denominators."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    fma = ops.fma_peak_tflops()
    tf32 = ops.tf32_peak_tflops()
    print(f"measured peaks: FFMA2 {fma:.1f} TFLOP/s, dense TF32 (tcgen05 SS, M128xN256xK8) {tf32:.1f} TFLOP/s")
    assert 40 < fma < 100
    assert 400 < tf32 < 1300
