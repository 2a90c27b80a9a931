"""The tcgen05 LFA kernel (lfa_tc.cu) against the FP32 SIMT kernel (lfa.cu) on identical
inputs, and end-to-end through the model (the golden / port parity tests in
test_gpu_models.py run with the tensor-core path enabled by default)."""
import numpy as np
import pytest
import torch

from open3d_ml_b200 import _lib as L
from conftest import rel_err

pytestmark = pytest.mark.gpu


def make(d, B, N, seed, gain=1.0):
    g = torch.Generator().manual_seed(seed)
    h = d // 2
    coords = (torch.rand(B * N, 3, generator=g) * 10).cuda()
    nidx = torch.randint(0, N, (B, N, 16), generator=g).cuda()
    feat = torch.randn(B * N, h, generator=g).cuda()
    w10 = (torch.randn(10, h, generator=g) * 0.3).cuda()
    s10, t10 = (torch.rand(h, generator=g) + 0.5).cuda(), (torch.randn(h, generator=g) * 0.1).cuda()
    wl2 = (torch.randn(h, h, generator=g) / h ** 0.5)          # [out, in]
    s2, t2 = (torch.rand(h, generator=g) + 0.5).cuda(), (torch.randn(h, generator=g) * 0.1).cuda()
    ws = (torch.randn(d, d, generator=g) / d ** 0.5) * gain    # [out, in]; gain spreads the scores
    bs = torch.randn(d, generator=g).cuda()
    return coords, nidx, feat, w10, s10, t10, wl2, s2, t2, ws, bs


@pytest.mark.parametrize("d", [16, 32, 64, 128, 256])
@pytest.mark.parametrize("stage", [1, 2])
@pytest.mark.parametrize("B,N,gain", [(1, 8, 1.0), (2, 1000, 1.0), (3, 2817, 1.0), (2, 500, 40.0)])
def test_lfa_tc_matches_simt(d, stage, B, N, gain):
    """gain = 40 makes the scores of neighbouring points differ by hundreds: the softmax max must be
    taken per point (a wrong group maximum only shows up once exp() underflows)."""
    coords, nidx, feat, w10, s10, t10, wl2, s2, t2, ws, bs = make(d, B, N, 7 * d + stage, gain)
    ref = torch.full((B * N, d), float("nan")).cuda()
    out = torch.full((B * N, d), float("nan")).cuda()
    wl2t, wst = wl2.t().contiguous().cuda(), ws.t().contiguous().cuda()
    L.check(L.lib().o3dml_randla_lfa_pool(stage, d, L.ptr(coords), L.ptr(nidx), 1, 16, L.ptr(feat), B, N,
                                          L.ptr(w10), L.ptr(s10), L.ptr(t10), L.ptr(wl2t), L.ptr(s2), L.ptr(t2),
                                          L.ptr(wst), L.ptr(bs), L.ptr(ref), L.stream()))
    img_l2 = L.pack_operand_image(wl2) if d >= 32 else None
    img_s = L.pack_operand_image(ws)
    L.check(L.lib().o3dml_randla_lfa_pool_tc(stage, d, L.ptr(coords), L.ptr(nidx), 1, 16, L.ptr(feat), B, N,
                                             L.ptr(w10), L.ptr(s10), L.ptr(t10), L.ptr(img_l2), L.ptr(wl2t),
                                             L.ptr(s2), L.ptr(t2), L.ptr(img_s), L.ptr(out), L.stream()))
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()
    assert rel_err(out, ref) < (2e-5 if gain == 1.0 else 2e-3), rel_err(out, ref)
