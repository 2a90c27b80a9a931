import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
def emu(a, b, ftz):
    def split(x):
        h1 = x.astype(np.float16); h2 = (x - h1.astype(np.float32)).astype(np.float16)
        if ftz:
            tiny = np.float16(6.1035e-05)
            h1 = np.where(np.abs(h1) < tiny, np.float16(0), h1); h2 = np.where(np.abs(h2) < tiny, np.float16(0), h2)
        return h1.astype(np.float64), h2.astype(np.float64)
    a1, a2 = split(a); b1, b2 = split(b)
    d = a1 @ b1.T + a1 @ b2.T + a2 @ b1.T
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    return np.abs(d - ref).max() / np.abs(ref).max()
for scale in (1.0, 1e-1, 1e-2, 1e-3, 1e-4):
    g = torch.Generator().manual_seed(0)
    a = (torch.randn(128, 128, generator=g) * scale)
    b = (torch.randn(128, 128, generator=g) / 128 ** 0.5)
    line = "scale %g  emu(no ftz) %.2e  emu(ftz) %.2e" % (scale, emu(a.numpy(), b.numpy(), False), emu(a.numpy(), b.numpy(), True))
    if torch.cuda.is_available():
        from open3d_ml_b200 import _lib as L
        d = torch.empty(128, 128).cuda(); ac, bc = a.cuda(), b.cuda()
        L.check(L.lib().o3dml_tc_gemm_test(L.ptr(ac), L.ptr(bc), L.ptr(d), 128, 128, 3, L.stream()))
        ref = ac.double() @ bc.double().t()
        line += "  gpu %.2e" % float((d.double() - ref).abs().max() / ref.abs().max())
    print(line)
