"""Per-CTA timeline of the PointPillars neck / head layers (debug build, O3DML_DEBUG_TIMING=1)."""
import sys, ctypes; sys.path.insert(0, '.')
import torch, numpy as np
from open3d_ml_b200 import _lib as L

h = ctypes.CDLL(L.LIB_PATH)


def report(name, fn, ncta):
    big = torch.empty(64 << 20, device="cuda")
    for _ in range(3):
        fn()
    big.zero_(); big.zero_(); big.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    REP = 50
    ev[0].record()
    for _ in range(REP):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 7000)()
    h.o3dml_gt_debug_read(buf, 7000)
    a = np.array(buf[:]); t0 = a[4000]
    n = min(1000, ncta)
    g0, g1 = a[5000:5000 + n], a[6000:6000 + n]
    print("%-28s %7.1f us/launch | CTA0 cycles: range %6d  mainloop done %6d  end %6d | CTA ns: min %6d max %6d, first-wave span %6d" % (
        name, ev[0].elapsed_time(ev[1]) * 1e3 / REP, a[4001] - t0, a[4002] - t0, a[4003] - t0,
        (g1 - g0).min(), (g1 - g0).max(), g1.max() - g0.min()))


def deconv(H, W, C, s, co):
    x = torch.randn(1, H, W, C).cuda()
    w = torch.randn(C, co, s, s) / C ** 0.5
    pw = L.pack_linear(w.permute(0, 2, 3, 1).reshape(C, s * s * co))
    sc, sh = torch.ones(s * s * co).cuda(), torch.zeros(s * s * co).cuda()
    out = torch.empty(1, H * s, W * s, 384).cuda()
    return lambda: L.check(L.lib().o3dml_deconv_nhwc_tc(L.ptr(x), 1, H, W, C, s, L.ptr(pw.img), pw.k_pad, pw.n_pad,
                                                        L.ptr(sc), L.ptr(sh), 1, 0.0, L.ptr(out), 384, co, L.stream())), (x, out, sc, sh, pw)


def linear(rows, K, N, nchw=False):
    x = torch.randn(rows, K).cuda(); w = torch.randn(N, K) / K ** 0.5
    pw = L.pack_linear(w.t().contiguous()); t = torch.zeros(N).cuda()
    out = torch.empty(rows, N).cuda()
    return lambda: L.linear([L.make_src(x)], pw, out, None, t, act=None, num_rows=rows, out_channels=N,
                            out_nchw_plane=rows if nchw else 0), (x, out, t, pw)


def conv(H, W, C, co, stride):
    x = torch.randn(1, H, W, C).cuda(); w = torch.randn(co, C, 3, 3) / (9 * C) ** 0.5
    pw = L.pack_linear(w.permute(2, 3, 1, 0).reshape(9 * C, co)); s = torch.ones(co).cuda(); t = torch.zeros(co).cuda()
    OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    o = torch.empty(1, OH, OW, co).cuda()
    return lambda: L.check(L.lib().o3dml_conv3x3_nhwc_tc(L.ptr(x), 1, H, W, C, stride, L.ptr(pw.img), pw.k_pad, pw.n_pad,
                                                         L.ptr(s), L.ptr(t), 1, 0.0, L.ptr(o), co, L.stream())), (x, o, s, t, pw)


L.TC_MIN_K = 8
for name, (fn, keep), ncta in [
        ("deconv1 248x216 64->128 s1", deconv(248, 216, 64, 1, 128), 419),
        ("deconv2 124x108 128->128 s2", deconv(124, 108, 128, 2, 128), 420),
        ("deconv3 62x54 256->128 s4", deconv(62, 54, 256, 4, 128), 432),
        ("linear 53568x64->128", linear(53568, 64, 128), 419),
        ("head 53568x384->72 nchw", linear(53568, 384, 72, True), 419),
        ("head 53568x384->72 rows", linear(53568, 384, 72, False), 419),
        ("conv b1 248x216 64->64", conv(248, 216, 64, 64, 1), 419),
        ("conv b3 62x54 256->256", conv(62, 54, 256, 256, 1), 54)]:
    report(name, fn, ncta)
