"""Per-role wait / work cycles of gemm_tc_kernel CTA (0,0) at FULL clocks: the layer is launched a few hundred
times back to back (the SM clock ramps over tens of ms) and the timeline of the last launch is read.
Needs a build with O3DML_DEBUG_TIMING=1."""
import sys, ctypes; sys.path.insert(0, '.')
import torch, numpy as np
from open3d_ml_b200 import _lib as L
h = ctypes.CDLL(L.LIB_PATH)


def run(name, fn, nsl):
    for _ in range(400):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(100):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 7000)()
    h.o3dml_gt_debug_read(buf, 7000)
    a = np.array(buf[:]); t0 = a[4000]
    n = min(nsl, 200)
    sl = a[:4 * n].reshape(n, 4); mm = a[2000:2000 + 3 * n].reshape(n, 3)
    cyc = a[4003] - t0
    print("%-26s %6.1f us/launch | CTA0 %6d cycles | per slice: %5.0f cycles" % (
        name, ev[0].elapsed_time(ev[1]) * 10, cyc, (a[4002] - t0) / nsl))
    print("   CTA 0 phases       : prologue (start -> first slice landed) %6d | main loop %7d | epilogue (last chunk done -> exit) %6d cycles" % (
        a[1] - t0, a[4002] - a[1], a[4003] - a[4002]))
    k = slice(2, n - 1)
    print("   converter thread 0: wait_full %6.0f  split+fence %6.0f   (cycles, mean over slices)" % (
        (sl[k, 1] - sl[k, 0]).mean(), (sl[k, 2] - sl[k, 1]).mean()))
    print("   MMA warp          : wait_conv %6.0f  issue %5.0f  | slice period %6.0f" % (
        (mm[k, 1] - mm[k, 0]).mean(), (mm[k, 2] - mm[k, 1]).mean(), np.diff(mm[k, 0]).mean()))


x = torch.randn(1, 62, 54, 256).cuda(); w = torch.randn(256, 256, 3, 3) / 48
pw = L.pack_linear(w.permute(2, 3, 1, 0).reshape(9 * 256, 256)); s = torch.ones(256).cuda(); t = torch.zeros(256).cuda()
o = torch.empty(1, 62, 54, 256).cuda()
run("conv b3 62x54 256->256", lambda: L.check(L.lib().o3dml_conv3x3_nhwc_tc(L.ptr(x), 1, 62, 54, 256, 1, L.ptr(pw.img), pw.k_pad, pw.n_pad,
                                                                             L.ptr(s), L.ptr(t), 1, 0.0, L.ptr(o), 256, L.stream())), 72)
x1 = torch.randn(1, 248, 216, 64).cuda(); w1 = torch.randn(64, 64, 3, 3) / 24
pw1 = L.pack_linear(w1.permute(2, 3, 1, 0).reshape(9 * 64, 64)); s1 = torch.ones(64).cuda(); t1 = torch.zeros(64).cuda()
o1 = torch.empty(1, 248, 216, 64).cuda()
run("conv b1 248x216 64->64", lambda: L.check(L.lib().o3dml_conv3x3_nhwc_tc(L.ptr(x1), 1, 248, 216, 64, 1, L.ptr(pw1.img), pw1.k_pad, pw1.n_pad,
                                                                             L.ptr(s1), L.ptr(t1), 1, 0.0, L.ptr(o1), 64, L.stream())), 18)
L.TC_MIN_K = 8
x2 = torch.randn(5120, 3072).cuda(); w2 = torch.randn(3072, 1024) / 55
pw2 = L.pack_linear(w2); t2 = torch.zeros(1024).cuda(); o2 = torch.empty(5120, 1024).cuda()
run("linear 5120x3072->1024", lambda: L.linear([L.make_src(x2)], pw2, o2, None, t2, act="leaky", slope=0.1), 96)
