"""Timeline (clock64) of CTA (0,0) of gemm_tc_kernel on a small-grid linear layer
(needs a build with O3DML_DEBUG_TIMING=1)."""
import sys, ctypes; sys.path.insert(0, '.')
import torch, numpy as np
from open3d_ml_b200 import _lib as L
rows, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(rows, K).cuda(); w = torch.randn(N, K) / K ** 0.5
pw = L.pack_linear(w.t().contiguous()); s = torch.ones(N).cuda(); t = torch.zeros(N).cuda()
o = torch.empty(rows, N).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
REP = 100
big = torch.empty(64 << 20, device="cuda")
for it in range(3):
    L.linear([L.make_src(x)], pw, o, s, t, act="leaky", slope=0.2)
big.zero_(); big.zero_(); big.zero_()   # a long kernel queue so that the launches below are GPU-bound
ev[0].record()
for it in range(REP):
    L.linear([L.make_src(x)], pw, o, s, t, act="leaky", slope=0.2)
ev[1].record(); torch.cuda.synchronize()
print("rows %d K %d N %d: %.1f us per launch (event, %d back-to-back)" % (rows, K, N, ev[0].elapsed_time(ev[1]) * 1e3 / REP, REP))
h = ctypes.CDLL(L.LIB_PATH)
buf = (ctypes.c_longlong * 7000)()
h.o3dml_gt_debug_read(buf, 7000)
ncta = min(1000, ((rows + 127) // 128) * max(1, N // 128))
g0 = np.array(buf[5000:5000 + ncta]); g1 = np.array(buf[6000:6000 + ncta])
print("globaltimer: first CTA start 0, last CTA start %d ns, first end %d ns, last end %d ns; per-CTA duration min %d max %d ns" % (
    g0.max() - g0.min(), g1.min() - g0.min(), g1.max() - g0.min(), (g1 - g0).min(), (g1 - g0).max()))
a = np.array(buf[:]); t0 = a[4000]
print("range pass done %d | last chunk done %d | end %d  (cycles)" % (a[4001] - t0, a[4002] - t0, a[4003] - t0))
nsl = K // 32
sl = a[:4 * nsl].reshape(nsl, 4); ld = a[1000:1000 + 4 * nsl].reshape(nsl, 4); mm = a[2000:2000 + 3 * nsl].reshape(nsl, 3)
print("slice | conv: start@ wait_full convert | loader: start@ wait_empty issue arrive | mma: start@ wait_conv issue")
for i in range(nsl if len(sys.argv) > 4 else 0):
    print("%3d | %7d %6d %6d | %7d %6d %6d %6d | %7d %6d %6d" % (
        i, sl[i, 0] - t0, sl[i, 1] - sl[i, 0], sl[i, 2] - sl[i, 1],
        ld[i, 0] - t0, ld[i, 1] - ld[i, 0], ld[i, 2] - ld[i, 1], ld[i, 3] - ld[i, 2],
        mm[i, 0] - t0, mm[i, 1] - mm[i, 0], mm[i, 2] - mm[i, 1]))
