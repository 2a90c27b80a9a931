import sys, ctypes; sys.path.insert(0, '.')
import torch, numpy as np
from open3d_ml_b200 import _lib as L
out = torch.zeros(2, dtype=torch.int64).cuda()
for n in (32, 64, 128, 256):
    for reps in (16, 256):
        L.check(L.lib().o3dml_tc_mma_rate(n, reps, L.ptr(out), L.stream())); torch.cuda.synchronize()
        tot, iss = out.tolist()
        print("N %3d reps %4d: %7d cycles total (%.1f per MMA), issue %7d (%.1f per MMA); ideal %.1f" % (n, reps, tot, tot / (6 * reps), iss, iss / (6 * reps), 128 * n * 16 / 4096))
# per-slice timeline of one block-3 conv CTA (62x54x256 -> 256, K = 2304)
x = torch.randn(1, 62, 54, 256).cuda(); w = torch.randn(256, 256, 3, 3) / 48
pw = L.pack_linear(w.permute(2, 3, 1, 0).reshape(9 * 256, 256)); s = torch.ones(256).cuda(); t = torch.zeros(256).cuda()
o = torch.empty(1, 62, 54, 256).cuda()
for _ in range(3):
    L.check(L.lib().o3dml_conv3x3_nhwc_tc(L.ptr(x), 1, 62, 54, 256, 1, L.ptr(pw.img), pw.k_pad, pw.n_pad, L.ptr(s), L.ptr(t), 1, 0.0, L.ptr(o), 256, L.stream()))
torch.cuda.synchronize()
h = ctypes.CDLL(L.LIB_PATH)
buf = (ctypes.c_longlong * 4100)()
print("read rc", h.o3dml_gt_debug_read(buf, 4100))
a = np.array(buf[:])
t0 = a[4000]
print("kernel start 0 | range pass done %d | last chunk done %d | end %d  (cycles)" % (a[4001] - t0, a[4002] - t0, a[4003] - t0))
sl = a[:288].reshape(72, 4)
print("slice: wait_full  convert+fence+arrive   start@")
for i in range(0, 72, 3):
    print("%3d %8d %8d   %9d" % (i, sl[i, 1] - sl[i, 0], sl[i, 2] - sl[i, 1], sl[i, 0] - t0))

ld = a[1000:1000 + 288].reshape(72, 4)
mm = a[2000:2000 + 216].reshape(72, 3)
print("loader rt0: slice  wait_empty  issue  arrive | MMA warp: wait_conv  issue")
for i in range(0, 72, 4):
    print("%3d %8d %8d %6d   | %8d %8d" % (i, ld[i, 1] - ld[i, 0], ld[i, 2] - ld[i, 1], ld[i, 3] - ld[i, 2], mm[i, 1] - mm[i, 0], mm[i, 2] - mm[i, 1]))
