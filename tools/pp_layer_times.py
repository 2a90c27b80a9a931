"""In-situ (warm clocks, warm L2) per-layer durations of the PointPillars dense part: CUDA events
around every launch of the eager path, averaged over steps.  usage: python tools/pp_layer_times.py"""
import sys; sys.path.insert(0, '.')
import torch
import bench
from open3d_ml_b200 import _lib as L

wl = bench.PointPillarsWorkload(1)
sd = bench.load_weights(wl)
wl_model = wl.make_model(sd)
wl_model.use_graph = False
frames = [f.cuda() for f in wl.build_inputs_gpu(0)]
recs = []
orig_conv, orig_lin = wl_model._conv, L.linear
lib = L.lib()


def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record()
    recs.append((name, a, b))
    return r


wl_model._conv = lambda x, B, H, W, name, stride, cin, cout: timed(
    "conv %s %dx%d %d->%d s%d" % (name, H, W, cin, cout, stride), lambda: orig_conv(x, B, H, W, name, stride, cin, cout))
real_deconv = lib.o3dml_deconv_nhwc_tc


class LibProxy:
    def __getattr__(self, k):
        f = getattr(lib, k)
        if k == "o3dml_deconv_nhwc_tc":
            return lambda *a: timed("deconv %dx%d %d s%d" % (a[2], a[3], a[4], a[5]), lambda: f(*a))
        if k == "o3dml_linear_tc":
            return lambda *a: timed("head linear", lambda: f(*a))
        return f


L_lib = L.lib
L.lib = lambda: LibProxy()
for _ in range(5):
    wl_model(frames)
recs.clear()
STEPS = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    wl_model(frames)
e1.record()
torch.cuda.synchronize()
tot = {}
order = []
for name, a, b in recs:
    if name not in tot:
        tot[name] = 0.0
        order.append(name)
    tot[name] += a.elapsed_time(b) * 1e3 / STEPS
print("step %.1f us (eager, events around each launch)" % (e0.elapsed_time(e1) * 1e3 / STEPS))
s = 0.0
for n in order:
    print("%8.1f us  %s" % (tot[n], n)); s += tot[n]
print("%8.1f us  sum of dense layers" % s)
