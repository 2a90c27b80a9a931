"""Builds the RandLA-Net k-NN pyramid (8 x 45 056 points) once eagerly: run under
`ncu --metrics gpu__time_duration.sum --profile-from-start off` to get the per-launch list."""
import sys; sys.path.insert(0, '.')
import torch
import bench
wl = bench.RandLAWorkload(8)
sd = bench.load_weights(wl)
net = wl.make_model(sd)
from open3d_ml_b200 import synth
pts = torch.stack([torch.from_numpy(synth.semantickitti_cloud(wl.N, 1000 + u)) for u in range(8)]).cuda()
for _ in range(3):
    net.build_pyramid(pts)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
net.build_pyramid(pts)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    net.build_pyramid(pts)
b.record(); torch.cuda.synchronize()
print("pyramid eager %.3f ms" % (a.elapsed_time(b) / 10))
