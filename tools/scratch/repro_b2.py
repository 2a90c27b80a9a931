import sys; sys.path.insert(0, '.')
import torch, bench
from open3d_ml_b200 import PipelinedRunner
wl = bench.RandLAWorkload(2)
sd = bench.load_weights(wl)
net = wl.make_model(sd)
host = wl.build_inputs_gpu([4, 5])
dev = bench.to_dev(host, "cuda")
for i in range(3):
    out = net.forward_graphed(dev)
torch.cuda.synchronize(); print("resident ok", float(out.abs().mean()))
runner = PipelinedRunner(lambda d: net.forward_graphed(d), "cuda")
for r in runner.run(host for _ in range(5)):
    pass
torch.cuda.synchronize(); print("runner ok")
pts = dict(points=host["coords"][0].clone().pin_memory())
runner2 = PipelinedRunner(lambda d: net.forward_points_graphed(d["points"]), "cuda")
for r in runner2.run(pts for _ in range(5)):
    pass
torch.cuda.synchronize(); print("points runner ok")
