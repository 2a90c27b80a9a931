import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
import helpers as H
import open3d_ml_b200 as M

g = H.golden("randlanet_small.npz")
sd, _ = H.state_dict("randlanet_semantickitti.manifest.json", g["weight_seed"])
inp = H.randla_inputs(int(g["B"]), int(g["N"]), int(g["seed0"]))
rec = {}
for tc_on in (False, True):
    net = M.RandLANetB200(sd, use_tc=tc_on)
    orig = net._lfa_pool
    calls = []
    def hook(stage, d, coords, nidx, feat, B, N, p, agg, orig=orig, calls=calls):
        orig(stage, d, coords, nidx, feat, B, N, p, agg)
        torch.cuda.synchronize()
        calls.append((stage, d, N, feat.clone(), agg.clone()))
    net._lfa_pool = hook
    out = net(inp)
    rec[tc_on] = calls
for (a, b) in zip(rec[False], rec[True]):
    s, d, N, f0, g0 = a
    _, _, _, f1, g1 = b
    nan = torch.isnan(g1)
    err = ((g1 - g0).abs().max() / g0.abs().max()).item()
    print("stage", s, "d", d, "N", N, "feat equal", torch.equal(f0, f1), "feat absmax %.3g" % f0.abs().max().item(),
          "agg absmax %.3g" % g0.abs().max().item(), "rel err %.3e" % err, "nan rows", nan.any(1).nonzero().flatten().tolist()[:20],
          "nan cols", nan.any(0).nonzero().flatten().tolist()[:20])
    if err > 1e-3 or nan.any():
        bad = ((g1 - g0).abs() > 1e-3 * g0.abs().max()) | nan
        print("   bad rows", bad.any(1).nonzero().flatten().tolist()[:40])
        print("   bad cols", bad.any(0).nonzero().flatten().tolist()[:40])
