import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch, numpy as np
import helpers as H
import open3d_ml_b200 as M
from open3d_ml_b200 import _lib as L

sd, _ = H.state_dict("randlanet_semantickitti.manifest.json", 7)
inp = H.randla_inputs(1, 45056 // 4, 500)   # level 3 has 176 points
net = M.RandLANetB200(sd, use_tc=False)
cap = {}
orig = net._lfa_pool
def hook(stage, d, coords, nidx, feat, B, N, p, agg):
    orig(stage, d, coords, nidx, feat, B, N, p, agg)
    if d == 256 and stage == 2:
        cap.update(coords=coords.clone(), nidx=nidx.clone(), feat=feat.clone(), B=B, N=N, p=p, agg=agg.clone())
net._lfa_pool = hook
net(inp)
w = net.w; p = cap["p"]
tcnet = M.RandLANetB200(sd, use_tc=True)
wt = tcnet.w

def run(coords, nidx, feat, weights=None, tag=""):
    B, N = cap["B"], cap["N"]
    ww, wwt = (w, wt) if weights is None else weights
    ref = torch.full((B * N, 256), float("nan")).cuda()
    out = torch.full((B * N, 256), float("nan")).cuda()
    L.check(L.lib().o3dml_randla_lfa_pool(2, 256, L.ptr(coords), L.ptr(nidx), 1, 16, L.ptr(feat), B, N,
        L.ptr(ww[p + ".lse1.mlp.wt"]), L.ptr(ww[p + ".lse1.mlp.s"]), L.ptr(ww[p + ".lse1.mlp.t"]),
        L.ptr(ww[p + ".lse2.mlp.wt"]), L.ptr(ww[p + ".lse2.mlp.s"]), L.ptr(ww[p + ".lse2.mlp.t"]),
        L.ptr(ww[p + ".pool2.score.wt"]), L.ptr(ww[p + ".pool2.score.b"]), L.ptr(ref), L.stream()))
    L.check(L.lib().o3dml_randla_lfa_pool_tc(2, 256, L.ptr(coords), L.ptr(nidx), 1, 16, L.ptr(feat), B, N,
        L.ptr(ww[p + ".lse1.mlp.wt"]), L.ptr(ww[p + ".lse1.mlp.s"]), L.ptr(ww[p + ".lse1.mlp.t"]),
        L.ptr(wwt[p + ".lse2.mlp.img"]), L.ptr(ww[p + ".lse2.mlp.wt"]), L.ptr(ww[p + ".lse2.mlp.s"]),
        L.ptr(ww[p + ".lse2.mlp.t"]), L.ptr(wwt[p + ".pool2.score.img"]), L.ptr(out), L.stream()))
    torch.cuda.synchronize()
    bad = ((out - ref).abs() > 1e-3 * ref.abs().max()) | torch.isnan(out)
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print("%-28s N=%d rel err %.3e  bad elems %d  bad rows %d (first %s) bad cols %d (first %s)" % (
        tag, N, err, int(bad.sum()), int(bad.any(1).sum()), bad.any(1).nonzero().flatten().tolist()[:12],
        int(bad.any(0).sum()), bad.any(0).nonzero().flatten().tolist()[:12]))
    return out, ref

g = torch.Generator().manual_seed(0)
c, n, f = cap["coords"], cap["nidx"], cap["feat"]
run(c, n, f, tag="captured")
run(c * 0.1, n, f, tag="coords x0.1")
run(c, n, torch.randn(f.shape, generator=g).cuda(), tag="feat randn")
run(c, torch.randint(0, cap["N"], n.shape, generator=g).cuda(), f, tag="nidx random")
run(torch.rand(c.shape, generator=g).cuda() * 10, n, f, tag="coords uniform[0,10)")
f2 = f.clone(); f2[:] = 1.0
run(c, n, f2, tag="feat ones")
print("---- scale sweep")
for sc in (0.01, 0.1, 0.3, 0.6, 1.0, 2.0):
    run(c, n, f * sc, tag="feat x%g" % sc)
out, ref = run(c, n, f, tag="captured again")
bad = (((out - ref).abs() > 1e-3 * ref.abs().max()) | torch.isnan(out)).nonzero()
for r, col in bad[:12].tolist():
    print("  bad (%d,%d): out %g ref %g" % (r, col, out[r, col].item(), ref[r, col].item()))
# neg-only / pos-only features
run(c, n, f.abs(), tag="feat abs")
run(c, n, -f.abs(), tag="feat -abs")
run(c, n, f.clamp(-20, 20), tag="feat clamp 20")
