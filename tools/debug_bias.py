import sys; sys.path.insert(0, '.')
import torch
from open3d_ml_b200 import _lib as L
# systematic (signed) error of the tcgen05 accumulation: positive operands expose truncation
for k in (64, 256, 1024):
    for terms in (1, 3):
        g = torch.Generator().manual_seed(k)
        a = (torch.rand(128, k, generator=g) + 0.5).cuda()
        b = (torch.rand(16, k, generator=g) + 0.5).cuda()
        if terms == 1:   # make operands exactly fp16 so that only the accumulation errs
            a, b = a.half().float(), b.half().float()
        d = torch.empty(128, 16).cuda()
        L.check(L.lib().o3dml_tc_gemm_test(L.ptr(a), L.ptr(b), L.ptr(d), 16, k, terms, L.stream()))
        ref = a.double() @ b.double().t()
        rel = (d.double() - ref) / ref
        f32 = ((a @ b.t()).double() - ref) / ref
        print("K %5d terms %d: mean signed rel err %+.3e  max |rel| %.3e   (fp32 cuBLAS: mean %+.3e max %.3e)" % (
            k, terms, rel.mean().item(), rel.abs().max().item(), f32.mean().item(), f32.abs().max().item()))
