"""How does the tcgen05 fp32 accumulator round?  (GPU box)  Operands are exactly fp16-representable, so in the
single-MMA `f16` mode every product is exact and the only error is the accumulation inside / across MMAs.
Reports the signed error split by the sign of the exact result: truncation toward zero shrinks both signs,
truncation toward -inf shifts both signs down."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd = O.random_state_dict(0)
eng = Engine(sd, 0)
g = torch.Generator().manual_seed(1)
for dist in ('normal', 'positive'):
    for K in (64, 256, 1024, 3200):
        a = torch.randn(512, K, generator=g).half().float()
        b = torch.randn(512, K, generator=g).half().float()
        if dist == 'positive':
            a, b = a.abs(), b.abs()
        exact = a.double() @ b.double().t()
        ref32 = (a @ b.t()).double()
        for mode in ('f16', 'f16x3'):
            c = eng.gemm_tn_test(a, b, mode).cpu().double()
            e = c - exact
            ulp = torch.tensor(np.spacing(exact.abs().float().numpy())).double()
            eu = e / ulp
            pos, neg = exact > 0, exact < 0
            print('%-8s K=%4d %-5s  err/ulp: mean(+) %+.3f mean(-) %+.3f  rms %.3f  max %.1f | rel rms %.2e | torch fp32 matmul rms %.3f ulp' % (
                dist, K, mode, float(eu[pos].mean()), float(eu[neg].mean()) if neg.any() else 0.0, float(eu.pow(2).mean().sqrt()),
                float(eu.abs().max()), float((e.pow(2).mean() / exact.pow(2).mean()).sqrt()),
                float(((ref32 - exact) / ulp).pow(2).mean().sqrt())), flush=True)

# ---- what would a two-level accumulation buy?  Short tcgen05 chains (one launch per K-chunk) summed in fp32 with RN.
print('two-level accumulation, normal data, K=3200, f16 mode (products exact):', flush=True)
a = torch.randn(512, 3200, generator=g).half().float()
b = torch.randn(512, 3200, generator=g).half().float()
exact = a.double() @ b.double().t()
rms = float(exact.pow(2).mean().sqrt())
ref32 = (a @ b.t()).double()
print('   torch fp32 matmul          rel rms %.2e' % (float((ref32 - exact).pow(2).mean().sqrt()) / rms), flush=True)
for chunk in (3200, 640, 256, 128, 64):
    tot = torch.zeros(512, 512)
    for k0 in range(0, 3200, chunk):
        tot += eng.gemm_tn_test(a[:, k0:k0 + chunk].contiguous(), b[:, k0:k0 + chunk].contiguous(), 'f16').cpu()
    e = tot.double() - exact
    delta = float((e * exact).sum() / (exact * exact).sum())
    resid = e - delta * exact
    print('   chunk K=%4d (%3d MMAs/chain) rel rms %.2e   shrink delta %+.2e   residual after removing it %.2e' % (
        chunk, chunk // 16, float(e.pow(2).mean().sqrt()) / rms, delta, float(resid.pow(2).mean().sqrt()) / rms), flush=True)
eng.close()
