"""Real-data check of a two-level accumulation for the 5x5 conv (GPU box): the conv of one ResNet block on real
activations as an explicit im2col GEMM through the tcgen05 pipeline (dmp2_gemm_tn_test), once as ONE accumulation
chain over K = 3200 and once as short chains (one launch per K-chunk) summed in fp32 with round-to-nearest.
Compared with the exact (fp64) result next to oneDNN's fp32 conv and torch's fp32 GEMM."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd = O.load_state_dict(os.path.join(ROOT, 'dmpfold2_b200', 'trained_model'))
orc = O.Oracle(sd)
msa = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
taps = {}
orc.fold(msa, iterations=0, minsteps=0, taps=taps)
eng = Engine(sd, 0)
for blk, key in ((1, 'stem'), (9, 'block8'), (16, 'block15')):
    x = taps[key]                                                  # (1,128,L,L)
    w = sd['resnet.%d.layer1.lin.weight' % blk].float()            # (512,128,5,5)
    a = F.unfold(x, 5, padding=2)[0].t().contiguous()              # (L*L, 3200), K order (c, ky, kx)
    b = w.reshape(512, 3200).contiguous()
    exact = a.double() @ b.double().t()
    rms, mx = float(exact.pow(2).mean().sqrt()), float(exact.abs().max())

    def rep(name, y):
        e = y.double() - exact
        delta = float((e * exact).sum() / (exact * exact).sum())
        r = e - delta * exact
        print('   %-34s rel-to-rms %.2e (rel-to-max %.2e)  shrink %+.2e  residual %.2e' % (
            name, float(e.pow(2).mean().sqrt()) / rms, float(e.pow(2).mean().sqrt()) / mx, delta, float(r.pow(2).mean().sqrt()) / rms), flush=True)

    print('block %d: |out| rms %.3g max %.3g' % (blk, rms, mx), flush=True)
    y_dnn = F.conv2d(x, w, None, padding=2)[0].permute(1, 2, 0).reshape(-1, 512)
    rep('oneDNN fp32 conv2d (reference)', y_dnn)
    rep('torch fp32 matmul', a @ b.t())
    for mode in ('f16x3', 'f16'):
        for chunk in (3200, 640, 128, 64):
            tot = torch.zeros(a.shape[0], 512)
            for k0 in range(0, 3200, chunk):
                tot += eng.gemm_tn_test(a[:, k0:k0 + chunk].contiguous(), b[:, k0:k0 + chunk].contiguous(), mode).cpu()
            rep('tcgen05 %s, chains of K=%d' % (mode, chunk), tot)
eng.close()
