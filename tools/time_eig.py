"""Eigen step (top-8 of a symmetric L x L matrix) at several L: time, phase split, accuracy against fp64 eigh."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd, _ = bench.load_weights()
eng = Engine(sd, 0)
for l in [int(a) for a in sys.argv[1:]] or [300, 640, 700, 1024, 1500, 2048]:
    g = torch.Generator().manual_seed(l)
    x = torch.randn(l, 3, generator=g, dtype=torch.float64).cumsum(0) * 2.0
    d2 = ((x[:, None] - x[None]) ** 2).sum(-1)
    m = 0.5 * (d2[0][None, :] + d2[:, 0][:, None] - d2) + torch.randn(l, l, generator=g, dtype=torch.float64) * 0.5
    m = ((m + m.t()) / 2).float().cuda()
    for _ in range(2):
        vals, vecs = eng.eig_top8(m)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        vals, vecs = eng.eig_top8(m)
    b.record()
    torch.cuda.synchronize()
    w, v = torch.linalg.eigh(m.double().cpu())
    idx = v.abs().argmax(0, keepdim=True)
    v = v * torch.gather(v, 0, idx).sign()
    print('L=%5d  %.3f ms  phases %s  max|dval| %.2e  max|dvec| %.2e' % (
        l, a.elapsed_time(b) / 3, {k: round(v_ / 1e3, 3) for k, v_ in eng.eig_phases(l).items()},
        float((vals.cpu().double() - w[-8:]).abs().max()), float((vecs.cpu().double() - v[:, -8:]).abs().max())), flush=True)
