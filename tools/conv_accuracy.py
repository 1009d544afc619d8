"""Accuracy of the conv kernel modes against an fp64 CPU convolution on realistic activations (GPU box only)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd, _ = bench.load_weights()
orc = O.Oracle(sd)
msa = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
taps = {}
orc.fold(msa, iterations=0, minsteps=0, taps=taps)
eng = Engine(sd, 0)
for blk, key in ((1, 'stem'), (9, 'block8'), (16, 'block15')):
    x = taps[key]                                             # input of block `blk`
    w = sd[f'resnet.{blk}.layer1.lin.weight'].double()
    b = sd[f'resnet.{blk}.layer1.lin.bias'].double()
    y = F.conv2d(x.double(), w, b, padding=2)
    ref = y.view(1, 128, 4, y.shape[2], y.shape[3]).max(dim=2)[0][0].permute(1, 2, 0)
    y32 = F.conv2d(x, w.float(), b.float(), padding=2)
    cpu32 = y32.view(1, 128, 4, y.shape[2], y.shape[3]).max(dim=2)[0][0].permute(1, 2, 0).double()
    scale = ref.abs().max()
    print(f'block {blk}: |x|max {float(x.abs().max()):.1f} |out|max {float(scale):.1f}  oneDNN-fp32 rel err {float((cpu32 - ref).abs().max() / scale):.2e}', flush=True)
    for mode in ('ffma', 'f16x3', 'f16f8', 'f16'):
        eng.set_conv_mode(mode)
        got = eng.conv5_maxout(blk, x[0].permute(1, 2, 0).contiguous()).cpu().double()
        err = got - ref
        print(f'   {mode:6s} max rel err {float(err.abs().max() / scale):.2e}  rms rel {float(err.pow(2).mean().sqrt() / scale):.2e}  mean signed {float((err * ref.sign()).mean() / scale):+.2e}', flush=True)
