"""Time the vgru stage alone (device events) for the mode selected by DMP2_VGRU and compare it with the per-step mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402
from dmpfold2_b200.predict import read_aln, encode_aln  # noqa: E402
from dmpfold2_b200.synth import synth_msa_structured  # noqa: E402

sd, _ = bench.load_weights()
base = encode_aln(read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
res = {}
for mode in ('steps', 'persist'):
    os.environ['DMP2_VGRU'] = mode
    eng = Engine(sd, 0)
    for (L, N) in ((82, 252), (128, 300), (300, 1000), (384, 64), (500, 200), (1024, 64)):
        msa = synth_msa_structured(base, L, N, 0)
        for _ in range(2):
            out = eng.vgru(msa)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            out = eng.vgru(msa)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        res[(mode, L, N)] = out.cpu()
        print('%-8s L=%4d N=%4d  %.3f ms  (%.2f us/row incl. H2D of the MSA)' % (mode, L, N, ms, ms * 1e3 / N), flush=True)
    del eng
for (L, N) in ((82, 252), (128, 300), (300, 1000), (384, 64), (500, 200), (1024, 64)):
    d = (res[('steps', L, N)] - res[('persist', L, N)]).abs().max()
    print('L=%d N=%d persist vs steps max abs diff %.3e' % (L, N, float(d)), flush=True)
