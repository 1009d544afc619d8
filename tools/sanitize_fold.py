"""Tiny fold for compute-sanitizer (memcheck / synccheck): exercises every kernel once at small ragged sizes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd, _ = bench.load_weights()
eng = Engine(sd, 0)
base = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
for (l, n) in ((27, 13), (9, 1)):
    msa = np.ascontiguousarray(base[:n, 5:5 + l])
    for mode in ('f16f8', 'f16x3', 'f16', 'ffma'):
        eng.set_conv_mode(mode)
        c, f = eng.fold_host(msa, None, 1, 3)
        assert np.isfinite(c).all()
    c2, f2 = eng.fold_host(msa, c[:, 1].copy(), 0, 0)
    print(l, n, 'ok', float(f.mean()), flush=True)
# direct (non-Woodbury) DCA path: N >= 21 L
msa = O.synth_msa_random(8, 200, 1)
print('direct dca', float(eng.dca(msa).abs().max()), flush=True)
eng.close()
