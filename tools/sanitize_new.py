"""compute-sanitizer targets added after the first sanitizer pass: the fused cluster eigensolver and the vgru step
changes (inside a tiny fold), the halo-sharded path with one strip (window conv, k_push, k_stats_finalize), and the
conv epilogue with fused InstanceNorm sums (DMP2_FUSE_STATS=1)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402
from dmpfold2_b200.predict import read_aln, encode_aln  # noqa: E402

sd, _ = bench.load_weights()
base = encode_aln(read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
msa = np.ascontiguousarray(base[:13, 5:32])
eng = Engine(sd, 0)
c, f = eng.fold_host(msa, None, 1, 3)
print('fold ok', float(f.mean()), flush=True)
h, win = eng.strip_setup(0, 1, msa.shape[1], msa.shape[0])
eng.strip_attach_local([win])
c2, f2 = eng.fold_strip_host(msa, None, 1, 3)
eng.strip_detach()
print('strip(1) ok', float(np.abs(c2 - c).max()), flush=True)
eng.close()
os.environ['DMP2_FUSE_STATS'] = '1'
eng = Engine(sd, 0)
c3, f3 = eng.fold_host(msa, None, 1, 3)
print('fused stats ok', float(np.abs(c3 - c).max()), flush=True)
eng.close()
