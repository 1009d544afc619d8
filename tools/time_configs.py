"""Wall time of the engine on the other BASELINE.json configs (parity-test shapes, not bench lines)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402
from dmpfold2_b200.predict import read_aln, encode_aln  # noqa: E402
from dmpfold2_b200.synth import synth_msa_structured  # noqa: E402

sd, _ = bench.load_weights()
eng = Engine(sd, 0)
base = encode_aln(read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))


def timed(msa, tmpl, n, m, reps=2):
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        c, f = eng.fold_host(msa, tmpl, n, m)
        best = min(best, time.perf_counter() - t)
    return best, c, f, eng.stage_times()


t, c, f, st = timed(base, None, 0, 0, 3)
print('cfg1 PF10963 L=82 N=252 n=0 m=0: %.1f ms' % (t * 1e3), st, flush=True)
t, c, f, st = timed(base, None, 10, 100, 3)
print('cfg1 PF10963 defaults (10/100): %.1f ms, mean conf %.4f' % (t * 1e3, f.mean()), flush=True)
msa = synth_msa_structured(base, 150, 512, 0)
t, c, f, st = timed(msa, None, 10, 100, 3)
print('cfg3 per target L=150 N=512 10/100: %.1f ms' % (t * 1e3), st, flush=True)
msa = synth_msa_structured(base, 1024, 2048, 0)
t0, c0, f0, _ = timed(msa, None, 0, 0, 1)
t, c, f, st = timed(msa, c0[:, 1].copy(), 30, 200, 1)
print('cfg4 L=1024 N=2048 30/200 template-seeded: %.1f ms' % (t * 1e3), st, flush=True)
msa = synth_msa_structured(base, 2048, 3000, 0)
t, c, f, st = timed(msa, None, 10, 100, 1)
print('cfg5 size on ONE GPU L=2048 N=3000 10/100: %.1f ms' % (t * 1e3), st, flush=True)
