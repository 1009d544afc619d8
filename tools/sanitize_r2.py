"""compute-sanitizer targets of round 2 (tiny ragged sizes, every new code path once): the persistent conv kernel in its
three forms (CTA pairs, independent CTAs, 2-CTA weight multicast) x three precision modes with the fused InstanceNorm
sums, on few SMs (many units per CTA: chunk-buffer phase wrap-around); the tcgen05 GEMM service with all epilogues
(stem, Gram / Woodbury / covariance, GRU projections with bias) incl. the direct-DCA branch; the four-chain vgru step;
the 16-CTA eigensolver with two inverse-iteration sweeps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402  (synthetic generator only)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd, _ = bench.load_weights()
base = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
msa = np.ascontiguousarray(base[:13, 5:38])                   # L = 33: 3 x 5 tiles, ragged in both directions
ref = None
for form, sms in (('pair', '0'), ('pair', '4'), ('1', '3'), ('2', '6')):
    os.environ['DMP2_CONV_CLUSTER'] = form
    os.environ['DMP2_CONV_SMS'] = sms
    eng = Engine(sd, 0)
    for mode in ('f16f8', 'f16x3', 'f16'):
        eng.set_conv_mode(mode)
        c, f = eng.fold_host(msa, None, 1, 3)
        assert np.isfinite(c).all()
        if mode == 'f16x3':
            if ref is None:
                ref = c
            print('conv form', form, 'sms', sms, mode, 'max|dcoords| vs first form %.2e' % float(np.abs(c - ref).max()), flush=True)
    eng.close()
os.environ.pop('DMP2_CONV_CLUSTER'); os.environ.pop('DMP2_CONV_SMS')
eng = Engine(sd, 0)
deep = O.synth_msa_random(9, 230, 1)                          # N >= 21 L: direct covariance branch on the GEMM service
print('direct dca', float(eng.dca(deep).abs().max()), flush=True)
wide = O.synth_msa_structured(base, 210, 24, 2)               # L >= 200: 16-CTA eigensolver; L > 128: two vgru row tiles
c, f = eng.fold_host(wide, None, 1, 3)
print('L=210 fold', float(f.mean()), bool(np.isfinite(c).all()), flush=True)
eng.close()
