"""Where does the cfg2 (L=300, N=1000) deviation from the oracle come from?  Diagnostic, GPU box only."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
sd, _ = bench.load_weights()
base = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
msa = O.synth_msa_structured(base, L, N, 0)
orc = O.Oracle(sd)


def rmsd(a, b):
    return O.kabsch_rmsd(np.asarray(a)[:, 1], np.asarray(b)[:, 1])


t = time.time()
refs = {}
for n, m in ((0, 0), (10, 100)):
    refs[(n, m)] = orc.fold(msa, iterations=n, minsteps=m)
print('oracle time', time.time() - t, flush=True)
torch.set_num_threads(4)
alt = {}
for n, m in ((0, 0), (10, 100)):
    alt[(n, m)] = orc.fold(msa, iterations=n, minsteps=m)
torch.set_num_threads(os.cpu_count())
for k in refs:
    print('oracle self-noise (all threads vs 4 threads)', k, 'CA-RMSD %.2e' % rmsd(refs[k][0].numpy(), alt[k][0].numpy()),
          'conf %.2e' % float((refs[k][1] - alt[k][1]).abs().max()), flush=True)

eng = Engine(sd, 0)
# stage-level
v_ref = orc.vgru_last(torch.from_numpy(msa))
v = eng.vgru(msa).cpu()
print('vgru max abs err %.2e (max |ref| %.2f)' % (float((v - v_ref).abs().max()), float(v_ref.abs().max())), flush=True)
h_ref = orc.hgru_out(v_ref)
print('hgru (teacher-forced) max abs err %.2e' % float((eng.hgru(v_ref).cpu() - h_ref).abs().max()), flush=True)
f_ref = O.msa_features(torch.from_numpy(msa))
f = eng.dca(msa).cpu()
print('dca max abs err %.2e (max |ref| %.2f)' % (float((f - f_ref).abs().max()), float(f_ref.abs().max())), flush=True)
del f, f_ref

for mode in ('ffma', 'f16x3', 'f16f8', 'f16'):
    eng.set_conv_mode(mode)
    for k in refs:
        c, cf = eng.fold_host(msa, None, k[0], k[1])
        print(mode, k, 'CA-RMSD %.2e' % rmsd(c, refs[k][0].numpy()), 'vs alt %.2e' % rmsd(c, alt[k][0].numpy()),
              'conf %.2e' % float(np.abs(cf - refs[k][1].numpy()).max()), 'mean conf %.4f (ref %.4f)' % (cf.mean(), float(refs[k][1].mean())), flush=True)
eng.close()
os.environ['DMP2_VGRU'] = 'ffma'
eng = Engine(sd, 0, conv_mode='ffma')
for k in refs:
    c, cf = eng.fold_host(msa, None, k[0], k[1])
    print('ffma conv + ffma vgru', k, 'CA-RMSD %.2e' % rmsd(c, refs[k][0].numpy()), flush=True)
