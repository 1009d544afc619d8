"""Fold synthetic targets on the GPU in every conv mode and dump the coordinates (GPU box).  The comparison with the
fp32 / fp64 oracle fixtures (tools/fp64_triangulate.py) happens wherever those fixtures are.

    python tools/diag_fold_dump.py NAME L N SEED n m [generator] [modes,comma,separated]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402  (synthetic generators + metric only)
from dmpfold2_b200.engine import Engine  # noqa: E402

name, L, N, seed, n, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
gen = sys.argv[7] if len(sys.argv) > 7 else 'structured'
modes = sys.argv[8].split(',') if len(sys.argv) > 8 else ['ffma', 'f16x3', 'f16f8']
tmpl_kind = sys.argv[9] if len(sys.argv) > 9 else 'none' 
sd = O.load_state_dict(os.path.join(ROOT, 'dmpfold2_b200', 'trained_model'))
base = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
msa = getattr(O, 'synth_msa_' + gen)(base, L, N, seed)
tmpl = None
if tmpl_kind == 'domains':
    tmpl = O.synth_template_domains(np.load(os.path.join(ROOT, 'tests', 'golden', 'pf10963_n10_m100.npz'))['coords'][:, 1], L)
eng = Engine(sd, 0)
out = {}
gold = None
gp = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
if os.path.isfile(gp):
    gold = np.load(gp)
for mode in modes:
    eng.set_conv_mode(mode)
    for tag, nm in (('pass', (0, 0)), ('full', (n, m))):
        c, f = eng.fold_host(msa, tmpl, nm[0], nm[1])
        out['%s_%s_coords' % (mode, tag)], out['%s_%s_confs' % (mode, tag)] = c, f
        line = '%s %-6s %s mean conf %.4f' % (name, mode, nm, float(f.mean()))
        if gold is not None:
            sfx = '_pass' if tag == 'pass' else ''
            r32 = O.kabsch_rmsd(c[:, 1], gold['ref32%s_coords' % sfx][:, 1])
            r64 = O.kabsch_rmsd(c[:, 1], gold['ref64%s_coords' % sfx][:, 1])
            rr = O.kabsch_rmsd(gold['ref32%s_coords' % sfx][:, 1], gold['ref64%s_coords' % sfx][:, 1])
            line += '  CA-RMSD eng-ref32 %.2e  eng-ref64 %.2e  (ref32-ref64 %.2e)' % (r32, r64, rr)
        print(line, flush=True)
eng.close()
os.makedirs(os.path.join(ROOT, 'gpurun_out', 'diag'), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'diag', 'fold_%s.npz' % name), **out)
