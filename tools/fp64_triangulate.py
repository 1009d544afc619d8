"""Precision triangulation of a fold: how far is the fp32 reference algorithm from its own fp64 evaluation?

The recycling loop amplifies rounding-level perturbations: at cfg2 the fp32 oracle differs from
ITSELF by ~5e-4 A between thread counts.  A distance "engine vs fp32 oracle" is therefore only meaningful next to
the distance "fp32 oracle vs exact arithmetic".  This script runs, on the CPU, the oracle (= the reference's
algorithm, pinned against /root/reference by oracle/make_golden.py) three times on one synthetic target:

    ref32      fp32, all host threads          (the parity target)
    ref32_alt  fp32, 4 host threads            (the reference's own irreproducibility)
    ref64      fp64, same fp32-valued weights  (ground truth)

and writes a small fixture (coords/confs of all three + the generator arguments) to tests/golden/.  The GPU tests
compare the engine with ref32 (the bar) and report its distance to ref64 beside the reference's.

    python tools/fp64_triangulate.py NAME L N SEED n m [generator] [template]

template = "domains": seed the fold with copies of the reference's PF10963 prediction (O.synth_template_domains).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402


def make_msa(gen, base, L, N, seed):
    if gen == 'structured':
        return O.synth_msa_structured(base, L, N, seed)
    if gen == 'tandem':
        return O.synth_msa_tandem(base, L, N, seed)
    raise SystemExit('unknown generator ' + gen)


def main():
    name, L, N, seed, n, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    gen = sys.argv[7] if len(sys.argv) > 7 else 'structured'
    tmpl_kind = sys.argv[8] if len(sys.argv) > 8 else 'none'
    wdir = os.path.join(ROOT, 'dmpfold2_b200', 'trained_model')
    sd = O.load_state_dict(wdir)
    base = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
    msa = make_msa(gen, base, L, N, seed)
    tmpl = None
    if tmpl_kind == 'domains':
        dom = np.load(os.path.join(ROOT, 'tests', 'golden', 'pf10963_n10_m100.npz'))['coords'][:, 1]
        tmpl = O.synth_template_domains(dom, L)
    out = {'L': L, 'N': N, 'seed': seed, 'iterations': n, 'minsteps': m, 'generator': gen, 'template': tmpl_kind,
           'torch': torch.__version__, 'threads': torch.get_num_threads()}

    def rmsd(a, b):
        return O.kabsch_rmsd(np.asarray(a)[:, 1], np.asarray(b)[:, 1])

    orc = O.Oracle(sd)
    for tag, nm in (('ref32', (n, m)), ('ref32_pass', (0, 0))):
        t = time.time()
        c, f = orc.fold(msa, template_ca=tmpl, iterations=nm[0], minsteps=nm[1])
        out[tag + '_coords'], out[tag + '_confs'] = c.numpy(), f.numpy()
        print(tag, nm, 'time %.1fs mean conf %.4f' % (time.time() - t, float(f.mean())), flush=True)
    nthr = torch.get_num_threads()
    torch.set_num_threads(4 if nthr != 4 else 3)
    for tag, nm in (('ref32_alt', (n, m)), ('ref32_alt_pass', (0, 0))):
        c, f = orc.fold(msa, template_ca=tmpl, iterations=nm[0], minsteps=nm[1])
        out[tag + '_coords'], out[tag + '_confs'] = c.numpy(), f.numpy()
    torch.set_num_threads(nthr)
    print('self-noise  %d/%d: %.3e A   0/0: %.3e A' % (n, m, rmsd(out['ref32_coords'], out['ref32_alt_coords']),
                                                         rmsd(out['ref32_pass_coords'], out['ref32_alt_pass_coords'])), flush=True)
    orc64 = O.Oracle(sd, dtype=torch.float64)
    for tag, nm in (('ref64_pass', (0, 0)), ('ref64', (n, m))):
        t = time.time()
        c, f = orc64.fold(msa, template_ca=tmpl, iterations=nm[0], minsteps=nm[1])
        out[tag + '_coords'], out[tag + '_confs'] = c.numpy(), f.numpy()
        print(tag, nm, 'time %.1fs mean conf %.4f' % (time.time() - t, float(f.mean())), flush=True)
    print('fp32 vs fp64  %d/%d: %.3e A (alt threads %.3e)   0/0: %.3e A (alt %.3e)' % (
        n, m, rmsd(out['ref32_coords'], out['ref64_coords']), rmsd(out['ref32_alt_coords'], out['ref64_coords']),
        rmsd(out['ref32_pass_coords'], out['ref64_pass_coords']), rmsd(out['ref32_alt_pass_coords'], out['ref64_pass_coords'])),
        flush=True)
    dst = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, os.path.getsize(dst), 'bytes')


if __name__ == '__main__':
    main()
