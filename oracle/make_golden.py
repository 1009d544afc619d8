"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported from /root/reference) in the build
container.  TEST INFRASTRUCTURE ONLY -- never imported by the product.

The reference needs one shim to run on torch >= 1.13: `torch.symeig` was removed (network.py:247,292).  The
shim below is "oracle B" of SURVEY.md section A.1: eigh on the upper triangle + canonical eigenvector sign.
Nothing else of the reference is modified; its modules, weights loader and forward run as shipped.

Usage (build container only; /root/reference does not exist on the GPU box):
    python oracle/make_golden.py            # writes tests/golden/pf10963_*.npz
"""
import os
import sys
import tempfile

import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, '..', 'tests', 'golden')


def install_shim():
    def _symeig(a, eigenvectors=False, upper=True):
        w, v = torch.linalg.eigh(a, UPLO='U' if upper else 'L')
        idx = v.abs().argmax(dim=-2, keepdim=True)
        v = v * torch.gather(v, -2, idx).sign()
        return w, v
    torch.symeig = _symeig


def write_ca_pdb(path, ca):
    with open(path, 'w') as fh:
        for i, (x, y, z) in enumerate(ca):
            fh.write("ATOM  %5d  CA  ALA A%4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, i + 1, x, y, z))
        fh.write("END\n")


def main():
    install_shim()
    sys.path.insert(0, REF)
    import dmpfold
    from dmpfold import network as refnet

    aln = os.path.join(REF, 'dmpfold', 'example', 'PF10963.aln')
    torch.set_num_threads(8)

    # ---- capture intermediates of one n=0,m=0 run through forward hooks on the reference modules
    taps = {}
    orig_forward = refnet.GRUResNet.forward

    def tap(name, fn):
        def hook(m, i, o):           # must return None, otherwise torch replaces the module output
            if name not in taps:
                taps[name] = fn(i, o).clone()
        return hook

    def hooked_forward(self, x, x2, nloops=5, refine_steps=0):
        hs = []
        hs.append(self.vgru.register_forward_hook(tap('vgru_last', lambda i, o: o[0][-1])))
        hs.append(self.hgru.register_forward_hook(tap('mat1d', lambda i, o: o[0][:, 0].t())))
        hs.append(self.resnet[0].register_forward_hook(tap('stem', lambda i, o: o[0])))
        hs.append(self.resnet[1].layer1.lin.register_forward_hook(tap('conv1', lambda i, o: o[0])))
        hs.append(self.resnet[1].register_forward_hook(tap('block1', lambda i, o: o[0])))
        hs.append(self.resnet[16].register_forward_hook(tap('block16', lambda i, o: o[0])))
        hs.append(self.resnet[17].register_forward_hook(tap('head', lambda i, o: o[0])))
        hs.append(self.coord_gru.register_forward_hook(tap('mds', lambda i, o: i[0][0, :, 512:])))
        hs.append(self.coord_fc.register_forward_hook(tap('ca0', lambda i, o: o[0])))
        taps['x2_sub'] = x2[0, :, ::9, ::7].clone()
        taps['x2_apc'] = x2[0, 441].clone()
        try:
            return orig_forward(self, x, x2, nloops, refine_steps)
        finally:
            for h in hs:
                h.remove()

    refnet.GRUResNet.forward = hooked_forward
    coords, confs, alnmat = dmpfold.aln_to_coords(aln, iterations=0, minsteps=0, return_alnmat=True)
    refnet.GRUResNet.forward = orig_forward

    from dmpfold.predict import reweight as ref_reweight
    import torch.nn.functional as F
    hot = F.one_hot(torch.clamp(torch.from_numpy(alnmat).long(), max=20), 21).float()
    w = ref_reweight(hot, 0.8)

    out = {
        'alnmat': alnmat, 'coords': coords.numpy(), 'confs': confs.numpy(), 'w': w.numpy(),
        'vgru_last': taps['vgru_last'].numpy(), 'mat1d': taps['mat1d'].numpy(),
        'x2_sub': taps['x2_sub'].numpy(), 'x2_apc': taps['x2_apc'].numpy(),
        'stem_sub': taps['stem'][:, ::8, ::8].numpy(), 'conv1_sub': taps['conv1'][:, ::8, ::8].numpy(),
        'block1_sub': taps['block1'][:, ::8, ::8].numpy(), 'block16_sub': taps['block16'][:, ::8, ::8].numpy(),
        'head': taps['head'].numpy(), 'mds': taps['mds'].numpy(), 'ca0': taps['ca0'].numpy(),
    }
    np.savez_compressed(os.path.join(GOLD, 'pf10963_n0_m0.npz'), **out)
    print('n0m0 mean conf', float(confs.mean()))

    # ---- the CLI's stdout for the same n=0, m=0 run (predict.py:195-208), for the PDB-writer byte-exactness test
    import contextlib
    import io
    buf = io.StringIO()
    argv = sys.argv
    sys.argv = ['dmpfold', '-i', aln, '-n', '0', '-m', '0']
    try:
        with contextlib.redirect_stdout(buf):
            dmpfold.run_dmpfold()
    finally:
        sys.argv = argv
    with open(os.path.join(GOLD, 'pf10963_n0_m0.pdb'), 'w') as fh:
        fh.write(buf.getvalue())

    for n, m in ((2, 20), (10, 100)):
        c, f = dmpfold.aln_to_coords(aln, iterations=n, minsteps=m)
        np.savez_compressed(os.path.join(GOLD, f'pf10963_n{n}_m{m}.npz'), coords=c.numpy(), confs=f.numpy())
        print(f'n{n}m{m} mean conf', float(f.mean()))

    # ---- template mode: the reference's own n=0 CA trace as a CA-only PDB (3FGX.pdb is not PF10963)
    with tempfile.TemporaryDirectory() as td:
        pdb = os.path.join(td, 'tmpl.pdb')
        write_ca_pdb(pdb, coords[:, 1].numpy())
        with open(pdb) as fh:
            pdb_text = fh.read()
        c, f = dmpfold.aln_to_coords(aln, template=pdb, iterations=1, minsteps=10)
        np.savez_compressed(os.path.join(GOLD, 'pf10963_tmpl_n1_m10.npz'), coords=c.numpy(), confs=f.numpy(),
                            pdb_text=np.array(pdb_text))
        print('template mean conf', float(f.mean()))

        # ---- single-sequence path (predict.py:139: zero DCA features)
        single = os.path.join(td, 'single.aln')
        with open(aln) as fh, open(single, 'w') as fo:
            fo.write(fh.readline())
        c, f = dmpfold.aln_to_coords(single, iterations=1, minsteps=0)
        np.savez_compressed(os.path.join(GOLD, 'pf10963_single_n1_m0.npz'), coords=c.numpy(), confs=f.numpy())
        print('single mean conf', float(f.mean()))


def synthetic_golden():
    """A second pin at another size: the structured synthetic alignment the halo-sharded tests use (L=100, N=96,
    seed 3), written out as an .aln file and folded by the reference (n=2, m=20)."""
    install_shim()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, '..'))
    import dmpfold
    from oracle import dmpfold_oracle as O
    torch.set_num_threads(8)
    base = O.encode_aln(O.read_aln(os.path.join(REF, 'dmpfold', 'example', 'PF10963.aln')))
    msa = O.synth_msa_structured(base, 100, 96, 3)
    letters = 'ARNDCQEGHILKMFPSTWYVX-'                      # codes 0..19, 20 = unknown, 21 = gap (predict.py:124-128)
    with tempfile.TemporaryDirectory() as td:
        aln = os.path.join(td, 'synth.aln')
        with open(aln, 'w') as fh:
            for row in msa:
                fh.write(''.join(letters[c] for c in row) + '\n')
        c, f, alnmat = dmpfold.aln_to_coords(aln, iterations=2, minsteps=20, return_alnmat=True)
    assert np.array_equal(alnmat, msa), 'the .aln round trip must reproduce the synthetic codes'
    np.savez_compressed(os.path.join(GOLD, 'synth_l100_n96_s3_n2_m20.npz'), msa=msa, coords=c.numpy(), confs=f.numpy())
    print('synthetic L=100 N=96 n2m20 mean conf', float(f.mean()))


if __name__ == '__main__':
    if '--synthetic' in sys.argv:
        synthetic_golden()                                   # python oracle/make_golden.py --synthetic
    else:
        main()
        synthetic_golden()
