"""CPU tests: the oracle restatement against the golden vectors produced by the reference itself
(oracle/make_golden.py), plus the oracle's internal consistency."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, needs_weights
from oracle import dmpfold_oracle as O


def _gold(name):
    return np.load(os.path.join(GOLDEN, name))


def test_encode_matches_reference_alnmat(pf10963):
    g = _gold('pf10963_n0_m0.npz')
    assert pf10963.dtype == np.uint8 and pf10963.shape == (252, 82)
    assert np.array_equal(pf10963, g['alnmat'])


def test_reweight_matches_reference(pf10963):
    g = _gold('pf10963_n0_m0.npz')
    w = O.reweight(O.one_hot_msa(torch.from_numpy(pf10963)))
    assert np.array_equal(w.numpy(), g['w'])


def test_dca_features_match_reference(pf10963):
    g = _gold('pf10963_n0_m0.npz')
    f = O.msa_features(torch.from_numpy(pf10963)).permute(2, 0, 1)
    np.testing.assert_allclose(f[:, ::9, ::7].numpy(), g['x2_sub'][:442], rtol=0, atol=2e-5)
    np.testing.assert_allclose(f[441].numpy(), g['x2_apc'], rtol=0, atol=2e-5)


@needs_weights
def test_forward_n0_m0_matches_reference(oracle, pf10963):
    g = _gold('pf10963_n0_m0.npz')
    taps = {}
    coords, conf = oracle.fold(pf10963, iterations=0, minsteps=0, taps=taps)
    np.testing.assert_allclose(taps['mat1d'].numpy(), g['mat1d'], atol=1e-5)
    np.testing.assert_allclose(taps['stem'][0][:, ::8, ::8].numpy(), g['stem_sub'], atol=1e-4)
    np.testing.assert_allclose(taps['block16'][0][:, ::8, ::8].numpy(), g['block16_sub'], atol=2e-3)
    np.testing.assert_allclose(taps['head'][0].numpy(), g['head'], atol=2e-3)
    np.testing.assert_allclose(taps['mds'][0].numpy(), g['mds'], atol=2e-3)
    assert O.kabsch_rmsd(coords[:, 1].numpy(), g['coords'][:, 1]) < 1e-4
    np.testing.assert_allclose(conf.numpy(), g['confs'], atol=1e-5)


@needs_weights
def test_forward_n2_m20_matches_reference(oracle, pf10963):
    g = _gold('pf10963_n2_m20.npz')
    coords, conf = oracle.fold(pf10963, iterations=2, minsteps=20)
    assert O.kabsch_rmsd(coords[:, 1].numpy(), g['coords'][:, 1]) < 2e-4
    np.testing.assert_allclose(coords.numpy(), g['coords'], atol=2e-3)
    np.testing.assert_allclose(conf.numpy(), g['confs'], atol=1e-4)


@needs_weights
def test_forward_synthetic_l100_matches_reference(oracle, pf10963):
    """Second pin at another size: the structured synthetic alignment of the halo-sharded tests, folded by the
    reference itself (oracle/make_golden.py --synthetic).  Also pins the generator: oracle copy == product copy == the
    codes the reference parsed back from the .aln text."""
    from dmpfold2_b200.synth import synth_msa_structured
    g = _gold('synth_l100_n96_s3_n2_m20.npz')
    msa = O.synth_msa_structured(pf10963, 100, 96, 3)
    assert np.array_equal(msa, g['msa']) and np.array_equal(synth_msa_structured(pf10963, 100, 96, 3), g['msa'])
    coords, conf = oracle.fold(msa, iterations=2, minsteps=20)
    rmsd = O.kabsch_rmsd(coords[:, 1].numpy(), g['coords'][:, 1])
    print(f'oracle vs reference on the synthetic L=100 target: CA-RMSD {rmsd:.2e} A')
    assert rmsd < 2e-4
    np.testing.assert_allclose(conf.numpy(), g['confs'], atol=1e-4)


@needs_weights
def test_template_and_single_sequence_match_reference(oracle, pf10963, tmp_path):
    g = _gold('pf10963_tmpl_n1_m10.npz')
    pdb = tmp_path / 'tmpl.pdb'
    pdb.write_text(str(g['pdb_text']))
    ca = O.read_template_ca(str(pdb))
    assert ca.shape == (82, 3)
    coords, conf = oracle.fold(pf10963, template_ca=ca, iterations=1, minsteps=10)
    assert O.kabsch_rmsd(coords[:, 1].numpy(), g['coords'][:, 1]) < 2e-4
    g1 = _gold('pf10963_single_n1_m0.npz')
    coords, conf = oracle.fold(pf10963[:1], iterations=1, minsteps=0)
    assert O.kabsch_rmsd(coords[:, 1].numpy(), g1['coords'][:, 1]) < 2e-4
    np.testing.assert_allclose(conf.numpy(), g1['confs'], atol=1e-4)


def test_gru_restatement_equals_aten_gru(oracle, pf10963):
    msa = torch.from_numpy(pf10963[:24, :20])
    x = torch.nn.functional.one_hot(msa.long(), 22).float()
    a = O.gru_stack(x, oracle.sd, 'vgru', 2, False)[-1]
    b = oracle.vgru_last(msa)
    assert (a - b).abs().max() < 1e-5
    v = torch.randn(20, 512, generator=torch.Generator().manual_seed(1))
    a = O.gru_stack(v.unsqueeze(1), oracle.sd, 'hgru', 2, True)[:, 0]
    b = oracle.hgru_out(v)
    assert (a - b).abs().max() < 1e-5


def test_refine_properties():
    g = torch.Generator().manual_seed(3)
    ca = torch.cumsum(torch.randn(40, 3, generator=g) * 2.2, dim=0)
    out = O.refine_coords(ca, 200)
    bonds = (out[1:] - out[:-1]).norm(dim=1)
    assert (bonds - 3.78).abs().max() < 0.2                 # bond springs pull neighbours to 3.78 A
    assert torch.equal(O.refine_coords(ca, 0), ca)


def test_backbone_is_rigid_motion_equivariant():
    g = torch.Generator().manual_seed(5)
    ca = torch.cumsum(torch.randn(1, 30, 3, generator=g) * 2.2, dim=1)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    t = torch.tensor([1.0, -2.0, 3.0])
    a = O.calpha_to_main_chain(ca) @ q.t() + t
    b = O.calpha_to_main_chain(ca @ q.t() + t)
    assert (a - b).abs().max() < 1e-3
    out = O.calpha_to_main_chain(ca).view(30, 5, 3)
    assert torch.equal(out[:, 1], ca[0])                     # CA passes through unchanged, atom index 1


def test_canonical_sign_rule():
    v = torch.tensor([[0.1, -0.9], [-0.7, 0.2], [0.7, 0.3]])
    s = O.canonical_sign(v)
    assert s[1, 0] > 0 and s[0, 1] > 0                       # lowest index wins the |0.7| tie in column 0
    assert torch.allclose(s.abs(), v.abs())


def test_synthetic_generators_are_deterministic(pf10963):
    a = O.synth_msa_random(50, 20, 7)
    assert np.array_equal(a, O.synth_msa_random(50, 20, 7)) and a.shape == (20, 50) and a.max() <= 21
    assert (a[0] < 20).all()
    s = O.synth_msa_structured(pf10963, 120, 64, 3)
    assert s.shape == (64, 120) and np.array_equal(s, O.synth_msa_structured(pf10963, 120, 64, 3))
    assert np.array_equal(s[0, :82], pf10963[0])


def test_kabsch():
    g = np.random.default_rng(0)
    a = g.normal(size=(25, 3))
    q, _ = np.linalg.qr(g.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    assert O.kabsch_rmsd(a, a @ q.T + 3.0) < 1e-7
    assert O.kabsch_rmsd(a, a) < 1e-6
