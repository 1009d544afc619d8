"""GPU parity tests, stage by stage: every C-ABI stage entry point against the CPU oracle on seeded inputs
(teacher-forced: each stage gets the oracle's inputs, so errors do not compound)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, needs_weights
from oracle import dmpfold_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng(state_dict):
    from dmpfold2_b200.engine import Engine
    e = Engine(state_dict, 0)
    yield e
    e.close()


def _rand_msa(n, l, seed):
    return O.synth_msa_random(l, n, seed)


def _nhwc(x_nchw):
    return x_nchw[0].permute(1, 2, 0).contiguous()


def _nchw(x_nhwc):
    return x_nhwc.permute(2, 0, 1).unsqueeze(0).contiguous()


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('n,l,seed', [(37, 45, 1), (252, 82, None), (130, 31, 2), (1, 16, 3)])
def test_reweight_bit_exact(eng, pf10963, n, l, seed):
    msa = pf10963 if seed is None else _rand_msa(n, l, seed)
    ref = O.reweight(O.one_hot_msa(torch.from_numpy(msa)))
    got = eng.reweight(msa).cpu()
    assert torch.equal(got, ref)


def test_dca_features(eng, pf10963):
    ref = O.msa_features(torch.from_numpy(pf10963))
    got = eng.dca(pf10963).cpu()
    assert got.shape == ref.shape == (82, 82, 442)
    assert (got - ref).abs().max() < 3e-4 * max(1.0, float(ref.abs().max()))
    # ragged sizes (N, 21L not multiples of the tile sizes) and the single-sequence branch (predict.py:139)
    msa = _rand_msa(53, 23, 11)
    ref = O.msa_features(torch.from_numpy(msa))
    got = eng.dca(msa).cpu()
    assert (got - ref).abs().max() < 3e-4 * max(1.0, float(ref.abs().max()))
    assert float(eng.dca(msa[:1]).abs().max()) == 0.0


def test_gemm_paths_tensor_core_vs_cuda_core(state_dict, oracle, pf10963, monkeypatch):
    """The dense contractions of the MSA features (Gram / covariance / Woodbury products) and the stem GEMM run on the
    tcgen05 pipeline by default; DMP2_GEMM=ffma keeps them on the CUDA-core fp32 GEMM.  Both must agree with the oracle
    (Woodbury branch, direct-covariance branch, ragged sizes) and with each other."""
    from dmpfold2_b200.engine import Engine
    cases = [pf10963, _rand_msa(53, 23, 11), O.synth_msa_structured(pf10963, 30, 700, 5)]       # last: N >= 21 L (direct branch)
    outs = {}
    for mode in ('tc', 'ffma'):
        if mode == 'ffma':
            monkeypatch.setenv('DMP2_GEMM', 'ffma')
        e = Engine(state_dict, 0)
        try:
            outs[mode] = [e.dca(m).cpu() for m in cases]
            m1 = torch.tanh(torch.randn(33, 512, generator=torch.Generator().manual_seed(3)))
            outs[mode].append(e.resnet_pass(m1, outs[mode][1].new_zeros(33, 33, 442).normal_(generator=torch.Generator().manual_seed(4)),
                                            torch.full((33, 33), -1.0)).cpu())
        finally:
            e.close()
    for i, m in enumerate(cases):
        ref = O.msa_features(torch.from_numpy(m))
        scale = max(1.0, float(ref.abs().max()))
        for mode in ('tc', 'ffma'):
            assert (outs[mode][i] - ref).abs().max() < 3e-4 * scale, (mode, i, float((outs[mode][i] - ref).abs().max()))
        assert (outs['tc'][i] - outs['ffma'][i]).abs().max() < 1e-4 * scale
    assert _rel(outs['tc'][3], outs['ffma'][3]) < 1e-4


def test_vgru(eng, oracle, pf10963):
    for msa in (pf10963[:60, :41], _rand_msa(9, 70, 4)):
        ref = oracle.vgru_last(torch.from_numpy(msa))
        got = eng.vgru(msa).cpu()
        assert (got - ref).abs().max() < 5e-5


def test_hgru(eng, oracle):
    g = torch.Generator().manual_seed(7)
    for l in (19, 82):
        v = torch.tanh(torch.randn(l, 512, generator=g))
        ref = oracle.hgru_out(v)
        got = eng.hgru(v).cpu()
        assert (got - ref).abs().max() < 5e-5


def _conv_ref(oracle, block, x_nchw):
    sd = oracle.sd
    y = F.conv2d(x_nchw, sd[f'resnet.{block}.layer1.lin.weight'], sd[f'resnet.{block}.layer1.lin.bias'], padding=2)
    return y.view(1, 128, 4, y.shape[2], y.shape[3]).max(dim=2)[0]


@pytest.mark.parametrize('mode,tol', [('ffma', 2e-5), ('f16x3', 2e-5), ('f16f8', 3e-4), ('f16', 3e-3)])
@pytest.mark.parametrize('l', [16, 27, 82])
def test_conv5_maxout(eng, oracle, mode, tol, l):
    g = torch.Generator().manual_seed(100 + l)
    x = torch.randn(1, 128, l, l, generator=g) * 3
    eng.set_conv_mode(mode)
    try:
        for block in (1, 16):
            ref = _nhwc(_conv_ref(oracle, block, x))
            got = eng.conv5_maxout(block, _nhwc(x)).cpu()
            assert _rel(got, ref) < tol, (mode, l, block, _rel(got, ref))
    finally:
        eng.set_conv_mode('ffma')


@pytest.mark.parametrize('mode,tol', [('ffma', 5e-5), ('f16x3', 5e-5), ('f16f8', 5e-4)])
def test_resblock(eng, oracle, mode, tol):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(1, 128, 33, 33, generator=g) * 2
    eng.set_conv_mode(mode)
    try:
        for block in (1, 9):
            ref = _nhwc(O.resnet_block(x, oracle.sd, block))
            got = eng.resblock(block, _nhwc(x)).cpu()
            assert _rel(got, ref) < tol, (mode, block, _rel(got, ref))
    finally:
        eng.set_conv_mode('ffma')


def test_cse_gate_is_the_weights_only_constant(oracle):
    # fact 8 of SURVEY.md: avgpool(InstanceNorm_affine(x)) == beta, so the cSE gate does not depend on x
    g = torch.Generator().manual_seed(2)
    y = torch.randn(1, 128, 20, 20, generator=g)
    sd = oracle.sd
    yn = F.instance_norm(y, weight=sd['resnet.3.layer1.norm.weight'], bias=sd['resnet.3.layer1.norm.bias'])
    assert (yn.mean(dim=(2, 3))[0] - sd['resnet.3.layer1.norm.bias']).abs().max() < 1e-5


@needs_weights
@pytest.mark.parametrize('mode,tol', [('ffma', 1e-3), ('f16x3', 1e-3), ('f16f8', 2e-3)])
def test_resnet_pass_teacher_forced(eng, oracle, pf10963, mode, tol):
    taps = {}
    oracle.fold(pf10963, iterations=0, minsteps=0, taps=taps)
    x2 = taps['x2'][0]                                   # (443, L, L)
    feat = x2[:442].permute(1, 2, 0).contiguous()
    eng.set_conv_mode(mode)
    try:
        got = eng.resnet_pass(taps['mat1d'].t().contiguous(), feat, x2[442]).cpu()
    finally:
        eng.set_conv_mode('ffma')
    ref = taps['head'][0]
    assert _rel(got, ref) < tol, _rel(got, ref)


@needs_weights
@pytest.mark.parametrize('gemm', ['tc', 'ffma'])
def test_stem_and_head_teacher_forced(state_dict, oracle, pf10963, gemm, monkeypatch):
    # SURVEY 8(b) stage list: dmp2_stem (network.py:194 on the 955-channel input of :227-229) and dmp2_head (network.py:207)
    from dmpfold2_b200.engine import Engine
    monkeypatch.setenv('DMP2_GEMM', gemm)
    taps = {}
    oracle.fold(pf10963, iterations=0, minsteps=0, taps=taps)
    x2 = taps['x2'][0]
    feat = x2[:442].permute(1, 2, 0).contiguous()
    e = Engine(state_dict, 0)
    try:
        got = e.stem(taps['mat1d'].t().contiguous(), feat, x2[442]).cpu()
        ref = _nhwc(taps['stem'])
        assert _rel(got, ref) < 1e-4, _rel(got, ref)
        head = e.head(_nhwc(taps['block16'])).cpu()
        assert _rel(head, taps['head'][0]) < 1e-5, _rel(head, taps['head'][0])
        # the recycling input (network.py:264-270): channel 954 becomes the distance map of the previous coordinates
        ca = taps['ca0']
        dmap = torch.clamp((ca.unsqueeze(0) - ca.unsqueeze(1)).pow(2).sum(dim=2), min=1e-8).sqrt()
        m1 = taps['mat1d']
        resinp = torch.cat(((m1.unsqueeze(1) * m1.unsqueeze(2)).unsqueeze(0), x2[None, :442], dmap[None, None]), dim=1)
        sd = oracle.sd
        ref2 = O.maxout_norm(resinp, sd['resnet.0.lin.weight'], sd['resnet.0.lin.bias'], sd['resnet.0.norm.weight'],
                             sd['resnet.0.norm.bias'], 3, 0)
        got2 = e.stem(m1.t().contiguous(), feat, dmap).cpu()
        assert _rel(got2, _nhwc(ref2)) < 1e-4, _rel(got2, _nhwc(ref2))
    finally:
        e.close()


@needs_weights
def test_head_mds_on_golden_head(eng):
    g = np.load(os.path.join(GOLDEN, 'pf10963_n0_m0.npz'))
    head = torch.from_numpy(g['head'])
    conf_r, m_r, mds_r = O.head_to_mds(head.unsqueeze(0))
    conf, m, mds = eng.head_mds(head)
    assert (conf.cpu() - conf_r[0]).abs().max() < 1e-5
    assert torch.equal(m.cpu(), m_r[0])                  # same fp32 operation order -> bit exact
    assert (mds.cpu() - mds_r[0]).abs().max() < 3e-3     # oracle eigenvectors are fp32 LAPACK
    assert (mds.cpu() - torch.from_numpy(g['mds'])).abs().max() < 3e-3


@pytest.mark.parametrize('l', [8, 33, 150, 520, 700, 1100])      # cluster of 8, of 16, whole-GPU tridiagonalisation
def test_eig_top8_against_fp64_eigh(eng, l):
    g = torch.Generator().manual_seed(l)
    q, _ = torch.linalg.qr(torch.randn(l, l, generator=g, dtype=torch.float64))
    lam = torch.cat((torch.tensor([-5000.0]), torch.linspace(-50, 40, l - 5, dtype=torch.float64),
                     torch.tensor([100.0, 100.5, 250.0, 900.0])))[:l]
    m = (q * lam) @ q.t()
    m = ((m + m.t()) / 2).float()
    w, v = torch.linalg.eigh(m.double())
    v = O.canonical_sign(v)
    vals, vecs = eng.eig_top8(m)
    assert (vals.cpu().double() - w[-8:]).abs().max() < 1e-3
    assert (vecs.cpu().double() - v[:, -8:]).abs().max() < 2e-4


def test_coord_gru(eng, oracle):
    g = torch.Generator().manual_seed(9)
    l = 41
    m1 = torch.tanh(torch.randn(512, l, generator=g))
    mds = torch.randn(l, 8, generator=g) * 5
    ref = oracle.coord_head(m1, mds)
    got = eng.coord_gru(m1.t().contiguous(), mds).cpu()
    assert (got - ref).abs().max() < 2e-4


@pytest.mark.parametrize('l,steps', [(12, 1), (82, 100), (300, 20), (1100, 3)])
def test_refine(eng, l, steps):
    g = torch.Generator().manual_seed(l)
    ca = torch.cumsum(torch.randn(l, 3, generator=g) * 1.6, dim=0)
    ref = O.refine_coords(ca, steps)
    got = eng.refine(ca, steps).cpu()
    assert (got - ref).abs().max() < 5e-4


@pytest.mark.parametrize('l', [8, 82, 513])
def test_backbone(eng, l):
    g = torch.Generator().manual_seed(l)
    ca = torch.cumsum(torch.randn(l, 3, generator=g) * 2.2, dim=0)
    ref = O.calpha_to_main_chain(ca.unsqueeze(0))[0].view(l, 5, 3)
    got = eng.backbone(ca).cpu()
    assert (got - ref).abs().max() < 1e-4
    assert torch.equal(got[:, 1], ca)


def test_bad_arguments_return_errors(eng):
    from dmpfold2_b200.engine import Dmp2Error
    with pytest.raises(Dmp2Error):
        eng.fold(torch.zeros(4, 5, dtype=torch.uint8))           # L < 8
    with pytest.raises(Dmp2Error):
        eng.conv5_maxout(17, torch.zeros(8, 8, 128))


@pytest.mark.parametrize('mode,tol', [('f16x3', 5e-5), ('f16', 2e-3)])
@pytest.mark.parametrize('m,n,k,chunk', [(128, 512, 64, 0), (300, 512, 256, 128), (1000, 512, 3200, 128), (777, 384, 1024, 256),
                                         (200, 36, 128, 64)])
def test_tensor_core_gemm_core(eng, mode, tol, m, n, k, chunk):
    """TMA -> tcgen05 -> TMEM pipeline self-test on a plain GEMM (descriptors, swizzle, barriers, the double-buffered
    chunk accumulators and the register running sums), ragged M and N included."""
    g = torch.Generator().manual_seed(m + k)
    a = torch.randn(m, k, generator=g)
    b = torch.randn(n, k, generator=g) * 0.1
    ref = a.double() @ b.double().t()
    got = eng.gemm_tn_test(a, b, mode, chunk).cpu().double()
    assert _rel(got, ref) < tol, _rel(got, ref)


def test_two_level_accumulation_matches_fp32_cpu_gemm(eng):
    """The point of the per-tap accumulation chains: with chains of K=128 summed in fp32 registers the tensor-core
    GEMM is as close to the exact product as torch's fp32 CPU GEMM (within 2x), while ONE chain over K=3200 is an order
    of magnitude further away (tcgen05 accumulators truncate on every MMA)."""
    g = torch.Generator().manual_seed(5)
    a = torch.randn(640, 3200, generator=g)
    b = torch.randn(512, 3200, generator=g)
    exact = a.double() @ b.double().t()

    def resid(y):                                # error after removing a uniform scale (InstanceNorm absorbs that part)
        e = y.double() - exact
        e = e - float((e * exact).sum() / (exact * exact).sum()) * exact
        return float(e.pow(2).mean().sqrt() / exact.pow(2).mean().sqrt())
    cpu = resid(a @ b.t())
    chained = resid(eng.gemm_tn_test(a, b, 'f16x3', 128).cpu())
    single = resid(eng.gemm_tn_test(a, b, 'f16x3', 0).cpu())
    print(f'fp32 CPU GEMM {cpu:.2e}, tcgen05 chains of 128: {chained:.2e}, one chain: {single:.2e}')
    assert chained < 2 * cpu and single > 3 * chained


def test_vgru_long_wavefront(eng, oracle, pf10963):
    """N much larger than the 3-stage wavefront depth, L spanning two 128-row tiles with a ragged tail."""
    msa = O.synth_msa_structured(pf10963, 150, 203, 5)
    ref = oracle.vgru_last(torch.from_numpy(msa))
    got = eng.vgru(msa).cpu()
    assert (got - ref).abs().max() < 1e-4
    one = eng.vgru(msa[:1]).cpu()                      # N = 1: shortest wavefront
    assert (one - oracle.vgru_last(torch.from_numpy(msa[:1]))).abs().max() < 1e-5


@pytest.mark.parametrize('cluster,sms', [('pair', '0'), ('pair', '6'), ('1', '3'), ('2', '20')])
def test_conv_dynamic_schedule_is_bit_identical(state_dict, oracle, pf10963, cluster, sms, monkeypatch):
    """Throughput mode's dynamic unit schedule (the clusters claim units from a device counter; dmp2_set_conv_dynamic):
    every unit is computed exactly as in the static schedule whichever cluster runs it, so the conv output must be
    BIT-identical, launch after launch (the counter is never reset: every launch owns the next range of values);
    the fused InstanceNorm sums (fp64, run-dependent order) must give the same block output to fp32 rounding."""
    from dmpfold2_b200.engine import Engine
    monkeypatch.setenv('DMP2_CONV_CLUSTER', cluster)
    monkeypatch.setenv('DMP2_CONV_SMS', sms)
    es, ed = Engine(state_dict, 0), Engine(state_dict, 0)
    ed.set_conv_dynamic(True)
    try:
        g = torch.Generator().manual_seed(78)
        for l in (19, 82, 150):
            x = _nhwc(torch.randn(1, 128, l, l, generator=g) * 3)
            for mode in ('f16x3', 'f16f8', 'f16'):
                es.set_conv_mode(mode)
                ed.set_conv_mode(mode)
                want = es.conv5_maxout(5, x)
                for rep in range(3):
                    assert torch.equal(ed.conv5_maxout(5, x), want), (cluster, sms, mode, l, rep)
                if mode != 'f16':
                    a, b = es.resblock(5, x), ed.resblock(5, x)
                    assert _rel(b.cpu(), a.cpu()) < 1e-6, (cluster, sms, mode, l)
        if cluster == 'pair' and sms == '0':                 # and a whole fold
            msa = O.synth_msa_structured(pf10963, 120, 64, 5)
            c0, f0 = es.fold_host(msa, None, 2, 20)
            c1, f1 = ed.fold_host(msa, None, 2, 20)
            assert O.kabsch_rmsd(c0[:, 1], c1[:, 1]) < 1e-5 and np.abs(f0 - f1).max() < 1e-5
    finally:
        es.close()
        ed.close()


@pytest.mark.parametrize('cluster,sms', [('1', '0'), ('2', '20'), ('1', '3')])
def test_conv_cluster_variants(state_dict, oracle, cluster, sms, monkeypatch):
    """The persistent conv kernel without weight multicast (cluster of 1), and on a restricted number of SMs (every
    CTA then walks through many work units: exercises the tile scheduler, both TMEM chunk buffers' phase wrap-around
    and the cross-unit InstanceNorm sums), must agree with the oracle exactly like the default (2-CTA multicast)."""
    from dmpfold2_b200.engine import Engine
    monkeypatch.setenv('DMP2_CONV_CLUSTER', cluster)
    monkeypatch.setenv('DMP2_CONV_SMS', sms)
    e = Engine(state_dict, 0)
    try:
        g = torch.Generator().manual_seed(77)
        for l in (19, 82):                                   # odd tile counts exercise the padded dummy tiles
            x = torch.randn(1, 128, l, l, generator=g) * 3
            ref = _nhwc(_conv_ref(oracle, 5, x))
            for mode, tol in (('f16x3', 2e-5), ('f16f8', 3e-4), ('f16', 3e-3)):
                e.set_conv_mode(mode)
                got = e.conv5_maxout(5, _nhwc(x)).cpu()
                assert _rel(got, ref) < tol, (cluster, mode, l, _rel(got, ref))
    finally:
        e.close()
