"""GPU end-to-end parity: the whole fold against the golden vectors the reference produced (oracle B) and
against the oracle at other sizes.  Bar: CA-RMSD <= 1e-3 A (BASELINE.json north_star), atom order N,CA,C,O,CB."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, needs_weights
from oracle import dmpfold_oracle as O

pytestmark = pytest.mark.gpu
TOL_RMSD = 1e-3          # Angstrom, the north_star tolerance


@pytest.fixture(scope='module')
def eng(state_dict):
    from dmpfold2_b200.engine import Engine
    e = Engine(state_dict, 0)
    yield e
    e.close()


def _check(coords, conf, g, tol=TOL_RMSD):
    coords, conf = np.asarray(coords), np.asarray(conf)
    assert coords.shape == g['coords'].shape and np.isfinite(coords).all()
    rmsd = O.kabsch_rmsd(coords[:, 1], g['coords'][:, 1])
    rmsd_all = O.kabsch_rmsd(coords.reshape(-1, 3), g['coords'].reshape(-1, 3))
    print(f'CA-RMSD {rmsd:.2e} A, all-atom {rmsd_all:.2e} A, max|dconf| {np.abs(conf - g["confs"]).max():.2e}')
    assert rmsd <= tol and rmsd_all <= 2 * tol, (rmsd, rmsd_all)
    assert np.abs(conf - g['confs']).max() < 2e-3
    return rmsd


@needs_weights
@pytest.mark.parametrize('mode', ['ffma', 'f16x3', 'f16f8'])
@pytest.mark.parametrize('n,m', [(0, 0), (2, 20), (10, 100)])
def test_pf10963_matches_reference(eng, pf10963, mode, n, m):
    g = np.load(os.path.join(GOLDEN, f'pf10963_n{n}_m{m}.npz'))
    eng.set_conv_mode(mode)
    coords, conf = eng.fold(torch.from_numpy(pf10963), None, n, m)
    _check(coords.cpu().numpy(), conf.cpu().numpy(), g)
    c2, f2 = eng.fold_host(pf10963, None, n, m)                    # host-buffer entry point, same result
    assert np.array_equal(c2, coords.cpu().numpy()) and np.array_equal(f2, conf.cpu().numpy())


@needs_weights
def test_template_and_single_sequence(eng, pf10963, tmp_path):
    eng.set_conv_mode('f16x3')
    g = np.load(os.path.join(GOLDEN, 'pf10963_tmpl_n1_m10.npz'))
    pdb = tmp_path / 't.pdb'
    pdb.write_text(str(g['pdb_text']))
    ca = O.read_template_ca(str(pdb))
    coords, conf = eng.fold_host(pf10963, ca, 1, 10)
    _check(coords, conf, g)
    g1 = np.load(os.path.join(GOLDEN, 'pf10963_single_n1_m0.npz'))
    coords, conf = eng.fold_host(pf10963[:1], None, 1, 0)
    _check(coords, conf, g1)


@pytest.mark.parametrize('mode', ['f16x3', 'f16f8'])
@pytest.mark.parametrize('l,n,seed', [(57, 40, 1), (164, 96, 2)])
def test_structured_synthetic_vs_oracle(eng, oracle, pf10963, l, n, seed, mode):
    msa = O.synth_msa_structured(pf10963, l, n, seed)
    ref_c, ref_f = oracle.fold(msa, iterations=1, minsteps=10)
    eng.set_conv_mode(mode)
    coords, conf = eng.fold_host(msa, None, 1, 10)
    _check(coords, conf, {'coords': ref_c.numpy(), 'confs': ref_f.numpy()})


@needs_weights
def test_fold_is_deterministic_and_negative_counts_clamp(eng, pf10963):
    eng.set_conv_mode('f16x3')
    a = eng.fold_host(pf10963, None, 1, 5)
    b = eng.fold_host(pf10963, None, 1, 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = eng.fold_host(pf10963, None, -3, -1)                       # predict.py:121-122
    d = eng.fold_host(pf10963, None, 0, 0)
    assert np.array_equal(c[0], d[0])


@needs_weights
def test_python_api_and_cli(pf10963, tmp_path):
    from dmpfold2_b200 import aln_to_coords
    aln = os.path.join(GOLDEN, 'PF10963.aln')
    g = np.load(os.path.join(GOLDEN, 'pf10963_n0_m0.npz'))
    coords, confs, alnmat = aln_to_coords(aln, device='cuda:0', iterations=0, minsteps=0, return_alnmat=True)
    assert coords.is_cuda and coords.shape == (82, 5, 3) and confs.shape == (82,) and alnmat.dtype == np.uint8
    _check(coords.cpu().numpy(), confs.cpu().numpy(), g)
    coords, confs = aln_to_coords(aln, iterations=0, minsteps=0)   # default device string 'cpu' -> CPU tensors
    assert not coords.is_cuda
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bin', 'dmpfold'), '-i', aln, '-n', '0', '-m', '0'],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0].startswith('REMARK  CONF:  0.71') and out[-1] == 'END'
    n_gly = int((pf10963[0] == 7).sum())
    assert len(out) == 2 + 82 * 5 - n_gly
    with pytest.raises(RuntimeError):
        aln_to_coords(aln, template=os.path.join(GOLDEN, 'PF10963.aln'))   # no CA atoms -> size mismatch


def test_cfg3_shape_vs_oracle(eng, oracle, pf10963):
    """BASELINE.json configs[2] shape (L=150, N=512), 2 recycles + 20 minimiser steps, against the oracle."""
    msa = O.synth_msa_structured(pf10963, 150, 512, 7)
    ref_c, ref_f = oracle.fold(msa, iterations=2, minsteps=20)
    eng.set_conv_mode('f16f8')
    coords, conf = eng.fold_host(msa, None, 2, 20)
    _check(coords, conf, {'coords': ref_c.numpy(), 'confs': ref_f.numpy()})


def _geometry_ok(coords):
    ca = coords[:, 1]
    bonds = np.linalg.norm(ca[1:] - ca[:-1], axis=1)
    n_ca = np.linalg.norm(coords[:, 0] - ca, axis=1)
    assert np.isfinite(coords).all()
    assert bonds.min() > 2.5 and bonds.max() < 5.2, (bonds.min(), bonds.max())     # bond springs pull towards 3.78 A
    assert n_ca[1:].min() > 0.9 and n_ca[1:].max() < 2.1, (n_ca.min(), n_ca.max())


def test_cfg4_long_target_properties(eng, pf10963):
    """BASELINE.json configs[3] shape (L=1024, N=2048, template-seeded) at a bounded iteration count: the oracle
    needs ~1 h of CPU for this size, so the check is on size-independent properties -- finite output, chain
    geometry after the minimiser, bit-reproducibility, and that the template seed changes the result."""
    msa = O.synth_msa_structured(pf10963, 1024, 2048, 11)
    eng.set_conv_mode('f16f8')
    c0, f0 = eng.fold_host(msa, None, 0, 200)
    _geometry_ok(c0)
    assert f0.min() >= 0.0 and f0.max() <= 1.0
    c1, f1 = eng.fold_host(msa, c0[:, 1].copy(), 1, 200)              # template = own CA trace
    _geometry_ok(c1)
    c2, f2 = eng.fold_host(msa, c0[:, 1].copy(), 1, 200)
    assert np.array_equal(c1, c2) and np.array_equal(f1, f2)
    assert not np.array_equal(c1, c0)


def test_cfg5_size_single_gpu_properties(eng, pf10963):
    """BASELINE.json configs[4] image size (L=2048) on ONE GPU: with the 955/512-channel tensors never materialised
    the whole target needs ~45 GB, so no halo split is required for capacity.  Bounded (N=1024, one pass);
    size-independent properties only."""
    msa = O.synth_msa_structured(pf10963, 2048, 1024, 13)
    eng.set_conv_mode('f16f8')
    c0, f0 = eng.fold_host(msa, None, 0, 100)
    _geometry_ok(c0)
    assert c0.shape == (2048, 5, 3) and f0.shape == (2048,)
    assert f0.min() >= 0.0 and f0.max() <= 1.0


# (the headline configuration L=300 / N=1000 at its full 10 recycles + 100 minimiser steps, and the other full-length
#  BASELINE.json shapes, are checked in tests/test_gpu_parity_r2.py against fp32 + fp64 fixtures of the reference algorithm)


@pytest.mark.parametrize('l,n', [(8, 2), (9, 1), (17, 3)])
def test_minimal_sizes_vs_oracle(eng, oracle, pf10963, l, n):
    """Smallest legal shapes: L = 8 is the minimum (top-8 MDS embedding, network.py:250); N = 1 takes the zero-feature
    branch (predict.py:139)."""
    msa = np.ascontiguousarray(pf10963[:n, 20:20 + l])
    ref_c, ref_f = oracle.fold(msa, iterations=1, minsteps=5)
    eng.set_conv_mode('f16x3')
    coords, conf = eng.fold_host(msa, None, 1, 5)
    _check(coords, conf, {'coords': ref_c.numpy(), 'confs': ref_f.numpy()})


@needs_weights
def test_batch_api(pf10963, tmp_path):
    from dmpfold2_b200 import alns_to_coords, aln_to_coords
    aln = os.path.join(GOLDEN, 'PF10963.aln')
    short = tmp_path / 'short.aln'
    with open(aln) as fh:
        rows = [l.rstrip()[:40] for l in fh.readlines()[:30]]
    short.write_text('\n'.join(rows) + '\n')
    res = alns_to_coords([aln, str(short)], device='cuda:0', iterations=1, minsteps=5)
    assert len(res) == 2 and res[0][0].shape == (82, 5, 3) and res[1][0].shape == (40, 5, 3)
    c, f = aln_to_coords(str(short), device='cuda:0', iterations=1, minsteps=5)
    assert torch.equal(res[1][0], c.cpu()) and torch.equal(res[1][1], f.cpu())
    # throughput mode: three targets in flight on two streams give the same bits as one at a time
    res2 = alns_to_coords([aln, str(short), aln], device='cuda:0', iterations=1, minsteps=5, streams=2)
    assert len(res2) == 3 and torch.equal(res2[0][0], res[0][0]) and torch.equal(res2[1][1], res[1][1])
    assert torch.equal(res2[2][0], res[0][0])


@needs_weights
def test_shared_vgru_scan_is_bit_identical(state_dict, pf10963):
    """Throughput mode scans the alignment columns of several targets in ONE vgru call (the scan is independent per
    column) and hands every fold its slice (dmp2_set_vgru_input).  The slice must be bit-identical to the fold's own
    scan and so must the fold; the hand-over is one-shot (the next fold scans again)."""
    from dmpfold2_b200.engine import Engine
    from dmpfold2_b200.parallel import StreamPool
    from dmpfold2_b200.synth import synth_msa_structured
    msas = [torch.from_numpy(synth_msa_structured(pf10963, l, 48, 30 + l)).cuda() for l in (82, 150, 100, 64, 130)]
    e, scan = Engine(state_dict, 0), Engine(state_dict, 0)
    try:
        want = [e.fold(m, None, 1, 10) for m in msas]
        state = scan.vgru(torch.cat(msas[:3], dim=1).contiguous())
        assert torch.equal(state[82:232], e.vgru(msas[1]))
        c, f = e.fold(msas[1], None, 1, 10, vgru=state[82:232])
        assert torch.equal(c, want[1][0]) and torch.equal(f, want[1][1])
        c, f = e.fold(msas[3], None, 1, 10)                    # one-shot: this fold scans its own alignment
        assert torch.equal(c, want[3][0])
        with pytest.raises(ValueError):
            e.fold(msas[1], None, 1, 10, vgru=state[:100])
    finally:
        e.close()
        scan.close()
    # the pool: 5 targets on 2 streams, scans shared by (82+150+100), (64+130); static conv schedule -> same bits
    pool = StreamPool(state_dict, 0, streams=2, conv_dynamic=False, scan_rows=384)
    try:
        from dmpfold2_b200.parallel import scan_batches
        assert scan_batches([(48, int(m.shape[1])) for m in msas], 384) == [[0, 1, 2], [3, 4]]
        for threads in (True, False):
            got = pool.fold_all(msas, None, 1, 10, host_threads=threads)
            torch.cuda.synchronize()
            for (c, f), (wc, wf) in zip(got, want):
                assert torch.equal(c, wc) and torch.equal(f, wf)
    finally:
        pool.close()
    # default pool (dynamic conv schedule: InstanceNorm sums in a run-dependent fp64 order): same result to rounding
    pool = StreamPool(state_dict, 0, streams=3)
    try:
        got = pool.fold_all(msas, None, 1, 10)
        torch.cuda.synchronize()
        for (c, f), (wc, wf) in zip(got, want):
            assert O.kabsch_rmsd(c[:, 1].cpu().numpy(), wc[:, 1].cpu().numpy()) < 1e-5
    finally:
        pool.close()


@needs_weights
def test_fused_stats_variant(state_dict, pf10963, monkeypatch):
    """DMP2_FUSE_STATS=0: the InstanceNorm sums come from a separate k_in_stats pass instead of the conv epilogue (the
    default).  Same statistics up to the summation order, so the two folds must agree far below the parity tolerance."""
    from dmpfold2_b200.engine import Engine
    from dmpfold2_b200.synth import synth_msa_structured
    msa = synth_msa_structured(pf10963, 150, 64, 5)
    e0 = Engine(state_dict, 0)
    c0, f0 = e0.fold_host(msa, None, 2, 20)
    e0.close()
    monkeypatch.setenv('DMP2_FUSE_STATS', '0')
    e1 = Engine(state_dict, 0)
    c1, f1 = e1.fold_host(msa, None, 2, 20)
    n_fused = e1.launch_count
    e1.close()
    assert O.kabsch_rmsd(c1[:, 1], c0[:, 1]) < 1e-5 and np.abs(f1 - f0).max() < 1e-5
    assert np.isfinite(c1).all() and n_fused > 0


@needs_weights
@pytest.mark.parametrize('dynamic', [False, True])
def test_graph_replay_of_the_recycling_iterations(state_dict, pf10963, dynamic):
    """dmp2_set_graph: the recycling iterations (network.py:264-306) replayed from a CUDA graph captured once per
    (L, workspace, configuration) give the eager loop's result -- bit for bit with the static conv schedule -- and the
    launch bookkeeping, a change of L, a change of conv mode and the conv profiler keep working."""
    from dmpfold2_b200.engine import Engine
    e = Engine(state_dict, 0)
    try:
        e.set_conv_dynamic(dynamic)
        n0 = e.launch_count
        a = e.fold_host(pf10963, None, 4, 10)
        per_fold = e.launch_count - n0
        e.set_graph(True)
        b = e.fold_host(pf10963, None, 4, 10)              # captures the iteration, replays it 4 times
        c = e.fold_host(pf10963, None, 4, 10)              # replays the cached graph
        assert e.launch_count - n0 == 3 * per_fold         # a replay counts the launches it stands for
        if dynamic:                                        # which CTA sums which unit varies: fp64 statistics to rounding
            assert O.kabsch_rmsd(a[0][:, 1], b[0][:, 1]) < 1e-4 and O.kabsch_rmsd(a[0][:, 1], c[0][:, 1]) < 1e-4
        else:
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
        g = np.load(os.path.join(GOLDEN, 'pf10963_n10_m100.npz'))
        coords, conf = e.fold_host(pf10963, None, 10, 100)
        _check(coords, conf, g)
        # another alignment length and another conv mode re-capture
        msa2 = np.ascontiguousarray(pf10963[:90, :57])
        d1 = e.fold_host(msa2, None, 3, 5)
        e.set_conv_mode('f16x3')
        f1 = e.fold_host(msa2, None, 3, 5)
        e.set_graph(False)
        f0 = e.fold_host(msa2, None, 3, 5)
        e.set_conv_mode('f16f8')
        d0 = e.fold_host(msa2, None, 3, 5)
        if not dynamic:
            assert np.array_equal(d0[0], d1[0]) and np.array_equal(f0[0], f1[0])
        else:
            assert O.kabsch_rmsd(d0[0][:, 1], d1[0][:, 1]) < 1e-4 and O.kabsch_rmsd(f0[0][:, 1], f1[0][:, 1]) < 1e-4
        # conv profiler with the graph on: eager first pass (16 launches) + the graph's last replay (16 launches)
        e.set_graph(True)
        e.set_profile(True)
        e.fold_host(pf10963, None, 4, 10)
        n, ms = e.conv_profile()
        e.set_profile(False)
        print(f'conv launches timed with the graph on: {n}, mean {ms / max(n, 1) * 1e3:.1f} us')
        assert n in (16, 32) and 0.0 < ms / n < 5.0, (n, ms)   # 16: this driver does not time event-record nodes
    finally:
        e.close()
