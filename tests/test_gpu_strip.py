"""Halo-sharded fold (BASELINE.json configs[4]): one target, every L x L map split in row strips over several
ranks; conv halo rows, InstanceNorm sums and head rows move between the ranks' windows by peer stores + epoch flags
(csrc/strip.cu).

Every rank is its own PROCESS with its own engine (tests/strip_worker.py), exactly as under torchrun: the windows
are shared through CUDA IPC handles exchanged once over a gloo group.  Ranks use GPU (rank mod #GPUs), so on a
one-GPU box they time-slice one device and on a multi-GPU box the stores cross NVLink.  The sharded arithmetic equals
the single-engine fold except for the order in which the InstanceNorm sums are folded (fp64), so the agreement bar
is far below the 1e-3 A parity tolerance.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, needs_weights
from oracle import dmpfold_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def strip_fold(tmp_path, msa, world, runs, mode='f16f8', tmpl=None, timeout=600):
    """Fold `msa` on `world` ranks for every (iterations, minsteps) in runs -> per rank a list of (coords, confs)."""
    job = {'msa': np.ascontiguousarray(msa, dtype=np.uint8), 'runs': np.asarray(runs, dtype=np.int64), 'mode': np.asarray(mode)}
    if tmpl is not None:
        job['tmpl'] = np.ascontiguousarray(tmpl, dtype=np.float32)
    job_path = os.path.join(str(tmp_path), f'job_w{world}.npz')
    np.savez(job_path, **job)
    out_dir = os.path.join(str(tmp_path), f'out_w{world}')
    os.makedirs(out_dir, exist_ok=True)
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', 'strip_worker.py'), job_path, out_dir],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    try:
        for p in procs:
            out, _ = p.communicate(timeout=timeout)
            logs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, p in enumerate(procs):
        assert p.returncode == 0, f'rank {r} failed:\n{logs[r][-3000:]}'
    res = []
    for r in range(world):
        z = np.load(os.path.join(out_dir, f'rank{r}.npz'))
        res.append([(z[f'coords{i}'], z[f'confs{i}']) for i in range(len(runs))])
    return res


@needs_weights
@pytest.mark.parametrize('world', [2, 3])
def test_strip_fold_matches_single_engine(state_dict, pf10963, tmp_path, world):
    from dmpfold2_b200.engine import Engine
    from dmpfold2_b200.synth import synth_msa_structured
    msa = synth_msa_structured(pf10963, 100, 96, 3)                  # L = 100: strips of 56+44 / 40+40+20 rows
    runs = [(0, 0), (2, 20)]
    e = Engine(state_dict, 0)
    ref = [e.fold_host(msa, None, n, m) for n, m in runs]
    e.close()
    res = strip_fold(tmp_path, msa, world, runs)
    for r in range(world):
        for i, (c, f) in enumerate(res[r]):
            rmsd = O.kabsch_rmsd(c[:, 1], ref[i][0][:, 1])
            print(f'world {world} rank {r} run {runs[i]}: CA-RMSD vs single engine {rmsd:.2e} A, max|dconf| {np.abs(f - ref[i][1]).max():.2e}')
            assert rmsd < 1e-4 and np.abs(f - ref[i][1]).max() < 1e-4
            assert np.array_equal(c, res[0][i][0]) and np.array_equal(f, res[0][i][1])      # every rank returns the same bits


@needs_weights
@pytest.mark.parametrize('mode', ['f16f8', 'f16x3'])
def test_strip_fold_matches_reference_golden(pf10963, tmp_path, mode):
    g = np.load(os.path.join(GOLDEN, 'pf10963_n2_m20.npz'))
    res = strip_fold(tmp_path, pf10963, 2, [(2, 20)], mode=mode)      # L = 82: 48 + 34 rows
    c, f = res[1][0]
    rmsd = O.kabsch_rmsd(c[:, 1], g['coords'][:, 1])
    print(f'{mode}: 2-strip fold vs reference golden: CA-RMSD {rmsd:.2e} A')
    assert rmsd <= 1e-3 and np.abs(f - g['confs']).max() < 2e-3


@needs_weights
def test_strip_template_and_odd_length(state_dict, pf10963, tmp_path):
    from dmpfold2_b200.engine import Engine
    msa = np.ascontiguousarray(pf10963[:64, :77])                    # odd L: head rows are not 16-byte aligned
    e = Engine(state_dict, 0)
    c0, f0 = e.fold_host(msa, None, 0, 0)
    ref_c, ref_f = e.fold_host(msa, c0[:, 1].copy(), 1, 10)
    e.close()
    res = strip_fold(tmp_path, msa, 2, [(1, 10)], tmpl=c0[:, 1])
    for r in range(2):
        c, f = res[r][0]
        assert O.kabsch_rmsd(c[:, 1], ref_c[:, 1]) < 1e-4 and np.abs(f - ref_f).max() < 1e-4


@needs_weights
def test_strip_world1_and_bad_use(state_dict, pf10963):
    from dmpfold2_b200.engine import Engine, Dmp2Error
    e = Engine(state_dict, 0)
    try:
        with pytest.raises(Dmp2Error):
            e.fold_strip(torch.from_numpy(pf10963), None, 0, 0)        # no window yet
        with pytest.raises(Dmp2Error):
            e.strip_setup(0, 8, 40)                                    # 40 rows cannot feed 8 strips
        ref_c, ref_f = e.fold_host(pf10963, None, 1, 10)
        handle, win = e.strip_setup(0, 1, 82, 252)
        assert len(handle) == 64 and win != 0
        with pytest.raises(Dmp2Error):
            e.strip_setup(0, 1, 82)                                    # window exists
        e.strip_attach_local([win])
        c, f = e.fold_strip_host(pf10963, None, 1, 10)                 # one strip = the whole map, through the window path
        assert O.kabsch_rmsd(c[:, 1], ref_c[:, 1]) < 1e-5 and np.abs(f - ref_f).max() < 1e-5
        with pytest.raises(Dmp2Error):
            e.fold_strip_host(pf10963[:, :80], None, 0, 0)             # windows are sized for one L
        e.strip_detach()
    finally:
        e.close()
