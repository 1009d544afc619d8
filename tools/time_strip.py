"""Time the halo-sharded fold of ONE long target over the ranks of a torchrun job and check it against the
single-engine fold.   torchrun --nproc-per-node 4 tools/time_strip.py [L N iterations minsteps]
Device time, max over ranks; rank 0 prints.  Rank 0 also folds the target alone for the comparison."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmpfold2_b200.engine import Engine  # noqa: E402
from dmpfold2_b200.parallel import StripGroup, max_over_ranks  # noqa: E402
from dmpfold2_b200.predict import load_weights, read_aln, encode_aln  # noqa: E402
from dmpfold2_b200.synth import synth_msa_structured  # noqa: E402

L, N, n, m = (int(a) for a in (sys.argv[1:5] + ['2048', '3000', '10', '100'][len(sys.argv) - 1:]))
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dist.init_process_group('gloo', rank=rank, world_size=world)
dev = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(dev)
base = encode_aln(read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
msa = synth_msa_structured(base, L, N, 0)
eng = Engine(load_weights(None), dev)
grp = StripGroup(eng)


def kabsch_rmsd(a, b):
    a = a - a.mean(0); b = b - b.mean(0)
    u, s, vt = np.linalg.svd(a.T @ b)
    d = np.sign(np.linalg.det(u @ vt))
    r = u @ np.diag([1, 1, d]) @ vt
    return float(np.sqrt(((a @ r - b) ** 2).sum(1).mean()))


best, stages = 1e9, None
for rep in range(3):
    dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    c, f = grp.fold_host(msa, None, n, m)
    dt = max_over_ranks((time.perf_counter() - t) * 1e3)
    if rep > 0 and dt < best:
        best, stages = dt, eng.stage_times()
grp.close()
if rank == 0:
    print(f'halo-sharded fold L={L} N={N} n={n} m={m} on {world} GPU(s): {best:.1f} ms (max over ranks, host buffers in/out)', flush=True)
    print('  stage ms (rank 0):', {k: round(v, 1) for k, v in stages.items()}, flush=True)
    t = time.perf_counter()
    c1, f1 = eng.fold_host(msa, None, n, m)
    t = time.perf_counter()
    c1, f1 = eng.fold_host(msa, None, n, m)
    one = (time.perf_counter() - t) * 1e3
    print(f'  single engine: {one:.1f} ms -> speed-up {one / best:.2f}x; CA-RMSD sharded vs single {kabsch_rmsd(c[:, 1], c1[:, 1]):.2e} A, '
          f'max|dconf| {np.abs(f - f1).max():.2e}, mean conf {f.mean():.4f}', flush=True)
eng.close()
dist.barrier()
dist.destroy_process_group()
