"""One rank of a halo-sharded fold, run as its own process by tests/test_gpu_strip.py (and usable under torchrun):
    python tests/strip_worker.py <job.npz> <out_dir>
RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT come from the environment.  Each rank builds an engine on GPU
(rank mod #GPUs), joins a gloo group for the one-off exchange of the window IPC handles, folds the job's alignment
through parallel.StripGroup and writes its result to <out_dir>/rank<r>.npz."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    job = np.load(sys.argv[1], allow_pickle=False)
    out_dir = sys.argv[2]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dmpfold2_b200.engine import Engine
        from dmpfold2_b200.parallel import StripGroup
        from dmpfold2_b200.predict import load_weights
        dev = rank % torch.cuda.device_count()
        eng = Engine(load_weights(None), dev, conv_mode=str(job['mode']))
        grp = StripGroup(eng)
        tmpl = job['tmpl'] if 'tmpl' in job.files else None
        results = {}
        for i, (n, m) in enumerate(job['runs']):
            c, f = grp.fold_host(job['msa'], tmpl, int(n), int(m))
            results[f'coords{i}'], results[f'confs{i}'] = c, f
        grp.close()
        eng.close()
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), **results)
    finally:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
