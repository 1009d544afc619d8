"""The "existing Blackwell kernels" bar: the reference's algorithm (oracle restatement) executed by PyTorch's own
CUDA library kernels (cuDNN conv, cuBLAS, cuSOLVER eigh, cuDNN GRU) on the same B200, next to the engine.
Diagnostic only (GPU box); results are copied to profiles/."""
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

sd, _ = bench.load_weights()
msa = bench.make_msa(0)
eng = Engine(sd, 0)
eng.fold_host(msa, None, 10, 100)
t = time.perf_counter()
c_eng, f_eng = eng.fold_host(msa, None, 10, 100)
t_eng = time.perf_counter() - t
print('engine (f16f8)            %.1f ms/target' % (t_eng * 1e3), flush=True)
ref = None
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    orc = O.Oracle(sd, device='cuda:0')
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        c, f = orc.fold(msa, iterations=10, minsteps=100)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print('torch CUDA libraries, cudnn.allow_tf32=%s   %.1f ms/target   CA-RMSD vs engine %.2e A' %
          (tf32, dt * 1e3, O.kabsch_rmsd(c[:, 1].cpu().numpy(), c_eng[:, 1])), flush=True)
