"""The "existing Blackwell kernels" bar: the UNMODIFIED reference (oracle/_ref, with the
torch.symeig shim) run through its own `aln_to_coords(device='cuda')` on the same B200 -- PyTorch's library kernels
(cuDNN conv, cuBLAS, cuSOLVER eigh, cuDNN GRU) -- next to the engine, with the distance of both from the fp32 CPU
reference result and from the fp64 evaluation (tests/golden/cfg2_s0_n10_m100.npz).  Diagnostic only (GPU box); the
output is copied to profiles/."""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dmpfold_oracle as O  # noqa: E402
from oracle import ref_runner  # noqa: E402
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd, _ = bench.load_weights()
base = O.encode_aln(O.read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
msa = O.synth_msa_structured(base, 300, 1000, 0)
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'cfg2_s0_n10_m100.npz'))


def dist(c):
    return 'CA-RMSD vs fp32 CPU reference %.2e A, vs fp64 %.2e A' % (O.kabsch_rmsd(c[:, 1], g['ref32_coords'][:, 1]),
                                                                   O.kabsch_rmsd(c[:, 1], g['ref64_coords'][:, 1]))


eng = Engine(sd, 0)
eng.fold_host(msa, None, 10, 100)
t = time.perf_counter()
c_eng, f_eng = eng.fold_host(msa, None, 10, 100)
print('engine (default conv mode)                         %8.1f ms/target   %s' % ((time.perf_counter() - t) * 1e3, dist(c_eng)), flush=True)
eng.close()
ref = ref_runner.load()
wf = ref_runner.merged_weights_file()
with tempfile.TemporaryDirectory() as tmp:
    aln = os.path.join(tmp, 't.aln')
    ref_runner.write_aln(aln, msa)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        for rep in range(2):
            torch.cuda.synchronize()
            t = time.perf_counter()
            c, f = ref.aln_to_coords(aln, device='cuda', iterations=10, minsteps=100, weights_file=wf)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
        print('reference aln_to_coords(device="cuda"), cudnn.allow_tf32=%-5s %8.1f ms/target   %s' %
              (tf32, dt * 1e3, dist(c.detach().cpu().numpy())), flush=True)
print('(fp32 CPU reference vs fp64: %.2e A; the reference at another thread count vs fp64: %.2e A)' % (
    O.kabsch_rmsd(g['ref32_coords'][:, 1], g['ref64_coords'][:, 1]), O.kabsch_rmsd(g['ref32_alt_coords'][:, 1], g['ref64_coords'][:, 1])))
