"""Average device time of the 5x5-conv kernel inside a cfg2 fold, for the engine settings given in the environment
(DMP2_CONV_CHUNK, DMP2_CONV_CLUSTER, DMP2_CONV_SMS, DMP2_FUSE_STATS) and every conv mode.  GPU box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

sd, _ = bench.load_weights()
msa = torch.from_numpy(bench.make_msa(3)).cuda()
tag = ' '.join('%s=%s' % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith('DMP2_'))
for mode in (sys.argv[1:] or ['f16x3', 'f16f8', 'f16']):
    eng = Engine(sd, 0, conv_mode=mode)
    eng.fold(msa, None, 2, 10)
    torch.cuda.synchronize()
    eng.set_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.fold(msa, None, 10, 100)
    e1.record()
    torch.cuda.synchronize()
    n, ms = eng.conv_profile()
    print('[%s] %-6s conv %.4f ms/launch (%d launches, %.1f ms)  fold %.1f ms' % (tag, mode, ms / n, n, ms, e0.elapsed_time(e1)), flush=True)
    eng.close()
