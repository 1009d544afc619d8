"""One fold of the bench workload (for ncu).  usage: profile_fold.py [iterations] [conv_mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = sys.argv[2] if len(sys.argv) > 2 else 'f16x3'
sd, _ = bench.load_weights()
eng = Engine(sd, 0, conv_mode=mode)
msa = bench.make_msa(0)
coords, conf = eng.fold_host(msa, None, iters, bench.N_MIN)
print('mean conf', float(conf.mean()), 'launches', eng.launch_count, eng.stage_times())
