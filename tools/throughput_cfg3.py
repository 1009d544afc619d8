"""BASELINE.json configs[2]: a batch of 256 synthetic targets (L=150, N=512, 10 recycles + 100 minimiser steps), sharded
round-robin over the ranks (one process per GPU, no data-path collective) and, on every GPU, over K CUDA streams with
one engine each (dmpfold2_b200.parallel.StreamPool).  GPU box.

    python tools/throughput_cfg3.py [--targets 256] [--streams 1,2,3,4] [--conv-sms 0]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/throughput_cfg3.py --streams 3

Prints, per stream count, the wall time of the whole batch (device events, max over ranks) and targets/s.  Inputs are
device-resident before the clock starts (the alignments are 77 KB each); results stay on the device.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200 import parallel as P  # noqa: E402
from dmpfold2_b200.predict import read_aln, encode_aln  # noqa: E402
from dmpfold2_b200.synth import synth_msa_structured  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--targets', type=int, default=256)
ap.add_argument('--L', type=int, default=150)
ap.add_argument('--N', type=int, default=512)
ap.add_argument('--streams', type=str, default='1,2,3,4')
ap.add_argument('--conv-sms', type=int, default=0)
ap.add_argument('--conv-mode', type=str, default='f16f8')
ap.add_argument('--scan-rows', type=int, default=-1, help='alignment columns per shared vgru scan (0 = every fold scans its own; -1 = StreamPool default: 384 when streams > 1)')
ap.add_argument('--host-threads', type=int, default=1, help='1: one enqueueing host thread per stream; 0: a single thread')
ap.add_argument('--dynamic', type=int, default=-1, help='conv unit schedule: 1 dynamic, 0 static, -1 StreamPool default (dynamic when streams > 1)')
args = ap.parse_args()

rank = int(os.environ.get('RANK', '0'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(local_rank)
dev = torch.device('cuda', local_rank)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
sd, _ = bench.load_weights()
base = encode_aln(read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
mine = P.targets_for_rank(args.targets, rank, world)
msas = [torch.from_numpy(synth_msa_structured(base, args.L, args.N, t)).to(dev) for t in mine]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for k in [int(x) for x in args.streams.split(',')]:
    pool = P.StreamPool(sd, local_rank, streams=k, conv_mode=args.conv_mode, conv_sms=args.conv_sms,
                        conv_dynamic=None if args.dynamic < 0 else bool(args.dynamic), scan_rows=None if args.scan_rows < 0 else args.scan_rows)
    pool.fold_all(msas[:k], None, 10, 100, host_threads=bool(args.host_threads))                       # warm-up: workspaces, tensor maps, module load
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = pool.fold_all(msas, None, 10, 100, host_threads=bool(args.host_threads))
    e1.record()
    barrier()
    ms = P.max_over_ranks(e0.elapsed_time(e1), dev)
    ok = all(bool(torch.isfinite(c).all()) for c, _ in res)
    if rank == 0:
        print(json.dumps({'workload': '%d targets L=%d N=%d, 10 iter + 100 min-steps' % (args.targets, args.L, args.N), 'gpus': world,
                          'streams_per_gpu': k, 'conv_sms': args.conv_sms, 'conv_dynamic': pool.conv_dynamic, 'host_threads': bool(args.host_threads), 'scan_rows': pool.scan_rows, 'conv_mode': args.conv_mode, 'batch_ms': ms,
                          'ms_per_target': ms / args.targets, 'targets_per_s': args.targets / ms * 1e3, 'finite': ok}), flush=True)
    pool.close()
if world > 1:
    dist.destroy_process_group()
