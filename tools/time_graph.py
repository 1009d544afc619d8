"""A/B of dmp2_set_graph (CUDA-graph replay of the recycling iterations) on one B200: device time per fold with the eager
loop and with the graph, at the headline config (L=300, N=1000, 10 + 100) and at the cfg3 shape (L=150, N=512), plus the
host time spent enqueueing one fold (what a throughput scheduler's threads pay).

    python tools/time_graph.py [folds]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmpfold2_b200.engine import Engine  # noqa: E402
from dmpfold2_b200.predict import read_aln, encode_aln  # noqa: E402
from dmpfold2_b200.synth import synth_msa_structured  # noqa: E402

folds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sd, _ = bench.load_weights()
base = encode_aln(read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
dev = torch.device('cuda', 0)
eng = Engine(sd, 0)
for (l, n) in ((300, 1000), (150, 512)):
    msas = [torch.from_numpy(synth_msa_structured(base, l, n, 50 + i)).to(dev) for i in range(folds + 1)]
    res = {}
    for graph in (False, True, False, True):
        eng.set_graph(graph)
        eng.fold(msas[0], None, 10, 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = 0.0
        e0.record()
        outs = []
        for m in msas[1:]:
            t0 = time.perf_counter()
            outs.append(eng.fold(m, None, 10, 100))
            host += time.perf_counter() - t0
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / folds
        res.setdefault(graph, []).append((ms, host / folds * 1e3, [c.cpu().numpy() for c, _ in outs]))
        print(f'L={l} N={n} graph={int(graph)}: {ms:8.3f} ms/fold on the device, {host / folds * 1e3:6.2f} ms of host enqueue per fold', flush=True)
    same = all(np.array_equal(a, b) for a, b in zip(res[False][-1][2], res[True][-1][2]))
    print(f'L={l} N={n}: graph replay bit-identical to the eager loop: {same}', flush=True)
eng.close()
