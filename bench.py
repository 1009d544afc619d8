#!/usr/bin/env python
"""bench.py -- the DMPfold2 hot path on B200: ms/target at L=300, N=1000, 10 recycling iterations + 100
minimiser steps (BASELINE.json `metric`, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one whole fold (MSA features -> GRUs -> 11 ResNet passes -> MDS/coordinate GRU -> minimiser ->
backbone) of one synthetic target per GPU.  Multi-GPU: independent targets, one per rank per step, no
data-path collective (weak scaling); launched under torch.distributed.run by the driver.

Printed JSON (rank 0, one line): see the contract in the task statement; `value` is device-resident timing
(CUDA events), `e2e` goes through the host-buffer C-ABI call dmp2_fold_host (H2D + fold + D2H + sync).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_RES, N_SEQ, N_ITER, N_MIN = 300, 1000, 10, 100
WORKLOAD = f'single target L={L_RES} N={N_SEQ} synthetic MSA, {N_ITER} iter + {N_MIN} min-steps'
METRIC = 'ms/target (L=300, N=1k MSA, 10 iter+100 min)'
CONV_FLOPS_PER_LAUNCH = 2.0 * L_RES * L_RES * 3200 * 512          # one 5x5 conv, 294.9 GF


def load_weights():
    wdir = os.path.join(ROOT, 'dmpfold2_b200', 'trained_model')
    if all(os.path.isfile(os.path.join(wdir, f'FINAL_fullmap_e2e_model_part{p}.pt')) for p in (1, 2)):
        from dmpfold2_b200.predict import load_weights as lw
        return lw(None), 'trained DMPfold2 weights'
    from dmpfold2_b200.synth import random_state_dict
    return random_state_dict(0), 'random-init weights of the reference architecture (trained files absent)'


def make_msa(seed):
    """Structured synthetic MSA: PF10963 rows/columns resampled to (N_SEQ, L_RES)."""
    from dmpfold2_b200.predict import read_aln, encode_aln
    from dmpfold2_b200.synth import synth_msa_structured
    base = encode_aln(read_aln(os.path.join(ROOT, 'tests', 'golden', 'PF10963.aln')))
    return synth_msa_structured(base, L_RES, N_SEQ, seed)


class ClockSampler(threading.Thread):
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


def cpu_sample(sd, msa):
    """Bounded CPU sample of the same workload with the oracle PORT (only used when oracle/_ref is not staged): the
    once-per-target part (features, vgru, hgru) + ONE of the 11 ResNet/MDS/coord passes + the two minimiser calls,
    extrapolated as fixed + 11*pass + refine.  Returns (ms_per_target_extrapolated, description)."""
    import torch
    from oracle import dmpfold_oracle as O
    torch.set_num_threads(os.cpu_count())
    orc = O.Oracle(sd)
    msa_t = torch.from_numpy(np.ascontiguousarray(msa)).long()
    with torch.no_grad():
        t0 = time.perf_counter()
        feats = O.msa_features(msa_t).permute(2, 0, 1).unsqueeze(0)
        m1 = orc.mat1d(msa_t)
        outer = (m1.unsqueeze(1) * m1.unsqueeze(2)).unsqueeze(0)
        resinp = torch.cat((outer, feats, torch.zeros((1, 1, L_RES, L_RES)) - 1), dim=1)
        t1 = time.perf_counter()
        ca, conf = orc.one_pass(resinp, m1)
        t2 = time.perf_counter()
        O.refine_coords(ca, N_MIN)
        t3 = time.perf_counter()
    fixed, one_pass, refine = t1 - t0, t2 - t1, t3 - t2
    total = fixed + (N_ITER + 1) * one_pass + 2 * refine
    desc = (f'oracle port on {os.cpu_count()} host threads: once-per-target part {fixed:.1f}s + one of {N_ITER + 1} passes '
            f'{one_pass:.1f}s + one of 2 minimiser calls {refine:.2f}s, extrapolated fixed+{N_ITER + 1}*pass+2*refine')
    return total * 1e3, desc


def reference_fold_ms(msa):
    """ONE full, un-extrapolated fold of the cfg2 target by the UNMODIFIED reference (oracle/_ref, staged by
    oracle/make_ref.py) through its own public API aln_to_coords(device='cpu') on all host cores: .aln text in,
    module construction + weight load + features + 11 passes + minimiser inside, exactly as a reference user pays."""
    import torch
    from oracle import ref_runner
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    coords, confs = ref_runner.fold(msa, iterations=N_ITER, minsteps=N_MIN)
    ms = (time.perf_counter() - t0) * 1e3
    assert coords.shape == (L_RES, 5, 3) and np.isfinite(coords).all()
    return ms


def cpu_baseline(sd, msa):
    """The `cpu_baseline` object of the engine arm (rank 0, N=1): ONE full fold of the same target by the unmodified
    reference when oracle/_ref is staged, else the bounded sample of the oracle port."""
    from oracle import ref_runner
    if ref_runner.available():
        v = reference_fold_ms(msa)
        return {'value': v, 'unit': 'ms/target', 'cores': os.cpu_count(), 'kind': 'reference',
                'sample': 'ONE full fold of the same target by the unmodified reference dmpfold.aln_to_coords(device="cpu") '
                          '(oracle/_ref + torch.symeig shim), cold, no extrapolation'}
    v, desc = cpu_sample(sd, msa)
    return {'value': v, 'unit': 'ms/target', 'cores': os.cpu_count(), 'kind': 'port', 'sample': desc}


REF_BUDGET_S = 240.0          # the whole reference-arm run stays within a few minutes


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import ref_runner
    sd, wdesc = load_weights()
    if ref_runner.available():
        ref_runner.load()
        ref_runner.merged_weights_file()                     # untimed set-up (the reference reads ONE weights file)
        n_warm = min(args.warmup, 1)                         # one warm-up fold at most (page cache, thread pools, oneDNN primitives)
        first = reference_fold_ms(make_msa(1000))
        vals = [] if n_warm else [first]
        n_timed = max(1, min(args.steps, int((REF_BUDGET_S - first / 1e3) // (first / 1e3))))
        vals += [reference_fold_ms(make_msa(1001 + i)) for i in range(n_timed - len(vals))]
        kind = 'reference'
        desc = (f'unmodified reference dmpfold.aln_to_coords(device="cpu") (oracle/_ref + torch.symeig shim), full folds, no '
                f'extrapolation: {n_warm} warm-up + {n_timed} timed folds of {args.steps} requested (bounded to ~{REF_BUDGET_S:.0f}s), '
                f'{os.cpu_count()} host threads, torch {__import__("torch").__version__}')
    else:
        vals = []
        desc = ''
        for i in range(min(args.warmup, 1) + min(args.steps, 3)):
            v, desc = cpu_sample(sd, make_msa(1000 + i))
            if i >= min(args.warmup, 1):
                vals.append(v)
        kind = 'port'
    val = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'ms/target', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': val, 'higher_is_better': False, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': f'synthetic structured MSA; {wdesc}',
        'config': {'workload': WORKLOAD, 'note': 'the reference on host CPU cores; rank 0 only'},
        'cpu_baseline': {'value': val, 'unit': 'ms/target', 'cores': os.cpu_count(), 'kind': kind, 'sample': desc},
        'e2e': {'value': val, 'unit': 'ms/target', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='engine')
    ap.add_argument('--conv-mode', type=str, default='f16f8', choices=['f16f8', 'f16x3', 'f16', 'ffma'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the informational sections (other conv modes, throughput mode)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'              # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist
    from dmpfold2_b200.engine import Engine
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    sd, wdesc = load_weights()
    eng = Engine(sd, local_rank, conv_mode=args.conv_mode)
    dev = torch.device('cuda', local_rank)

    nsteps = args.warmup + args.steps
    msas = [make_msa(rank * 10007 + i) for i in range(nsteps)]
    msas_dev = [torch.from_numpy(m).to(dev) for m in msas]
    msas_pin = []
    for m in msas:
        t = torch.empty(m.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = m
        msas_pin.append(t.numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (`value`) ---------------------------------------------------------
    for i in range(args.warmup):
        eng.fold(msas_dev[i], None, N_ITER, N_MIN)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.set_profile(True)
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    outs = []
    for i in range(args.warmup, nsteps):
        outs.append(eng.fold(msas_dev[i], None, N_ITER, N_MIN))
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    n_conv, conv_ms = eng.conv_profile()
    eng.set_profile(False)
    sampler.stop_flag = True
    sampler.join()
    finite = all(bool(torch.isfinite(c).all()) for c, _ in outs)
    t_max = max_over_ranks(dev_ms)

    # ---- end-to-end timing through the host-buffer C-ABI call ---------------------------------------
    for i in range(min(args.warmup, 1)):
        eng.fold_host(msas_pin[i], None, N_ITER, N_MIN)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, nsteps):
        eng.fold_host(msas_pin[i], None, N_ITER, N_MIN)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_max = max_over_ranks(e2e_ms)
    stages = eng.stage_times()

    # ---- informational: the other conv modes (f16x3 = reference-accuracy conv, 3 MMAs per MAC; f16 = single fp16 MMA,
    # which does NOT hold the parity bar)
    fast_ms = {}
    if args.conv_mode == 'f16f8' and not args.no_extras:
        for fm in ('f16x3', 'f16'):
            eng.set_conv_mode(fm)
            eng.fold(msas_dev[0], None, N_ITER, N_MIN)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            eng.fold(msas_dev[-1], None, N_ITER, N_MIN)
            f1.record()
            torch.cuda.synchronize()
            fast_ms[fm] = f0.elapsed_time(f1)
        eng.set_conv_mode(args.conv_mode)

    # ---- informational: throughput mode (dmpfold2_b200.parallel.StreamPool): K engines on K streams of this GPU fold
    # independent targets concurrently (the one-target-per-stream layout of BASELINE.json configs[2]); the latency-bound
    # stages of one target overlap the convs of the others (profiles/round2_throughput_v2.txt has the full sweep).
    tp_ms = {}
    try:
        from dmpfold2_b200.parallel import StreamPool
        for k_streams, conv_sms in (() if args.no_extras else ((2, 0), (3, 0))):
            pool = StreamPool(sd, local_rank, streams=k_streams, conv_mode=args.conv_mode, conv_sms=conv_sms)
            n_tp = max(k_streams * 2, (args.steps // k_streams) * k_streams)
            idx = [args.warmup + (k % args.steps) for k in range(n_tp)]
            pool.fold_all([msas_dev[i] for i in idx[:k_streams]], None, N_ITER, N_MIN)
            torch.cuda.synchronize()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            pool.fold_all([msas_dev[i] for i in idx], None, N_ITER, N_MIN)
            t1e.record()
            torch.cuda.synchronize()
            tp_ms['%d_streams%s' % (k_streams, '_conv_on_%d_sms' % conv_sms if conv_sms else '')] = t0e.elapsed_time(t1e) / n_tp
            pool.close()
    except Exception as ex:                      # informational only
        tp_ms['failed'] = str(ex)

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.isfile(pk):
            peaks = json.load(open(pk))
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback 1.4 PFLOP/s sustained (of fallback)'
        avg_conv_ms = conv_ms / max(n_conv, 1)
        achieved = CONV_FLOPS_PER_LAUNCH / (avg_conv_ms * 1e-3) / 1e12 if n_conv else 0.0
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
        if os.path.isfile(tp):
            traffic = json.load(open(tp)).get('dram_bytes_per_launch')
        total_targets = args.steps * world
        line = {
            'metric': METRIC, 'value': t_max / total_targets, 'unit': 'ms/target', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': t_max / args.steps, 'higher_is_better': False, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {'f16x3': 'f16 hi/lo operand split (3 MMAs per MAC), per-tap tcgen05 chains summed in f32 registers (f32 elsewhere)',
                      'f16f8': 'f16 main term + fp8 hi/lo correction terms, per-tap tcgen05 chains summed in f32 registers (f32 elsewhere)'}.get(args.conv_mode, args.conv_mode),
            'data': f'synthetic structured MSA (PF10963 resampled, seeded); {wdesc}',
            'config': {'workload': WORKLOAD, 'conv_mode': args.conv_mode, 'targets_per_gpu_per_step': 1,
                       'l2': 'per-step working set ~0.9 GB > 126 MB L2 and every step folds a different target',
                       'outputs_finite': finite},
            'e2e': {'value': e2e_max / total_targets, 'unit': 'ms/target', 'h2d_bytes_per_step': int(N_SEQ * L_RES),
                    'd2h_bytes_per_step': int(L_RES * 16 * 4)},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
            'roofline': {'bound': 'tensor', 'kernel': 'k_conv5_tc (persistent, cta_group::2 pairs)', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved / peak_tf, 'traffic': traffic, 'launches_timed': n_conv,
                         'avg_launch_ms': avg_conv_ms, 'conv_share_of_step': avg_conv_ms * 16 * (N_ITER + 1) * args.steps / dev_ms,
                         'peak_source': peak_src,
                         'mma_fp16_equivalents_per_mac': {'f16f8': 2.0, 'f16x3': 3.0, 'f16': 1.0}.get(args.conv_mode),
                         'mma_equivalent_tflops': achieved * {'f16f8': 2.0, 'f16x3': 3.0, 'f16': 1.0}.get(args.conv_mode, 1.0),
                         'note': 'algorithmic FLOPs 2*L^2*3200*512 per launch; f16x3 issues 3 fp16 MMAs per algorithmic MAC, '
                                 'f16f8 one fp16 MMA + two fp8 MMAs (2 fp16-equivalents)'},
            'stage_ms_last_e2e_step': stages,
            'other_conv_modes_ms_per_target': fast_ms,
            'throughput_mode_ms_per_target': tp_ms,
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(sd, msas[-1])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
