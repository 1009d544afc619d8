"""Turn the raw output of tools/gpu_final_r2.sh (gpurun_out/final_r2/) into the committed summaries under profiles/:
per-kernel launch table of the bench command, selected raw metrics of the --set full captures (conv in both modes, the
tensor-core GEMMs), per-warp-role stall summary of the conv kernel, conv_traffic.json."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out', sys.argv[1] if len(sys.argv) > 1 else 'final_r2')
P = os.path.join(ROOT, 'profiles')
TAG = 'round2'

# ---- launch list -> per-kernel table
lcsv = os.path.join(G, 'bench_launches.csv')
if os.path.isfile(lcsv):
    lines = [l for l in open(lcsv) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}[row['Metric Unit']]
        k = row['Kernel Name'].split('(<unnamed>')[0].split('(const')[0][:110]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f'{TAG}_final_bench_launches_summary.txt'), 'w') as fh:
        fh.write('# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras\n')
        fh.write('# (1 warm-up + 2 timed device-resident folds, 1 warm-up + 2 timed host-buffer folds; default conv mode f16f8)\n')
        fh.write('# per-launch times are cold-cache and serialised: compare SHARES\n')
        fh.write(f'# total kernel time {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f'{v[1]:10.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:6d}  avg={1e3 * v[1] / v[0]:9.1f} us  {k}\n')

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster', 'sm__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic',
        'sm__cycles_active.avg', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed']


def raw_summary(rep, out, title):
    if not os.path.isfile(rep):
        return None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    iname = hdr.index('Kernel Name')
    rd = wr = None
    with open(out, 'w') as fh:
        fh.write(title + '\n')
        fh.write('# kernels (launch ids in column order): ' + ' | '.join(r[iname].split('(<unnamed>')[0][-60:] for r in data) + '\n')
        for i, h in enumerate(hdr):
            if any(w in h for w in WANT):
                fh.write(f'{h} [{units[i]}] = ' + ', '.join(r[i] for r in data) + '\n')
            if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                vals = [float(r[i].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[i]] for r in data]
                if h.endswith('read.sum'):
                    rd = vals
                else:
                    wr = vals
    return [a + b for a, b in zip(rd, wr)] if rd and wr else None


def role_stalls(rep, out_fh, label):
    """warp-stall samples of the conv kernel aggregated by warp role (address ranges between the setmaxnreg markers)"""
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
    ia, isrc, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples')
    seen, d1 = set(), []
    for r in data:
        if r[ia] in seen:
            break
        seen.add(r[ia])
        d1.append(r)

    def n(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    idx = {}
    for i, r in enumerate(d1):
        s = r[isrc]
        if 'USETMAXREG.DEALLOC' in s and 'dealloc' not in idx:
            idx['dealloc'] = i
        if 'USETMAXREG.TRY_ALLOC' in s and 'alloc' not in idx:
            idx['alloc'] = i
        if 'UTMALDG' in s:
            idx['last_tma'] = i
    if len(idx) < 3:
        return
    out_fh.write(f'## {label}: warp-stall samples by warp role\n')
    for name, a, b in (('producer warp', idx['dealloc'], idx['last_tma'] + 30), ('MMA issuer warp', idx['last_tma'] + 30, idx['alloc']),
                       ('8 epilogue warps (+ idle warps at the final barrier)', idx['alloc'], len(d1))):
        tot = sum(n(r[isamp]) for r in d1[a:b])
        agg = {}
        for r in d1[a:b]:
            for x in stalls:
                agg[x] = agg.get(x, 0) + n(r[hdr.index(x)])
        top = sorted(((v, k[6:]) for k, v in agg.items() if v), reverse=True)[:5]
        out_fh.write(f'   {name}: {tot} samples; ' + ', '.join(f'{k}={v}' for v, k in top) + '\n')


t = raw_summary(os.path.join(G, 'conv_f16f8.ncu-rep'), os.path.join(P, f'{TAG}_conv_ncu_summary.txt'),
                '# ncu --set full --clock-control none --import-source on -k regex:k_conv5_tc<2 -s 40 -c 2, python tools/time_conv.py f16f8 '
                '(the default conv mode inside a cfg2 fold: persistent cta_group::2 kernel, 5-tap accumulation chains)')
if t:
    json.dump({'dram_bytes_per_launch': sum(t) / len(t),
               'source': f'{TAG}_conv_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(t)} launches)'},
              open(os.path.join(P, 'conv_traffic.json'), 'w'))
raw_summary(os.path.join(G, 'conv_f16x3.ncu-rep'), os.path.join(P, f'{TAG}_conv_f16x3_ncu_summary.txt'),
            '# same capture for the f16x3 conv (3 MMAs per MAC, one tap per accumulation chain): python tools/time_conv.py f16x3')
raw_summary(os.path.join(G, 'gemm_tc.ncu-rep'), os.path.join(P, f'{TAG}_gemm_ncu.txt'),
            '# ncu --set full of the first 10 launches of the tcgen05 GEMM service (k_conv5_tc<f16x3> in gemm mode) inside a cfg2 fold, in launch '
            'order: MSA-feature GEMMs on the side stream (Gram system 1000x1000x6400, wy^T = X K^-1 6300x1000x1024, inverse (I - X wy)/ridge '
            '6300x6300x1024 = the covariance-sized contraction), hgru input projections (300x1536x512), stem slabs (32768x384x1024)')
for rep, label in (('conv_f16f8.ncu-rep', 'f16f8'), ('conv_f16x3.ncu-rep', 'f16x3')):
    pth = os.path.join(G, rep)
    if os.path.isfile(pth):
        with open(os.path.join(P, f'{TAG}_conv_ncu_summary.txt' if label == 'f16f8' else f'{TAG}_conv_f16x3_ncu_summary.txt'), 'a') as fh:
            role_stalls(pth, fh, label)
print('ok')
