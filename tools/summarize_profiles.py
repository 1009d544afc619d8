"""Turn gpurun_out/<tag>_launches.csv and <tag>_conv.ncu-rep into the committed summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else 'r3'
out_tag = sys.argv[2] if len(sys.argv) > 2 else 'round1'
what = sys.argv[3] if len(sys.argv) > 3 else 'one fold L=300 N=1000 n=1 m=100 (tools/profile_fold.py 1)'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)

# ---- launch list -> per-kernel table
lcsv = os.path.join(G, f'{tag}_launches.csv')
lines = [l for l in open(lcsv) if not l.startswith('==')] if os.path.isfile(lcsv) else []
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines) if lines else []:
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}[row['Metric Unit']]
    k = row['Kernel Name'].split('(')[0][:100]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f'{out_tag}_launches_summary.txt') if agg else os.devnull, 'w') as fh:
    fh.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none, {what}\n')
    fh.write('# per-launch times are cold-cache and serialised: compare SHARES\n')
    fh.write(f'# total kernel time {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f'{v[1]:10.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg={1e3 * v[1] / v[0]:9.1f} us  {k}\n')

# ---- conv kernel full-set capture -> selected raw metrics
rep = os.path.join(G, f'{tag}_conv.ncu-rep')
if os.path.isfile(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active',
            'sm__pipe_tensor_subpipe_hmma', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
            'sm__throughput.avg.pct', 'gpu__dram_throughput.avg.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
            'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__warps_active.avg.pct', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
            'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_uniform', 'gpc__cycles_elapsed.max', 'sm__cycles_active.avg']
    with open(os.path.join(P, f'{out_tag}_conv_ncu_summary.txt'), 'w') as fh:
        fh.write(f'# ncu --set full --clock-control none -k regex:k_conv5_tc (launch ids in column order), source {tag}_conv.ncu-rep\n')
        idx_name = hdr.index('Kernel Name')
        fh.write('# kernels: ' + ' | '.join(r[idx_name][:60] for r in data) + '\n')
        rd = wr = None
        for i, h in enumerate(hdr):
            if any(w in h for w in want):
                fh.write(f'{h} [{units[i]}] = ' + ', '.join(r[i] for r in data) + '\n')
            if h == 'dram__bytes_read.sum':
                rd = [float(r[i].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[i]] for r in data]
            if h == 'dram__bytes_write.sum':
                wr = [float(r[i].replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[i]] for r in data]
    if rd and wr:
        per = sum(a + b for a, b in zip(rd, wr)) / len(rd)
        json.dump({'dram_bytes_per_launch': per, 'source': f'{out_tag}_conv_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(rd)} launches)'},
                  open(os.path.join(P, 'conv_traffic.json'), 'w'))
print('ok')
