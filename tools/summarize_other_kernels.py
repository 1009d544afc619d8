"""Selected raw metrics of the --set full captures of the NON-conv kernels of a fold (tools/gpu_sessions/r2/run31_final.sh):
duration, DRAM bytes and achieved DRAM GB/s, L2 hit rate, SM / memory throughput %, grid / block / cluster, registers.

    python tools/summarize_other_kernels.py gpurun_out/r2_final > profiles/round2_other_kernels_ncu.txt
"""
import collections
import csv
import os
import subprocess
import sys

G = sys.argv[1]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'usecond': 1e-6,
        'nsecond': 1e-9, 'msecond': 1e-3, 'second': 1.0}


def load(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return float('nan')


print('# ncu --set full --clock-control none, one fold of the bench workload (L=300, N=1000, f16f8), non-conv kernels.')
print('# per kernel: launches captured, mean duration, DRAM read+write per launch and the achieved DRAM rate, L2 hit rate,')
print('# compute (SM) and memory throughput as % of peak, launch shape.  Cold-cache, serialised, clocks not locked.')
for rep in ('other_kernels.ncu-rep', 'vgru_step.ncu-rep'):
    path = os.path.join(G, rep)
    if not os.path.isfile(path):
        print(f'# {rep}: missing')
        continue
    hdr, units, data = load(path)
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name):
        i = col.get(name)
        if i is None:
            return float('nan')
        return num(r[i]) * UNIT.get(units[i], 1.0)
    agg = collections.OrderedDict()
    for r in data:
        k = r[col['Kernel Name']].split('(')[0].replace('<unnamed>::', '')
        agg.setdefault(k, []).append(r)
    for k, rs in agg.items():
        n = len(rs)
        dur = sum(get(r, 'gpu__time_duration.sum') for r in rs) / n
        rd = sum(get(r, 'dram__bytes_read.sum') for r in rs) / n
        wr = sum(get(r, 'dram__bytes_write.sum') for r in rs) / n
        hit = sum(get(r, 'lts__t_sector_hit_rate.pct') for r in rs) / n
        smt = sum(get(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed') for r in rs) / n
        mem = sum(get(r, 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed') for r in rs) / n
        drt = sum(get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') for r in rs) / n
        r0 = rs[0]
        grid = r0[col['launch__grid_size']] if 'launch__grid_size' in col else '?'
        blk = r0[col['launch__block_size']] if 'launch__block_size' in col else '?'
        regs = r0[col['launch__registers_per_thread']] if 'launch__registers_per_thread' in col else '?'
        clus = r0[col['launch__cluster_dim_x']] if 'launch__cluster_dim_x' in col else '-'
        print(f'{k:28s} n={n:2d}  {dur * 1e6:9.1f} us  dram {rd / 1e6:8.2f} MB rd + {wr / 1e6:8.2f} MB wr = {(rd + wr) / dur / 1e9:7.1f} GB/s '
              f'({drt:5.1f} % of peak)  L2 hit {hit:5.1f} %  SM {smt:5.1f} %  mem {mem:5.1f} %  grid {grid} x {blk} thr, cluster {clus}, {regs} regs')
