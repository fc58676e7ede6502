#!/usr/bin/env python
"""Turns gpurun_out/launches.csv (ncu launch list) and gpurun_out/prof.ncu-rep (one
`ncu --set full` capture) into the small text summaries committed under profiles/.

    python profiles/summarize.py <tag>     # writes profiles/<tag>_launches.txt, <tag>_kernels.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum']


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
        agg.setdefault(name, []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(out, 'w') as fh:
        fh.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold cache, serialised:\n'
                 '# compare SHARES, not absolutes).  Command: CDR_NO_CUDA_GRAPH=1 python bench.py '
                 '--steps 2 --warmup 1 --cpu-steps 0\n')
        fh.write('%-64s %5s %10s %10s %10s %7s\n' % ('kernel', 'n', 'mean_us', 'min_us', 'max_us', 'share'))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write('%-64s %5d %10.2f %10.2f %10.2f %7.3f\n' %
                     (k[:64], len(v), sum(v) / len(v), min(v), max(v), sum(v) / total))


def kernels(rep, out):
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv']).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    # DRAM traffic per launch (read + write) of each kernel -> profiles/ncu_traffic.json,
    # which bench.py reports as roofline.traffic
    import json
    traffic = {}
    for row in rows[2:]:
        name = re.sub(r'\(.*', '', row[hdr.index('Kernel Name')]).replace('void ', '').split('<')[0]
        tot = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(m)
            v = float(row[i].replace(',', ''))
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
            tot += v * scale
        traffic.setdefault(name, []).append(tot)
    with open('profiles/ncu_traffic.json', 'w') as fh:
        json.dump({'source': out + ' (ncu --set full --clock-control none, dram__bytes_read.sum + '
                             'dram__bytes_write.sum per launch)',
                   'dram_bytes_per_launch': {k: sum(v) / len(v) for k, v in traffic.items()}}, fh,
                  indent=1)
    with open(out, 'w') as fh:
        fh.write('# ncu --set full --clock-control none --import-source on (one capture per kernel)\n')
        for row in rows[2:]:
            fh.write('\n%s\n' % re.sub(r'\(.*', '', row[hdr.index('Kernel Name')]))
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    fh.write('  %-66s %s %s\n' % (m, row[i], units[i]))


if __name__ == '__main__':
    tag = sys.argv[1]
    launches('gpurun_out/launches.csv', 'profiles/%s_launches.txt' % tag)
    kernels('gpurun_out/prof.ncu-rep', 'profiles/%s_kernels.txt' % tag)
