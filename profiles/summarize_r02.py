#!/usr/bin/env python
"""Turns the outputs of profiles/run_profile_r02.sh (gpurun_out/r02p_*) into the text
summaries committed under profiles/r02/ and refreshes profiles/ncu_traffic.json.

    python profiles/summarize_r02.py
"""
import collections
import csv
import io
import json
import os
import re
import subprocess

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__shared_mem_per_block_dynamic']
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def launches(path, out, command):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
        agg.setdefault(name, []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(out, 'w') as fh:
        fh.write('# ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none\n'
                 '# (serialised launches: compare SHARES, not absolutes).  Command: %s\n' % command)
        fh.write('%-64s %5s %10s %10s %10s %7s\n' % ('kernel', 'n', 'mean_us', 'min_us', 'max_us', 'share'))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write('%-64s %5d %10.2f %10.2f %10.2f %7.3f\n' %
                     (k[:64], len(v), sum(v) / len(v), min(v), max(v), sum(v) / total))


def kernels(reps, out, traffic):
    with open(out, 'w') as fh:
        fh.write('# ncu --set full --clock-control none --import-source on (cold caches, replayed)\n')
        for rep in reps:
            if not os.path.exists(rep):
                continue
            raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv']).decode()
            rows = list(csv.reader(io.StringIO(raw)))
            hdr, units = rows[0], rows[1]
            fh.write('\n## %s\n' % os.path.basename(rep))
            for row in rows[2:]:
                full = re.sub(r'\(.*', '', row[hdr.index('Kernel Name')])
                fh.write('\n%s\n' % full)
                for m in METRICS:
                    if m in hdr:
                        i = hdr.index(m)
                        fh.write('  %-66s %s %s\n' % (m, row[i], units[i]))
                name = full.replace('void ', '').replace('cdr::', '').split('<')[0]
                tot = 0.0
                for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                    i = hdr.index(m)
                    tot += float(row[i].replace(',', '')) * SCALE[units[i]]
                traffic.setdefault(name, []).append(tot)


def main():
    os.makedirs('profiles/r02', exist_ok=True)
    launches('gpurun_out/r02p_launches.csv', 'profiles/r02/r02p_launches.txt',
             'CDR_NO_CUDA_GRAPH=1 python bench.py --no-stress --no-kmeans --steps 3 --warmup 3 '
             '--cpu-steps 0 --min-timed-ms 1')
    traffic = {}
    kernels(['gpurun_out/r02p_aa.ncu-rep', 'gpurun_out/r02p_gpnh.ncu-rep',
             'gpurun_out/r02p_syrk.ncu-rep', 'gpurun_out/r02p_gemm64.ncu-rep'],
            'profiles/r02/r02p_kernels.txt', traffic)
    with open('profiles/ncu_traffic.json', 'w') as fh:
        json.dump({'source': 'profiles/r02/r02p_kernels.txt (ncu --set full --clock-control none, '
                             'dram__bytes_read.sum + dram__bytes_write.sum per launch)',
                   'dram_bytes_per_launch': {k: sum(v) / len(v) for k, v in traffic.items()}}, fh,
                  indent=1)


if __name__ == '__main__':
    main()
