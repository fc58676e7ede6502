#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libcdr_b200.so (cuobjdump -sass): the instructions that
show which hardware paths a kernel uses -- DMMA (fp64 tensor pipe), UBLKCP (cp.async.bulk),
LDGSTS (cp.async), SYNCS (mbarrier), DFMA / DADD / DMUL (fp64 pipe), SHFL, MEMBAR, ATOM/RED.

    python profiles/sass_summary.py > profiles/r02/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'matrix-factorization-case-studies_b200', 'convex_dim_red', 'libcdr_b200.so')
KEYS = ['DMMA', 'UBLKCP', 'LDGSTS', 'SYNCS', 'DFMA', 'DADD', 'DMUL', 'SHFL', 'MEMBAR', 'ATOM', 'RED',
        'LDS', 'STS', 'LDG', 'STG', 'BAR']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)),
                           capture_output=True, text=True).stdout.splitlines()
    blocks = re.split(r'\n\s*Function : \S+\n', '\n' + sass)[1:]
    print('# cuobjdump -sass %s (sm_100a)' % os.path.relpath(LIB, ROOT))
    print('%-78s %s' % ('kernel', ' '.join('%7s' % k for k in KEYS)))
    rows = []
    for name, body in zip(names, blocks):
        counts = collections.Counter()
        for line in body.splitlines():
            m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
            if m:
                op = m.group(1).split('.')[0]
                for k in KEYS:
                    if op == k or (k in ('ATOM', 'RED') and op.startswith(k)):
                        counts[k] += 1
        short = re.sub(r'\(.*', '', name).replace('void ', '').replace('cdr::', '')
        rows.append((short, counts))
    for short, counts in sorted(rows):
        print('%-78s %s' % (short[:78], ' '.join('%7d' % counts[k] for k in KEYS)))
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    print('%-78s %s' % ('TOTAL (%d kernels)' % len(rows), ' '.join('%7d' % tot[k] for k in KEYS)))


if __name__ == '__main__':
    main()
