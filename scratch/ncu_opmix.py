#!/usr/bin/env python
"""
usage: scratch/ncu_opmix.py <file.ncu-rep> [pixels]
Executed-instruction mix of the (first) kernel in an `ncu --set full --import-source on` report: warp-level executed
counts per opcode from the SASS source page, top stall-sample sites, and lane-instructions per pixel when `pixels`
is given.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    pixels = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
    hdr = rows[0]
    i_src, i_exec, i_samp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    mix, samples, total, tot_s = collections.Counter(), collections.Counter(), 0, 0
    sites = []
    for r in rows[1:]:
        if len(r) <= i_exec or not r[i_exec].isdigit():
            continue
        text = re.sub(r'^@!?U?P\d+\s+', '', r[i_src].strip())
        op = text.split()[0].split('.')[0]
        n, s = int(r[i_exec]), int(r[i_samp] or 0)
        mix[op] += n
        samples[op] += s
        total += n
        tot_s += s
        sites.append((s, r[i_src].strip()[:70]))
    print(f'{total} warp-instructions' + (f' = {32 * total / pixels:.1f} lane-instructions per pixel' if pixels else ''))
    for op, n in mix.most_common(28):
        per = f'{32 * n / pixels:7.2f}/px' if pixels else ''
        print(f'  {op:10s} {100 * n / total:5.1f} % {per}   stall samples {100 * samples[op] / max(tot_s, 1):5.1f} %')
    print('top stall sites:')
    for s, t in sorted(sites, reverse=True)[:12]:
        print(f'  {100 * s / max(tot_s, 1):5.1f} %  {t}')


if __name__ == '__main__':
    main()
