#!/usr/bin/env python
"""
usage: scratch/sass_hist.py <object file> <kernel name substring> [--loop]
Opcode histogram of one kernel's SASS (cuobjdump -sass).  --loop: only the instructions inside the largest backward
branch (the steady-state row loop of the streaming kernels), which is what the per-pixel instruction budget is about.
"""
import collections
import re
import subprocess
import sys


def main():
    obj, sub = sys.argv[1], sys.argv[2]
    loop = '--loop' in sys.argv
    names = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    funcs = re.findall(r'Function : (\S+)', names)
    match = [f for f in funcs if all(s in f for s in sub.split(','))]
    if not match:
        sys.exit(f'no kernel matching {sub!r}; have {len(funcs)} functions')
    fn = match[0]
    out = subprocess.run(['cuobjdump', '-sass', '-fun', fn, obj], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    lo, hi = 0, len(ins)
    if loop:
        best = (0, 0, 0)
        for i, (addr, text) in enumerate(ins):
            m = re.search(r'\bBRA\b.*?0x([0-9a-f]+)', text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and addr - tgt > best[0]:
                    best = (addr - tgt, tgt, addr)
        lo = next(i for i, (a, _) in enumerate(ins) if a >= best[1])
        hi = next(i for i, (a, _) in enumerate(ins) if a >= best[2]) + 1
    hist = collections.Counter()
    for _, text in ins[lo:hi]:
        text = re.sub(r'^@!?U?P\d+\s+', '', text)
        op = text.split()[0].split('.')[0]
        hist[op] += 1
    total = sum(hist.values())
    print(f'{fn[:100]}\n{"loop" if loop else "kernel"}: {total} instructions')
    for op, n in hist.most_common(40):
        print(f'  {op:10s} {n:6d}  {100 * n / total:5.1f} %')


if __name__ == '__main__':
    main()
