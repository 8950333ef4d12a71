"""Per-kernel SASS mnemonic counts and resource usage of libcrbm_b200.so (profiles/rNN_sass.txt).

    python scratch/sass_summary.py > profiles/r02_sass.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'crbm_b200', 'libcrbm_b200.so')
PAT = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'HMMA', 'SYNCS', 'MUFU', 'ACQBULK', 'STG', 'LDG', 'RED', 'ATOM']


def demangle(names):
    r = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True)
    return r.stdout.split('\n')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts, cur, total = collections.OrderedDict(), None, {}
    for line in sass.split('\n'):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            total[cur] += 1
            op = m.group(1)
            for p in PAT:
                if op.startswith(p):
                    counts[cur][p] += 1
    res = subprocess.run(['cuobjdump', '--dump-resource-usage', LIB], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for line in res.split('\n'):
        m = re.match(r'\s*Function (\S+):', line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
            fn = None
    names = list(counts)
    pretty = demangle(names)
    print('# cuobjdump -sass / --dump-resource-usage of crbm_b200/libcrbm_b200.so (sm_100a): instructions, registers, static shared memory,')
    print('# and the counts of the mnemonics that identify the tensor-memory / TMA / mma.sync paths')
    for n, pn in sorted(zip(names, pretty), key=lambda t: t[1]):
        c = counts[n]
        reg, sh = usage.get(n, (None, None))
        tags = ' '.join('%s=%d' % (p, c[p]) for p in PAT if c[p])
        print('%-150s insts=%-6d regs=%-4s smem=%-6s %s' % (pn[:150], total[n], reg, sh, tags))


if __name__ == '__main__':
    main()
