"""Per-kernel census of tensor-core / TMA / memory instructions in libdsb200.so (cuobjdump -sass).  No GPU needed.
    python tools/sass_census.py [filter-regex] > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'deepsphere_tf1_b200', 'libdsb200.so')
MN = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDGSTS', 'FFMA']
flt = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
sass = subprocess.run(['cuobjdump', '-sass', SO], stdout=subprocess.PIPE, text=True).stdout
names = {}
cur = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1).split('.')[0]
        if op in MN:
            counts[cur][op] += 1
dem = subprocess.run(['c++filt'], input='\n'.join(counts.keys()), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print('SASS census of deepsphere_tf1_b200/libdsb200.so (cuobjdump -sass, sm_100a): instruction counts per kernel.')
print('UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG / UTMASTG = TMA tensor load / store,')
print('UBLKCP = cp.async.bulk (1-D TMA), HMMA = legacy mma.sync, LDGSTS = cp.async, FFMA = fp32 FMA.')
print()
print('{:<86s} '.format('kernel') + ' '.join('{:>7s}'.format(m) for m in MN))
for (mangled, c), name in zip(counts.items(), dem):
    name = re.sub(r'\(.*', '', name).replace('void ', '').replace('dsb200::', '').replace('(anonymous namespace)::', '')
    if flt is not None and not flt.search(name):
        continue
    if not any(c[m] for m in MN[:7]) and flt is None:
        continue                                          # plain kernels: listed only when asked for
    print('{:<86s} '.format(name[:86]) + ' '.join('{:>7d}'.format(c[m]) for m in MN))
