"""Per-kernel SASS mnemonic counts of the shipped library (runs anywhere cuobjdump is: no GPU needed).
Usage: python tools/sass_summary.py [path/to/libmvx_b200.so] > profiles/rN_sass_mnemonics.txt
The columns are the instructions that prove which hardware path a kernel uses on sm_100a: UTCHMMA / UTCQMMA = tcgen05.mma,
UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA) load / store, UBLKCP = cp.async.bulk, STTM / LDTM = tcgen05.st / tcgen05.ld (tensor
memory), UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS + LDGDEPBAR = cp.async, REDG / ATOMG = global reductions / atomics,
HMMA = legacy mma.sync (none expected)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'mvxnet_makise_b200', 'libmvx_b200.so')
KEYS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'STTM', 'LDTM', 'UTCBAR', 'SYNCS', 'LDGSTS', 'LDGDEPBAR', 'REDG', 'ATOMG', 'HMMA', 'FFMA', 'DFMA']
sass = subprocess.run(['cuobjdump', '-sass', lib], check=True, capture_output=True, text=True).stdout
res = subprocess.run(['cuobjdump', '-res-usage', lib], check=True, capture_output=True, text=True).stdout
demangle = lambda names: dict(zip(names, subprocess.run(['c++filt'] + names, check=True, capture_output=True, text=True).stdout.splitlines()))
regs = {}
for m in re.finditer(r'Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)', res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.match(r'\s+Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        counts[cur]['_total'] += 1
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
names = demangle(order)
print(f'# {os.path.relpath(lib, ROOT)}: {len(order)} kernels, sm_100a SASS (cuobjdump -sass); columns: instruction counts in the kernel body')
print('# ' + ' '.join(k.rjust(9) for k in ['instrs', 'regs', 'stack'] + KEYS) + '  kernel')
tot = collections.Counter()
for fn in order:
    c = counts[fn]
    r = regs.get(fn, (0, 0, 0))
    short = re.sub(r'\(anonymous namespace\)::|mvx::', '', names[fn])
    short = re.sub(r'\(.*\)$', '', short)[:110]
    print('  ' + ' '.join(str(v).rjust(9) for v in [c['_total'], r[0], r[1]] + [c[k] for k in KEYS]) + '  ' + short)
    tot.update({k: c[k] for k in KEYS})
print('# totals: ' + ', '.join(f'{k} {tot[k]}' for k in KEYS))
