"""Attribute ncu warp-stall samples to SOURCE LINES of one kernel (dev tool, runs where ncu / cuobjdump / nvdisasm are).
ncu's CSV export of the source page carries SASS addresses only; this joins them with nvdisasm's line table of the shipped library.
Usage: python tools/ncu_lines.py report.ncu-rep <launch index in the report> <kernel substring, e.g. tc3_layer_kernelILb1ELb1ELi0E> [top N]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, idx, ksub = sys.argv[1], int(sys.argv[2]), sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(idx), '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print('#', rows[0][1][:160])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
seen = {}
for r in rows[2:]:
    if len(r) == len(hdr) and r[ix['# Samples']].isdigit():
        seen[int(r[ix['Address']], 16)] = r
base = min(seen)
with tempfile.TemporaryDirectory() as td:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'mvxnet_makise_b200', 'libmvx_b200.so')], cwd=td, capture_output=True)
    lines = []
    for cub in glob.glob(os.path.join(td, '*.cubin')):
        if os.path.basename(cub).count('-') == 0:
            dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout
            if ksub in dis:
                lines = dis.splitlines()
                break
start = next(i for i, l in enumerate(lines) if l.strip().startswith('.section') and ksub in l)
off2line, off2op, cur = {}, {}, None
for l in lines[start + 1:]:
    if l.strip().startswith('.section') and '.text.' in l:
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m:
        off2line[int(m.group(1), 16)] = cur
        off2op[int(m.group(1), 16)] = m.group(2).split('.')[0]
# the line table comes from the library on disk: make sure it is the binary that was profiled (same opcode at every address)
same = tot_ops = 0
for a, r in seen.items():
    src = r[ix['Source']].strip().split()
    op = (src[1] if src and src[0].startswith('@') and len(src) > 1 else (src[0] if src else '')).split('.')[0]
    tot_ops += 1
    same += off2op.get(a - base) == op
print(f'# binary check: {same} of {tot_ops} instructions of the report match the library on disk' + ('' if same >= 0.98 * tot_ops else '  ** MISMATCH: rebuild the profiled source first, the lines below are not trustworthy **'))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
by, ins, why = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for a, r in seen.items():
    ln = off2line.get(a - base)
    by[ln] += int(r[ix['# Samples']])
    ins[ln] += int(r[ix['Instructions Executed']])
    for h in stalls:
        why[ln][h[6:]] += int(r[ix[h]])
print('# total samples', sum(by.values()), 'executed warp instructions', sum(ins.values()))
srcs = {}
for ln, c in by.most_common(topn):
    text = '?'
    if ln:
        fn = os.path.join(ROOT, 'mvxnet_makise_b200', 'csrc', ln[0])
        if os.path.exists(fn):
            srcs.setdefault(fn, open(fn).read().splitlines())
            text = srcs[fn][ln[1] - 1].strip()[:100]
    top2 = ', '.join(f'{k} {v}' for k, v in why[ln].most_common(2))
    print(f'{(ln[0] + ":" + str(ln[1])) if ln else "-":>22} {c:6d} {ins[ln]:10d}  [{top2}]  {text}')
