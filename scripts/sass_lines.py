#!/usr/bin/env python
"""Attribute executed warp-instructions of one kernel to CUDA source lines.

    ncu -i X.ncu-rep --page source --csv > src.csv
    python scripts/sass_lines.py src.csv tap-net_b200/lib/libtapenv.so 'step_kernelILi0ELb1'

Joins the per-SASS-instruction 'Instructions Executed' column of the ncu source page with the line table
that nvdisasm -g prints for the same cubin (instruction order is identical)."""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, lib, pat = sys.argv[1], sys.argv[2], sys.argv[3]

rows = list(csv.reader(open(src_csv)))
body = [r for r in rows if len(r) > 6 and r[0].startswith('0x')]
# several launches of the same kernel are concatenated: take the first block (addresses restart)
first = [body[0]]
for r in body[1:]:
    if r[0] == body[0][0]:
        break
    first.append(r)
d = tempfile.mkdtemp()
subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(sass) if l.startswith('.text.') and pat in l][0]
lines = []   # (file:line stack, instr text)
cur = '?'
for l in sass[start + 1:]:
    if l.startswith('//---') or l.startswith('\t.section'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        cur = '%s:%s' % (os.path.basename(m.group(1)), m.group(2))
        if inl:
            cur += ' <- ' + ' <- '.join('%s:%s' % (os.path.basename(a), b) for a, b in inl)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(2)))
n = min(len(lines), len(first))
if len(lines) != len(first):
    print('warning: %d SASS lines vs %d profiled instructions' % (len(lines), len(first)), file=sys.stderr)
tot = collections.Counter(); stall = collections.Counter()
for (loc, txt), r in zip(lines[:n], first[:n]):
    key = loc.split(' <- ')[0] if '--leaf' in sys.argv else (loc.split(' <- ')[-1] if '--root' in sys.argv else loc)
    tot[key] += int(r[5]); stall[key] += int(r[2])
warps = max(int(r[5]) for r in first)
print('warps', warps, 'instr/warp %.1f' % (sum(tot.values()) / warps))
for k, v in tot.most_common(int(os.environ.get("TOPN","45"))):
    print('%7.1f  stall=%5d  %s' % (v / warps, stall[k], k))
