#!/usr/bin/env python
"""Per-source-line executed warp-instructions and stall samples of one kernel from an ncu source-page CSV
(`ncu -i X.ncu-rep --page source --csv`), joined with the line table `nvdisasm -g` prints for the same cubin.

    python scripts/src_lines.py SRC.csv LIB.so KERNEL_SUBSTRING UNITS [TOPN]      # UNITS = environments in the launch
"""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, lib, pat, nenv = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(src_csv)))
hdr = [r for r in rows if r and r[0] == 'Address'][0]
body = [r for r in rows if len(r) > 6 and r[0].startswith('0x')]
d = tempfile.mkdtemp()
subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(sass) if l.startswith('.text.') and pat in l][0]
lines, cur = [], '?'
for l in sass[start + 1:]:
    if l.startswith('//---') or l.startswith('\t.section'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = '%s:%s' % (os.path.basename(m.group(1)), m.group(2))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(2)))
if len(lines) != len(body):
    print('warning: %d SASS lines vs %d profiled instructions (library rebuilt since the capture?)' % (len(lines), len(body)), file=sys.stderr)
si, ii = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
tot, st = collections.Counter(), collections.Counter()
for (loc, txt), r in zip(lines, body):
    tot[loc] += int(r[ii]); st[loc] += int(r[si])
T, S = sum(tot.values()), max(sum(st.values()), 1)
print('# %s: %.1f warp-instructions per unit, %d stall samples' % (pat, T / nenv, S))
byf, sf = collections.Counter(), collections.Counter()
for k, v in tot.items(): byf[k.split(':')[0]] += v
for k, v in st.items(): sf[k.split(':')[0]] += v
for k, v in byf.most_common(): print('%-26s %8.1f instr/unit  %5.1f%% of stall samples' % (k, v / nenv, 100.0 * sf[k] / S))
print('# top lines')
for k, v in tot.most_common(topn): print('%8.1f instr/unit  %5.1f%% stall  %s' % (v / nenv, 100.0 * st[k] / S, k))
