"""Join an ncu SASS source page (csv) with nvdisasm -g line info: instruction share and stall samples per CUDA source line.
usage: sass_lines.py <ncu_source_page.csv> <nvdisasm -g -c output> <kernel mangled substring> [kernel index]"""
import csv, re, sys
csvp, sassp, kname = sys.argv[1:4]
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(csvp)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
blk = rows[starts[kidx] + 1:(starts[kidx + 1] if kidx + 1 < len(starts) else None)]
hdr = blk[0]
ie, ss = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
inst = [(r[1].strip(), int(r[ie] or 0), int(r[ss] or 0)) for r in blk[1:] if len(r) > ie]
# nvdisasm: walk the kernel's section, track current line
lines = open(sassp).read().split('\n')
cur = None; in_k = False; seq = []
for l in lines:
    if l.startswith('.section') or l.lstrip().startswith('.section'):
        in_k = ('.text.' in l and kname in l)
        continue
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(2)))
print('ncu insts', len(inst), 'nvdisasm insts', len(seq), file=sys.stderr)
agg = {}
n = min(len(inst), len(seq))
for (src, cnt, smp), (loc, txt) in zip(inst[:n], seq[:n]):
    a = agg.setdefault(loc, [0, 0]); a[0] += cnt; a[1] += smp
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print('total warp-insts', ti, 'samples', ts)
for loc, a in sorted(agg.items(), key=lambda kv: (kv[0] is None, kv[0])):
    if a[0] > 0.003 * ti or a[1] > 0.005 * ts:
        print(loc, 'inst %.1f%%' % (100 * a[0] / ti), 'stall %.1f%%' % (100 * a[1] / ts))
