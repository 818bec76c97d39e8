"""Top stall-sample SASS instructions of a kernel in an ncu report, with the CUDA source line of each (from
nvdisasm -g of the built library).  python tools/hot_sass.py <report.ncu-rep> <kernel regex> [top]"""
import csv, re, subprocess, sys, os, tempfile, glob
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [r for r in rows if '# Samples' in r][0]
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith('0x')]
seen, uniq = set(), []
for r in data:                       # the CSV lists every instruction twice
    if r[0] in seen:
        continue
    seen.add(r[0]); uniq.append(r)
iS, iI, isrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
tot = sum(int(r[iS]) for r in uniq if r[iS].isdigit())
# line info
d = tempfile.mkdtemp()
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'sucem-fem_b200', 'lib', 'libsucem_b200.so')
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=d, capture_output=True)
lines = {}
for f in glob.glob(os.path.join(d, '*.cubin')):
    txt = subprocess.run(['nvdisasm', '-g', '-c', f], capture_output=True, text=True).stdout
    if kern.split('|')[0] not in txt:
        continue
    cur, infn = None, False
    for l in txt.splitlines():
        if '.text.' in l and 'section' in l:
            infn = re.search(kern, l) is not None
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m and infn:
            lines[int(m.group(1), 16)] = cur
base = min(int(r[0], 16) for r in uniq)
print('total samples', tot)
byline = {}
for r in uniq:
    if not r[iS].isdigit():
        continue
    li = lines.get(int(r[0], 16) - base)
    byline[li] = byline.get(li, 0) + int(r[iS])
print('--- by source line')
for li, n in sorted(byline.items(), key=lambda kv: -kv[1])[:top_n]:
    print('%6.2f%%  %s' % (100.0 * n / tot, li))
print('--- by instruction')
for r in sorted([r for r in uniq if r[iS].isdigit()], key=lambda r: -int(r[iS]))[:top_n]:
    print('%6.2f%%  x%-9s %-22s %s' % (100.0 * int(r[iS]) / tot, r[iI], lines.get(int(r[0], 16) - base), r[isrc].strip()[:80]))
