"""Correlate an ncu SASS page (per-instruction executed counts) with nvdisasm -g
line info: warp-instructions executed per source line.
usage: sass_by_line.py sass.csv dis.txt <mangled kernel substring> [top]"""
import collections, csv, re, sys
sass_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
H = rows[1]; data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) == len(H): data.append(r)
ix = {h: i for i, h in enumerate(H)}
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l)
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith('//---------------------'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): seq.append((cur, l.split('*/')[1].strip()))
print("sass rows", len(data), "disasm instr", len(seq))
agg = collections.Counter(); thr = collections.Counter()
n = min(len(data), len(seq))
for k in range(n):
    agg[seq[k][0]] += int(data[k][ix['Instructions Executed']]); thr[seq[k][0]] += int(data[k][ix['Thread Instructions Executed']])
tot = sum(agg.values())
src = {}
for key, c in agg.most_common(top):
    f, ln = key if key else ('?', 0)
    try:
        if f not in src: src[f] = open('/root/repo/hermespy-rt_b200/csrc/' + f).read().split('\n')
        text = src[f][ln - 1].strip()[:90]
    except Exception: text = ''
    print(f"{100*c/tot:5.1f}%  thr {thr[key]/max(c,1):4.1f}  {f}:{ln}  {text}")
