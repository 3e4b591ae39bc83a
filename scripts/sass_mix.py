"""Aggregate an `ncu --page source --print-source sass --csv` dump: opcode mix,
SIMT efficiency per opcode, stall reasons.  usage: sass_mix.py file.csv [kernel#]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; blocks.append(cur); continue
    if cur is not None: cur['rows'].append(r)
b = blocks[which]; H = b['rows'][0]; data = [r for r in b['rows'][1:] if len(r) == len(H)]
ix = {h: i for i, h in enumerate(H)}
I = lambda r, k: int(r[ix[k]] or 0)
tot_inst = sum(I(r, 'Instructions Executed') for r in data)
tot_thr = sum(I(r, 'Thread Instructions Executed') for r in data)
print(b['name'][:80]); print("sass lines", len(data), "warp inst %.3e" % tot_inst, "avg active threads %.2f" % (tot_thr / tot_inst))
op, opthr = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    o = m.group(2).split('.')[0] if m else '?'
    op[o] += I(r, 'Instructions Executed'); opthr[o] += I(r, 'Thread Instructions Executed')
for o, c in op.most_common(30):
    print(f"  {o:10s} {100*c/tot_inst:5.1f}% of warp-inst   avg threads {opthr[o]/max(c,1):5.1f}")
st = collections.Counter()
for r in data:
    for k in ix:
        if k.startswith('stall_') and 'Not Issued' not in k: st[k] += I(r, k)
ts = sum(st.values())
print("stalls:", {k: round(100 * v / ts, 1) for k, v in st.most_common(9)})
