"""Aggregate an `ncu --page source --csv` dump: stall reasons and opcode mix (SASS table only)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; tot = collections.Counter(); opc = collections.Counter(); opn = collections.Counter(); ops = collections.Counter()
totinst = 0; nstat = 0; table = 0; top = []
for r in rows:
    if len(r) > 5 and r[0] == 'Address':
        hdr = r; table += 1; continue
    if len(r) > 5 and r[0] in ('#', 'Line'):  # CUDA-C view
        hdr = None; continue
    if hdr is None or len(r) < len(hdr) or table != 1: continue
    si = hdr.index('Source'); ni = hdr.index('# Samples'); ei = hdr.index('Instructions Executed')
    src = r[si].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2).split('.')[0] if m else '?'
    n = int(r[ni]); e = int(r[ei])
    opc[op] += e; ops[op] += n; opn[op] += 1; totinst += e; nstat += 1
    top.append((n, src))
    for k, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h: tot[h] += int(r[k])
print('tables', table, 'total warp-instr executed', totinst, 'static instr', nstat)
S = sum(tot.values())
for k, v in tot.most_common(): print('%-28s %6d %5.1f%%' % (k, v, 100 * v / S))
print()
for k, v in opc.most_common(28): print('%-10s exec %10d (%4.1f%%) static %4d samples %6d (%4.1f%%)' % (k, v, 100 * v / totinst, opn[k], ops[k], 100*ops[k]/S))
print()
for n, s in sorted(top, reverse=True)[:25]: print(n, s)
