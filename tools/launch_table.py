"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_table.py file.csv"""
import collections, csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[ki][:64] + ' ' + r[gi], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in d.values())
    print(f)
    for k, v in d.items():
        print('  %-84s n=%3d mean %8.1f us  share %5.1f%%' % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
