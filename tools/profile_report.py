"""Summarise a launch list (ncu --metrics gpu__time_duration.sum --csv) and, optionally, an `ncu --set full` capture
and the bench line of the same visit into profiles/<tag>.md:
    python tools/profile_report.py TAG launches.csv [capture.ncu-rep] [bench.json]"""
import collections, csv, json, os, subprocess, sys

tag, launches = sys.argv[1], sys.argv[2]
rep = next((a for a in sys.argv[3:] if a.endswith('.ncu-rep')), None)
bench = next((a for a in sys.argv[3:] if a.endswith('.json')), None)
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki].split('(')[0].replace('void ', ''), []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in d.values())
lines = ['# %s' % tag, '', 'ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and',
         'serialised: shares, not absolutes, carry over to the timed run).', '',
         '| kernel | launches | mean us | total us | share |', '|---|---|---|---|---|']
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    lines.append('| %s | %d | %.1f | %.1f | %.1f%% |' % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
step = sum(sum(v) for k, v in d.items() if k.startswith('k_step'))
lines += ['', 'Share of the fused step kernel (all its launches) in the captured window: %.1f%%' % (100 * step / tot)]
if rep:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct']
    lines += ['', '## ncu --set full (`%s`)' % os.path.basename(rep), '', '| metric | ' + ' | '.join(r[h.index('Kernel Name')].replace('void ', '').split('(')[0] for r in rr[2:]) + ' | unit |',
              '|---|' + '---|' * (len(rr) - 1)]
    for k in keys:
        if k in h:
            i = h.index(k)
            lines.append('| %s | ' % k + ' | '.join(r[i] for r in rr[2:]) + ' | %s |' % u[i])
if bench and os.path.exists(bench):
    lines += ['', '## bench line of the same visit (not under ncu)', '', '```', open(bench).read().strip(), '```']
os.makedirs('profiles', exist_ok=True)
open(os.path.join('profiles', tag + '.md'), 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:40]))
