"""Summarise one GPU visit (tools/gpu_final.sh) into profiles/: launch list shares, raw ncu metrics of the
fused kernel, stall/opcode mix.  python tools/profile_summary.py <tag>"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1]
out = os.path.join('profiles', tag)
os.makedirs('profiles', exist_ok=True)
rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki].split('(')[0], []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in d.values())
lines = ['# %s: ncu launch list (gpu__time_duration.sum, --clock-control none) of `python bench.py --steps 20 --warmup 3`' % tag, '',
         '| kernel | launches | mean us | total us | share |', '|---|---|---|---|---|']
for k, v in d.items():
    lines.append('| %s | %d | %.1f | %.1f | %.1f%% |' % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
step = [sum(v) for k, v in d.items() if 'k_step' in k]
steady = [k for k in d if k.startswith('k_fill') or 'k_step' in k or k.startswith('k_fs') or k.startswith('k_red')]
lines += ['', 'Share of the fused kernel among the per-step kernels: %.1f%%' % (100 * sum(step) / max(1e-9, sum(sum(d[k]) for k in steady)))]
raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_step.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second']
lines += ['', '## ncu --set full, fused kernel (launch 1 of the capture)', '', '| metric | value | unit |', '|---|---|---|']
traffic = None
for r in rr[2:3]:
    for k in keys:
        if k in h:
            i = h.index(k); lines.append('| %s | %s | %s |' % (k, r[i], u[i]))
    def val(k):
        i = h.index(k); v = float(r[i].replace(',', '')); un = u[i]
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(un, 1)
    traffic = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
src = subprocess.run(['ncu', '-i', 'gpurun_out/prof_step.ncu-rep', '--page', 'source', '--csv', '--kernel-id', ':::1'], capture_output=True, text=True).stdout
open('/tmp/_src.csv', 'w').write(src)
mix = subprocess.run([sys.executable, 'tools/ncu_src.py', '/tmp/_src.csv'], capture_output=True, text=True).stdout
lines += ['', '## warp stall samples and executed opcode mix (source page)', '', '```', mix.strip(), '```']
if os.path.exists('gpurun_out/bench.json'):
    lines += ['', '## bench line of the same visit (not under ncu)', '', '```', open('gpurun_out/bench.json').read().strip(), '```']
open(out + '.md', 'w').write('\n'.join(lines) + '\n')
tj = os.path.join('profiles', 'traffic.json')
t = json.load(open(tj)) if os.path.exists(tj) else {}
t['cfg2'] = {'dram_bytes_per_launch': traffic, 'source': tag + ' ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch'}
json.dump(t, open(tj, 'w'), indent=1)
print(open(out + '.md').read()[:3000])
