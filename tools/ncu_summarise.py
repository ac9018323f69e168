"""Turn an ncu report (.ncu-rep, read with the local ncu CLI) and a launch-list CSV into the small
tracked summaries under profiles/.  Usage: python tools/ncu_summarise.py <rep> <launches.csv> <round-tag> <chunk_pixels> [<profiled command> [<launch-list command>]]"""
import collections, csv, json, re, subprocess, sys
rep, launches, tag, chunk = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
prof_cmd = sys.argv[5] if len(sys.argv) > 5 else 'tools/perf_probe.py 1000 1000 %d 40' % chunk
list_cmd = sys.argv[6] if len(sys.argv) > 6 else 'python bench.py --steps 1 --warmup 3 --no-cpu-baseline'

def short(name):
    name = re.sub(r'\(CUtensorMap.*|\(dmf.*|\(const.*|\(tc::.*|\(PatchSrc.*', '', name)
    return name.replace('void ', '').replace('dmf::tc::', 'tc::').replace('dmf::', '').replace('(int)', '').replace('(bool)', '').replace(' ', '')

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
to_us = {'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
dur_scale = to_us[units[idx['gpu__time_duration.sum']]]
K = {'dur': 'gpu__time_duration.sum', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
     'tensor': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'regs': 'launch__registers_per_thread', 'grid': 'launch__grid_size', 'block': 'launch__block_size', 'smem': 'launch__shared_mem_per_block_dynamic',
     'cyc': 'sm__cycles_elapsed.max'}
agg = collections.OrderedDict()
for r in rows[2:]:
    e = agg.setdefault(short(r[idx['Kernel Name']]), collections.defaultdict(list))
    e['dur'].append(float(r[idx[K['dur']]]) * dur_scale)
    e['bytes'].append(float(r[idx[K['rd']]]) * mult[units[idx[K['rd']]]] + float(r[idx[K['wr']]]) * mult[units[idx[K['wr']]]])
    e['rd'].append(float(r[idx[K['rd']]]) * mult[units[idx[K['rd']]]])
    e['tensor'].append(float(r[idx[K['tensor']]])); e['dram'].append(float(r[idx[K['dram']]]))
    e['mhz'].append(float(r[idx[K['cyc']]]) / (float(r[idx[K['dur']]]) * dur_scale))
    e['meta'] = [r[idx[K[k]]] for k in ('regs', 'grid', 'block', 'smem')]
out = {'source': 'ncu --set full --clock-control none --import-source on; %s; round %s' % (prof_cmd, tag),
       'workload': 'c2', 'chunk_pixels': chunk, 'kernels': {}}
for k, e in agg.items():
    n = len(e['dur'])
    out['kernels'][k] = {'captured_launches': n, 'avg_duration_us': round(sum(e['dur']) / n, 2), 'dram_bytes_per_launch': round(sum(e['bytes']) / n),
                         'dram_read_bytes_per_launch': round(sum(e['rd']) / n), 'tensor_pipe_active_pct': round(sum(e['tensor']) / n, 1),
                         'dram_throughput_pct': round(sum(e['dram']) / n, 1), 'sm_mhz_during_capture': round(sum(e['mhz']) / n),
                         'registers': int(e['meta'][0]), 'grid': e['meta'][1], 'block': e['meta'][2], 'smem_dynamic_kb': e['meta'][3]}
    print(k, out['kernels'][k])
json.dump(out, open('profiles/%s_ncu_summary.json' % tag, 'w'), indent=1)

rows = list(csv.DictReader(l for l in open(launches) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    v, u = float(r['Metric Value']), r['Metric Unit']
    us = v / 1000 if u.startswith('n') else v if u.startswith('u') else v * 1000
    a = agg.setdefault(short(r['Kernel Name']), [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
lines = ['# ncu --metrics gpu__time_duration.sum --clock-control none  ' + list_cmd,
         '# round %s, B200, chunk / band parameter %d; per-launch times are cold-cache and serialised: compare SHARES with bench.py roofline.stage_ms' % (tag, chunk),
         '%-48s %8s %12s %7s' % ('kernel', 'launches', 'total_us', 'share')]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append('%-48s %8d %12.1f %6.1f%%' % (k, n, us, 100 * us / tot))
open('profiles/%s_launches_summary.txt' % tag, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
