"""ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --csv --log-file ...`) -> per-kernel totals and shares.
    python tools/launches_summary.py gpurun_out/r02f_launches.csv profiles/r02_launches_summary.txt "<command that was profiled>" """
import collections, csv, re, sys
src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
def short(name):
    n = re.sub(r'\(CUtensorMap.*|\(dmf.*|\(const.*|\(tc::.*|\(PatchSrc.*|\(unsigned.*|\(float.*|\(long.*|\(int.*', '', name)
    return n.replace('void ', '').replace('dmf::tc::', 'tc::').replace('dmf::', '').replace(' ', '')
agg = collections.OrderedDict()
for r in rows:
    v, u = float(r['Metric Value'].replace(',', '')), r['Metric Unit']
    us = v / 1000 if u.startswith('n') else v if u.startswith('u') else v * 1000
    a = agg.setdefault(short(r['Kernel Name']), [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
lines = ['# ncu --metrics gpu__time_duration.sum --clock-control none -c 600   ' + cmd,
         '# the first %d launches of the run (scene preparation, warm-up and timed steps, e2e passes); per-launch times are cold-cache and' % len(rows),
         '# serialised: compare SHARES with bench.py roofline.stage_ms, not absolutes',
         '%-52s %8s %12s %7s' % ('kernel', 'launches', 'total_us', 'share')]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append('%-52s %8d %12.1f %6.1f%%' % (k[:52], n, us, 100 * us / tot))
open(dst, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
