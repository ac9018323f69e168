"""Where the warps of a kernel wait: mbarrier try-wait sites (with the samples of the spin branch behind them) and the opcode mix of the stall
samples, from the source page of an ncu report (read locally, no GPU).    python tools/ncu_roles.py <report.ncu-rep> <kernel regex> [launch index]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
kid = '::regex:%s:%s' % (rx, sys.argv[3] if len(sys.argv) > 3 else '1')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', kid], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = None, []
for r in rows:
    if r and r[0] == 'Kernel Name' and hdr is None:
        print(r[1][:150])
    if r and r[0] == 'Address':
        if hdr is None:
            hdr = r
            continue
        break
    if hdr and len(r) == len(hdr):
        data.append(r)
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
tot = sum(int(r[i_s]) for r in data)
print('stall samples %d, SASS instructions %d, warp instructions executed %d' % (tot, len(data), sum(int(r[i_ex]) for r in data)))
print('-- mbarrier waits: address, executions, samples at the wait + its spin branch')
for i, r in enumerate(data):
    if 'PHASECHK' in r[i_src] and int(r[i_ex]) > 1000:
        smp = int(r[i_s]) + int(data[i + 1][i_s])
        if smp * 200 > tot:
            print('  %s %9s %7d (%4.1f %%)  %s' % (r[0][-5:], r[i_ex], smp, 100.0 * smp / tot, r[i_src].strip()[:70]))
byop = collections.Counter()
for r in data:
    t = r[i_src].strip().split()
    byop[(t[1] if t[0].startswith('@') else t[0]).split('.')[0]] += int(r[i_s])
print('-- samples by opcode:', ', '.join('%s %d' % kv for kv in byop.most_common(12)))
