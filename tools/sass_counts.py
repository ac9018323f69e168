"""Blackwell evidence from the built library: per kernel, how many tcgen05 / TMEM / TMA instructions its SASS holds.

    python tools/sass_counts.py [path/to/libdmf_b200.so]  ->  profiles/r02_sass_counts.txt

SASS mnemonics (B200_PROFILING.md): tcgen05.mma -> UTC*MMA (UTCHMMA for kind::f16), tcgen05.ld -> LDTM, TMA tensor loads ->
UTMALDG, bulk copies (cp.async.bulk, both directions) -> UBLKCP, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, legacy mma.sync -> HMMA.
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, 'dual-modal-fusion_b200', 'dmf', 'libdmf_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
PAT = [('UTC*MMA (tcgen05.mma)', r'\bUTC[A-Z]*MMA'), ('LDTM (tcgen05.ld)', r'\bLDTM'), ('UTMALDG (TMA tensor load)', r'\bUTMALDG'),
       ('UBLKCP (bulk copy)', r'\bUBLKCP'), ('UTCBAR (tcgen05.commit)', r'\bUTCBAR'), ('SYNCS (mbarrier)', r'\bSYNCS'),
       ('F2FP (pack cvt)', r'\bF2FP'), ('HMMA (legacy mma.sync)', r'\bHMMA')]
rows, cur, i = [], None, -1
for line in sass.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        i += 1
        cur = collections.Counter()
        rows.append((names[i], cur))
        continue
    if cur is None:
        continue
    for key, pat in PAT:
        if re.search(pat, line):
            cur[key] += 1


def short(n):
    n = re.sub(r'\(.*', '', n).replace('void ', '').replace('dmf::tc::', 'tc::').replace('dmf::', '')
    return n.replace('(bool)', '').replace('(int)', '').replace(' ', '')


out = ['# cuobjdump -sass %s  (sm_100a); instruction counts per kernel, kernels without any of them omitted' % os.path.relpath(lib, REPO),
       '%-64s %s' % ('kernel', '  '.join('%s' % k.split(' ')[0] for k, _ in PAT))]
tot = collections.Counter()
for n, c in sorted(rows, key=lambda r: -r[1]['UTC*MMA (tcgen05.mma)']):
    if not any(c[k] for k, _ in PAT[:5]):
        continue
    out.append('%-64s %s' % (short(n)[:64], '  '.join('%*d' % (len(k.split(' ')[0]), c[k]) for k, _ in PAT)))
    tot.update(c)
out.append('%-64s %s' % ('TOTAL', '  '.join('%*d' % (len(k.split(' ')[0]), tot[k]) for k, _ in PAT)))
out.append('# legend: ' + '; '.join(k for k, _ in PAT))
text = '\n'.join(out) + '\n'
open(os.path.join(REPO, 'profiles', 'r02_sass_counts.txt'), 'w').write(text)
print(text)
