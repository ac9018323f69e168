"""Summarise an ncu report (read with the local ncu CLI, no GPU) into a small tracked JSON under profiles/.

    python tools/ncu_r02.py <report.ncu-rep> <profiles/out.json> <workload tag> "<command that was profiled>"

Per kernel (averaged over its captured launches): duration, DRAM bytes read / written, achieved DRAM GB/s (= bytes / duration,
cold-cache and serialised under ncu: compare with the CUDA-event numbers of tools/datapath_probe.py / bench.py), DRAM and
tensor-pipe utilisation, L1 / L2 hit rates (sector efficiency of the access pattern), SM throughput, launch geometry.
"""
import collections
import csv
import json
import re
import subprocess
import sys

rep, out_path, workload, cmd = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else '')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
to_us = {'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
M = {'dur': 'gpu__time_duration.sum', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
     'tensor': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'sm': 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1hit': 'l1tex__t_sector_hit_rate.pct', 'l2hit': 'lts__t_sector_hit_rate.pct',
     'regs': 'launch__registers_per_thread', 'grid': 'launch__grid_size', 'block': 'launch__block_size',
     'smem': 'launch__shared_mem_per_block_dynamic', 'cyc': 'sm__cycles_elapsed.max'}


def val(r, key, scale=None):
    if M[key] not in idx or r[idx[M[key]]] in ('', 'n/a'):
        return None
    v = float(r[idx[M[key]]].replace(',', ''))
    u = units[idx[M[key]]]
    if scale == 'bytes':
        v *= mult.get(u, 1)
    if scale == 'us':
        v *= to_us.get(u, 1)
    return v


def short(name):
    n = re.sub(r'\(CUtensorMap.*|\(dmf.*|\(const.*|\(tc::.*|\(PatchSrc.*|\(unsigned.*|\(float.*', '', name)
    n = n.replace('void ', '').replace('dmf::tc::', '').replace('dmf::', '').replace('(int)', '').replace('(bool)', '').replace(' ', '')
    for pat, key in ((r'conv_pool4_kernel<64,128', 'conv_pool4_64_128'), (r'conv_pool4_kernel<32,64', 'conv_pool4_32_64'), (r'fuse_rowsum_kernel', 'fuse_rowsum'),
                     (r'ms_stem_map_kernel', 'ms_stem_map'), (r'pan_stem_map_kernel', 'pan_stem_map'), (r'head_dense_kernel', 'head_dense')):
        if re.search(pat, n):
            return key
    return n


agg = collections.OrderedDict()
order = []
for r in rows[2:]:
    k = short(r[idx['Kernel Name']])
    e = agg.setdefault(k, collections.defaultdict(list))
    e['full_name'] = r[idx['Kernel Name']][:160]
    for key, scale in (('dur', 'us'), ('rd', 'bytes'), ('wr', 'bytes'), ('tensor', None), ('dram', None), ('sm', None), ('l1hit', None), ('l2hit', None), ('cyc', None)):
        v = val(r, key, scale)
        if v is not None:
            e[key].append(v)
    e['meta'] = [r[idx[M[k2]]] for k2 in ('regs', 'grid', 'block', 'smem')]
    e['per_launch'].append({'duration_us': round(val(r, 'dur', 'us'), 2), 'dram_read_MB': round(val(r, 'rd', 'bytes') / 1e6, 2),
                            'dram_write_MB': round(val(r, 'wr', 'bytes') / 1e6, 2), 'grid': r[idx[M['grid']]], 'block': r[idx[M['block']]]})
avg = lambda xs: sum(xs) / len(xs) if xs else None
out = {'source': 'ncu --set full --clock-control none; %s' % cmd, 'workload': workload, 'kernels': {}}
for k, e in agg.items():
    n = len(e['dur'])
    byt = avg(e['rd']) + avg(e['wr'])
    out['kernels'][k] = {'captured_launches': n, 'avg_duration_us': round(avg(e['dur']), 2), 'dram_bytes_per_launch': round(byt),
                         'dram_read_bytes_per_launch': round(avg(e['rd'])), 'dram_write_bytes_per_launch': round(avg(e['wr'])),
                         'dram_GBs_during_capture': round(byt / avg(e['dur']) / 1e3, 1),
                         'dram_throughput_pct': round(avg(e['dram']), 1) if e['dram'] else None,
                         'tensor_pipe_active_pct': round(avg(e['tensor']), 1) if e['tensor'] else None,
                         'sm_throughput_pct': round(avg(e['sm']), 1) if e['sm'] else None,
                         'l1_sector_hit_rate_pct': round(avg(e['l1hit']), 1) if e['l1hit'] else None,
                         'l2_sector_hit_rate_pct': round(avg(e['l2hit']), 1) if e['l2hit'] else None,
                         'sm_mhz_during_capture': round(avg(e['cyc']) / avg(e['dur'])) if e['cyc'] else None,
                         'registers': int(e['meta'][0]), 'grid': e['meta'][1], 'block': e['meta'][2], 'smem_dynamic_kb': e['meta'][3],
                         'launches': e['per_launch'] if n > 1 else None, 'kernel': e['full_name']}
    print(k, {kk: vv for kk, vv in out['kernels'][k].items() if kk not in ('launches', 'kernel')})
json.dump(out, open(out_path, 'w'), indent=1)
