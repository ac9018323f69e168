#!/usr/bin/env python
"""bench.py — scene pixels classified / second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3]

One "step" = one whole-scene classification pass (every pixel: co-registered MS/PAN window -> GMFNet -> argmax -> confusion
matrix + label map) over the synthetic scene of the workload.  The SAME workload at every N, so the 1/2/4/8 numbers are one
strong-scaling curve:
  C3  Hohhot-scale  MS 4x2001x2101 / PAN 8004x8404, 11 classes (+background = 12): the configuration BASELINE.json quotes the
      scaling metric on; at N > 1 row-band sharded with one int64 C*C all-reduce per step.  It fits one GPU.
  C2  Xi'an-scale   MS 4x1000x1000 / PAN 4000x4000, 12 classes: `--workload c2`, and a `secondary` entry of the default N = 1 run.
Scene = oracle.synthetic_scene_structured (labels depend on the rasters); network = the fitted GMFNet of oracle/fitted_net.py
(seed-3407 convolutions, calibrated BatchNorm, fitted head), so predictions vary and OA / AA / Kappa are not degenerate.
`value`: CUDA events around each step with the scene resident in HBM.  `e2e`: the public API (dmf.ScenePipeline) from pinned
HOST rasters — H2D + range + normalise/pad + inference + all-reduce + D2H of label band and matrix + OA/AA/Kappa every step,
software-pipelined (scene i+1 uploads while scene i is classified).  `--impl reference` times the reference's own CPU
objects (oracle/_ref through oracle/ref_runner.py; the oracle port if that copy is absent) on all host cores.
oracle/ in the b200 arm: it only GENERATES INPUTS before any timing starts — the synthetic rasters (numpy) and the weight values of
the fitted net, loaded into the product's own model.gmfnet.Net with load_state_dict like a checkpoint.  No oracle code computes
anything on the product path or inside a timed region; the oracle's compute runs in the `cpu_baseline` / `--impl reference` legs only.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))

import numpy as np
import torch

WORKLOADS = {
    'c1': dict(name='C1 synthetic MS 4x128x128 / PAN 512x512, p=16, 7 classes', H=128, W=128, classes=7),
    'c2': dict(name="C2 Xi'an-scale synthetic MS 4x1000x1000 / PAN 4000x4000, p=16, 12 classes", H=1000, W=1000, classes=12),
    'c3': dict(name='C3 Hohhot-scale synthetic MS 4x2001x2101 / PAN 8004x8404, p=16, 11 classes', H=2001, W=2101, classes=11),
}
P = 16
METRIC, UNIT = 'scene pixels classified/sec', 'px/s'


def make_config(wl_key):
    """The `config` object: identical in both arms (--impl b200 / reference) and at every N."""
    wl = WORKLOADS[wl_key]
    return {'workload': wl['name'], 'patch_size': P, 'classes_with_background': wl['classes'] + 1, 'pixels': wl['H'] * wl['W'],
            'scene': 'structured synthetic rasters, uint16 11-bit, labels depend on the rasters (oracle/fitted_net.scene: regions of 64 MS pixels, seeds 0 / 1)',
            'net': 'GMFNet, seed-3407 convolutions + calibrated BatchNorm + fitted head (oracle/fitted_net.py): predictions vary, Kappa > 0',
            'sharding': 'row bands over the ranks, scene replicated per rank, one int64 CxC all-reduce per step',
            'l2': 'explicit 256 MiB L2 flush between timed steps; the per-step intermediates (GBs) exceed L2 too'}


def peaks():
    try:
        return json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json'))), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock + throttle reasons sampled every 10 ms during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.index = index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40,
                     'hw_power_brake': 0x80}
            while not self._stop.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.reasons.update(k for k, bit in names.items() if r & bit)
                time.sleep(0.01)
        except Exception as e:               # clocks are evidence, not a dependency
            self.reasons.add('unavailable: %s' % type(e).__name__)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def workload_scene(wl_key):
    from oracle import fitted_net
    return fitted_net.scene(wl_key)


def cpu_reference_run(wl_key, budget_s, steps=1, warmup=0, scene=None):
    """The reference's CPU path on ALL host cores; returns (px/s, cores, kind, description, per-step ms).
    kind "reference": the reference's own Solver / DataLoader / dataset_dual / aa_oa objects (oracle/_ref, vendored by
    __graft_entry__.build()); kind "port": oracle/ref_pipeline.py when that copy is absent."""
    from oracle import fitted_net, ref_runner
    wl = WORKLOADS[wl_key]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)                          # torchrun exports OMP_NUM_THREADS=1: undo it explicitly
    ms, pan, label = scene if scene is not None else workload_scene(wl_key)
    state = fitted_net.fitted_state(wl_key)

    def factory(args):
        net = fitted_net.base_net(wl_key)
        net.load_state_dict(state)
        return net

    if ref_runner.available():
        kind = 'reference'
        run = ref_runner.RefRun(ms, pan, label, P, wl['classes'], factory)
        classify = run.classify
        prep_s = run.prep_s
    else:
        kind = 'port'
        from oracle.ref_pipeline import RefPipeline
        pipe = RefPipeline(ms, pan, label, P, wl['classes'] + 1, net=factory(None))
        order, state_ = np.arange(wl['H'] * wl['W']), {'c': 0}

        def classify(b):
            M, lm, done, secs = pipe.classify(order[state_['c']:state_['c'] + 200000], batch_size=300, budget_s=b)
            state_['c'] = (state_['c'] + done) % max(1, order.size - 200000)
            return M, lm, done, secs
        prep_s = pipe.prep_s
    px, secs_total, M_total = 0, 0.0, None
    for s in range(warmup + steps):
        M, _, done, secs = classify(budget_s if s >= warmup else min(budget_s, 3.0))
        if s >= warmup:
            px += done
            secs_total += secs
            M_total = M if M_total is None else M_total + M
    desc = ('%d px of the scene in %.1f s (%s: Solver.__init__ + dataloader(), DataLoader bs=300 num_workers=0 over '
            'dataset_dual.__getitem__, fp32 Net on the CPU with %d torch threads, per-sample confusion + label-map loops); '
            'scene prep (to_tensor, data_padding, split_data_old) %.1f s not counted'
            % (px, secs_total, 'the reference\'s own objects from oracle/_ref' if kind == 'reference' else 'oracle port', torch.get_num_threads(), prep_s))
    return px / secs_total, torch.get_num_threads(), kind, desc, 1e3 * secs_total / max(1, steps)


def fitted_product_net(wl_key, dev, **b200):
    from model.gmfnet import Net
    from oracle import fitted_net
    net = Net(dict(fitted_net.cfg_for(wl_key), b200=b200))
    net.load_state_dict(fitted_net.fitted_state(wl_key))
    return net.to(dev).eval()


def train_step_metric(dev, world, batch=512, steps=30, warmup=5):
    """BASELINE.json configs[3]: IHS-input training step, `batch` patches per GPU (data-parallel, weak scaling): the IHS product
    computed on the device straight into the scene (dmf_scene_set_mspan_ihs) -> K1 tri-gather from the resident scene -> native
    forward (train-mode BatchNorm) -> CrossEntropyLoss -> native backward -> one flat-gradient all-reduce -> FusedAdam.  CUDA events,
    max over ranks."""
    import random
    import torch.distributed as dist
    import dmf
    from oracle import dmf_oracle as orc
    from image_convert.IHS import draw_offsets
    from model.gmfnet import Net
    C, Hs, Ws = 12, 400, 400
    ms, pan, label = orc.synthetic_scene_structured(Hs, Ws, C - 1, seed=0, label_seed=1)
    random.seed(7)
    offs = draw_offsets(Hs, Ws, 4, 4)
    sc = dmf.Scene.from_raw(ms, pan, P, dev)
    sc.set_labels(label)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_d = torch.from_numpy(ms.view(np.int16)).to(dev)
    pan_d = torch.from_numpy(pan.view(np.int16)).to(dev)
    offs_d = torch.from_numpy(offs).to(dev)
    sc.set_mspan_ihs(ms_d, pan_d, offs_d)
    e0.record()
    sc.set_mspan_ihs(ms_d, pan_d, offs_d)
    e1.record()
    torch.cuda.synchronize()
    ihs_ms = e0.elapsed_time(e1)
    torch.manual_seed(0)
    net = Net({'Categories_Number': C, 'patch_size': P, 'schedule': {'activate': 'Relu'}, 'b200': {'max_train_batch': batch}}).to(dev).train()
    opt = dmf.FusedAdam(net.parameters(), lr=1e-3)
    labelled = torch.from_numpy(np.flatnonzero(label.reshape(-1) != 0)).to(dev)
    g = torch.Generator(device=dev).manual_seed(1 + (dist.get_rank() if world > 1 else 0))
    batches = [labelled[torch.randint(0, labelled.numel(), (batch,), device=dev, generator=g)] for _ in range(8)]
    first = None
    for i in range(warmup):
        loss = net.train_step_scene(sc, batches[i % 8], opt, use_mspan=True)
        first = float(loss) if first is None else first
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = dmf.launch_count()
    e0.record()
    for i in range(steps):
        loss = net.train_step_scene(sc, batches[i % 8], opt, use_mspan=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t)
    flops = 3 * net.native().flops_per_patch * batch * world
    return {'workload': 'C4 IHS-input training step, batch %d per GPU, p=16, 12 classes, data-parallel x%d' % (batch, world),
            'ms_per_step': ms_step, 'patches_per_s': batch * world / ms_step * 1e3, 'algorithmic_TFLOPs': flops / ms_step / 1e9,
            'kernels_per_step': (dmf.launch_count() - l0) / steps, 'loss_first_step': first, 'loss_after': float(loss), 'steps': steps,
            'ihs_product_on_device_ms': ihs_ms, 'ihs_scene': '%dx%d raw uint16 -> padded fp32 MSPAN in the scene, no host round trip' % (Hs, Ws),
            'collective': 'one NCCL all-reduce of the flat fp32 gradient (%.1f MB) per step' % (net.trainer().flat_grad.numel() * 4 / 1e6) if world > 1 else None}


def main():
    # stdout carries exactly one JSON line: park the real fd and point fd 1 at stderr while libraries
    # (NCCL's version banner, tqdm, ...) are active
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + '\n').encode())

    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c3', choices=list(WORKLOADS))
    ap.add_argument('--max-batch', type=int, default=16384)
    ap.add_argument('--mode', default='dense', choices=['dense', 'patch'],
                    help='whole-scene algorithm: scene-dense maps (default) or the per-patch kernels')
    ap.add_argument('--band', type=int, default=512, help='anchor rows per pass of the dense path')
    ap.add_argument('--cpu-budget-s', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the secondary C4 training-step measurement')
    ap.add_argument('--no-secondary', action='store_true', help='skip every secondary measurement (C2, per-patch path, C5 sweep, C4 step)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    wl_key = args.workload
    wl = WORKLOADS[wl_key]
    config = make_config(wl_key)

    if args.impl == 'reference':
        if rank != 0:
            return
        per_step = max(2.0, min(args.cpu_budget_s, 120.0 / max(1, args.steps + args.warmup)))
        v, cores, kind, desc, ms_step = cpu_reference_run(wl_key, per_step, steps=args.steps, warmup=min(args.warmup, 1))
        emit(({'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
               'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
               'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
               'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': desc},
               'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch.distributed as dist
    import dmf
    from model.gmfnet import Net
    from indicators.kappa import aa_oa
    from solver.mainsolver import row_band

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local_rank)
    dev = 'cuda:%d' % local_rank
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    C = wl['classes'] + 1
    H, W = wl['H'], wl['W']
    ms, pan, label = workload_scene(wl_key)
    net = fitted_product_net(wl_key, dev, max_batch=args.max_batch)
    handle = net.native()
    handle.set_dense(args.mode == 'dense', args.band)
    details = {'algorithm': ('scene-dense: every layer evaluated once per scene position and border class (csrc/dense.cu), bands of %d rows' % args.band
                             if args.mode == 'dense' else 'per-patch kernels, chunks of %d pixels' % args.max_batch)}
    r0, r1 = row_band(H, rank, world)
    npix_total = H * W

    # ---- resident scene for the device-timed arm
    scene = dmf.Scene.from_raw(ms, pan, P, dev)
    scene.set_labels(label)
    pred_map = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        cm.zero_()
        handle.infer_scene(scene, r0, r1, pred_map=pred_map, cm=cm)
        if world > 1:
            dist.all_reduce(cm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = dmf.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        for a, b in ev:
            flush.fill_(1)                      # L2 flush, outside the event pair
            a.record()
            step()
            b.record()
        barrier()
    launches = dmf.launch_count() - launches0
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    value = npix_total * args.steps / (total_ms / 1e3)
    cm_host = cm.cpu().numpy().astype(np.float64)
    assert cm_host.sum() == npix_total, 'confusion matrix does not cover the scene'

    # ---- e2e: the public API from pinned host rasters, every copy inside the timed region; 2-deep software pipeline
    ms_pin = torch.from_numpy(ms.view(np.int16)).pin_memory()
    pan_pin = torch.from_numpy(pan.view(np.int16)).pin_memory()
    lab_pin = torch.from_numpy(label).pin_memory()
    pipe = dmf.ScenePipeline(handle, H, W, P, r0, r1)

    def metrics_of(ticket):
        _, cm_h = pipe.result(ticket)
        with open(os.devnull, 'w') as null, _redirect(null):
            return aa_oa(cm_h.numpy().astype(np.float64)), cm_h

    def e2e_run(n, pipelined):
        prev, out = None, None
        for _ in range(n):
            t = pipe.submit(ms_pin, pan_pin, lab_pin)
            if not pipelined:
                out = metrics_of(t)
            else:
                if prev is not None:
                    out = metrics_of(prev)
                prev = t
        if prev is not None:
            out = metrics_of(prev)
        return out

    e2e_run(2, True)
    barrier()
    t0 = time.perf_counter()
    result, cm_e2e = e2e_run(args.steps, True)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = npix_total * args.steps / float(e2e_s)
    assert np.array_equal(cm_e2e.numpy(), cm.cpu().numpy()), 'e2e arm (band upload) and resident-scene arm disagree'
    barrier()
    t0 = time.perf_counter()
    e2e_run(3, False)
    barrier()
    lat_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lat_s, op=dist.ReduceOp.MAX)
    e2e_latency_ms = float(lat_s) / 3 * 1e3
    io_bytes = torch.tensor([pipe.h2d_bytes, pipe.d2h_bytes], dtype=torch.int64, device=dev)      # summed over the ranks
    if world > 1:
        dist.all_reduce(io_bytes)
    h2d_total, d2h_total = (int(v) for v in io_bytes.tolist())

    # ---- roofline of the dominant kernel: per-stage device events inside the library (one extra pass)
    pk, pk_src = peaks()
    n_local = (r1 - r0) * W
    conv = lambda cin, cout, k, h: 2 * cin * cout * k * k * h * h

    def ncu_traffic(tag, key, want_wl):
        try:
            prof = json.load(open(os.path.join(REPO, 'profiles', tag)))
            if want_wl == prof.get('workload'):
                return prof['kernels'].get(key, {}).get('dram_bytes_per_launch')
        except Exception:
            pass
        return None

    def fracs(achieved):
        return {'frac': achieved / pk['bf16_tflops'], 'frac_of_burst': achieved / pk['bf16_tflops'],
                'frac_of_sustained': achieved / pk['bf16_tflops_sustained'], 'peak': pk['bf16_tflops'],
                'peak_sustained': pk['bf16_tflops_sustained'],
                'peak_source': pk_src + ': burst cuBLAS bf16 figure (the kernel is timed in one isolated pass); the sustained figure beside it'}

    def patch_roofline(h, sc, rows0, rows1, Wd):
        h.set_timing(True)
        h.infer_scene(sc, rows0, rows1)
        stage = h.get_timing()
        h.set_timing(False)
        n_loc = (rows1 - rows0) * Wd
        n_chunks = -(-n_loc // args.max_batch)
        kernels = {   # kernel instance -> (stage keys, algorithmic FLOPs per pixel, launches per chunk)
            'conv_tc_kernel<64,128,9,pool,G=3> (ms2 + pan3)': (['conv_ms2', 'conv_pan3'], 2 * conv(64, 128, 3, P), 2),
            'conv_rowpair_kernel<32,G=3> (pan2, 32->64)': (['conv_pan2'], conv(32, 64, 3, 2 * P), 1),
            'conv_tc_kernel<256,128,1,G=2,gap> (fuse + pooling)': (['conv_fuse'], conv(256, 128, 1, P // 2), 1),
        }
        name, (keys, fl_px, per_chunk) = max(kernels.items(), key=lambda kv: sum(stage[k] for k in kv[1][0]))
        k_ms = sum(stage[k] for k in keys)
        achieved = fl_px * n_loc / (k_ms / 1e3) / 1e12
        conv_fl = 2 * conv(64, 128, 3, P) + conv(32, 64, 3, 2 * P) + conv(256, 128, 1, P // 2)
        conv_ms = sum(stage[k] for k in ('conv_ms2', 'conv_pan2', 'conv_pan3', 'conv_fuse'))
        roof = {'bound': 'tensor', 'kernel': name, 'achieved': achieved, 'unit': 'TFLOP/s', **fracs(achieved), 'traffic': None,
                'avg_launch_ms': k_ms / (n_chunks * per_chunk), 'launches': n_chunks * per_chunk,
                'flops_per_launch': fl_px * n_loc / (n_chunks * per_chunk),
                'stage_ms': {k: round(v, 3) for k, v in stage.items()},
                'whole_net_tflops': h.flops_per_patch * n_loc / (stage['total'] / 1e3) / 1e12}
        a = conv_fl * n_loc / (conv_ms / 1e3) / 1e12
        util = {'achieved_TFLOPs': a, 'frac_of_sustained_peak': a / pk['bf16_tflops_sustained'], 'frac_of_burst_peak': a / pk['bf16_tflops']}
        return roof, util

    def dense_roofline(h, sc, rows0, rows1, Wd, tag_wl):
        """FLOPs the dense kernels execute.  A fused conv + pool layer evaluates, per map position and pooled border class
        (first / interior / last per axis), the conv outputs of the pooling window with 2,3 / 3,3 / 3,2 live tap rows
        (columns): (5 + 6 + 5)^2 = 256 tap evaluations of 2*Cin*Cout FLOPs per position on the aligned grid (pan2).  On the
        stride-1 grids (ms2, pan3) the second sub-position of an interior cell is the first one of its neighbour and is
        evaluated once (dense_tc.cuh, SHARE): (5 + 3 + 5)^2 = 169.  No credit for tile padding (16/15, 8/7 on the shared
        axes), skipped taps or don't-care rows."""
        h.set_timing(True)
        h.get_dense_timing(reset=True)
        h.infer_scene(sc, rows0, rows1)
        stage = h.get_dense_timing()
        h.set_timing(False)
        band = max(1, min(args.band, rows1 - rows0))
        bands = [min(band, rows1 - b) for b in range(rows0, rows1, band)]
        pos = sum((nb + P - 1) * (Wd + P - 1) for nb in bands)                 # MS-resolution map positions of the bands
        def taps(nb, cells, step_, shared=True):
            """tap evaluations of one conv + pool layer over a band: per border class only the rows / columns some anchor uses"""
            tr = (5, 3, 5) if step_ == 2 and shared else (5, 6, 5)            # live tap rows (columns) evaluated per border class
            ext = (0, step_ * (cells - 3), 0)
            return sum(t * (nb + e) for t, e in zip(tr, ext)) * sum(t * (Wd + e) for t, e in zip(tr, ext))

        fl_s1 = sum(taps(nb, P // 2, 2) for nb in bands) * 2 * 64 * 128       # ms2, pan3: p/2 pooled cells at x + 2k
        fl_s1_unshared = sum(taps(nb, P // 2, 2, shared=False) for nb in bands) * 2 * 64 * 128
        fl_al = sum(taps(nb, P, 1) for nb in bands) * 2 * 32 * 64             # pan2: p pooled cells at x + k
        fl_fu = sum(3 * nb + P - 6 for nb in bands) * 3 * (Wd + P - 1) * 2 * 256 * 128
        kernels = {   # kernel -> (stage keys, FLOPs executed over all bands, launches per band)
            'conv_pool4_kernel<64,128> (ms2 + pan3: conv + stride-1 2x2 max, 9 pooled classes)': (['conv_ms2', 'conv_pan3'], 2 * fl_s1, 2),
            'conv_pool4_kernel<32,64> (pan2: conv + aligned 2x2 max)': (['conv_pan2'], fl_al, 1),
            'fuse_rowsum_kernel (1x1 fusion conv on 9 planes + row sums of the average pool; no credit for the 128/114 tile overlap)':
                (['conv_fuse'], fl_fu, 1),
        }
        ncu_key = {k: v for k, v in zip(kernels, ('conv_pool4_64_128', 'conv_pool4_32_64', 'fuse_rowsum'))}
        name, (keys, fl_k, per_band) = max(kernels.items(), key=lambda kv: sum(stage[k] for k in kv[1][0]))
        k_ms = sum(stage[k] for k in keys)
        achieved = fl_k / (k_ms / 1e3) / 1e12
        conv_fl = sum(v[1] for v in kernels.values())
        conv_ms = sum(stage[k] for k in ('conv_ms2', 'conv_pan2', 'conv_pan3', 'conv_fuse'))
        n_loc = (rows1 - rows0) * Wd
        roof = {'bound': 'tensor', 'kernel': name, 'achieved': achieved, 'unit': 'TFLOP/s', **fracs(achieved),
                'traffic': ncu_traffic('r02_dense_ncu_summary.json', ncu_key[name], tag_wl) if args.band == 512 else None,
                'avg_launch_ms': k_ms / (len(bands) * per_band), 'launches': len(bands) * per_band,
                'flops_per_launch': fl_k / (len(bands) * per_band),
                'stage_ms': {k: round(v, 3) for k, v in stage.items()},
                'map_positions': pos, 'flops_executed_per_pixel': conv_fl / n_loc,
                'stride1_layers': {'executed_TFLOPs': 2 * fl_s1 / ((stage['conv_ms2'] + stage['conv_pan3']) / 1e3) / 1e12,
                                   'unshared_equivalent_TFLOPs': 2 * fl_s1_unshared / ((stage['conv_ms2'] + stage['conv_pan3']) / 1e3) / 1e12,
                                   'note': 'ms2 + pan3: the cells of an interior border class share sub-positions with their neighbours (169 tap '
                                           'evaluations per position instead of 256, dense_tc.cuh SHARE).  executed = what the tensor pipe does '
                                           '(this is `achieved` when these layers dominate); unshared_equivalent = the same results at 256 per '
                                           'position / the same time, comparable with the lines before the sharing (profiles/r02_bench_before_sharing.json)'},
                'whole_step_executed_TFLOPs': conv_fl / (stage['total'] / 1e3) / 1e12,
                'per_patch_equivalent_tflops': h.flops_per_patch * n_loc / (stage['total'] / 1e3) / 1e12,
                'note': 'achieved = tensor-core FLOPs this kernel EXECUTES / its time (one isolated pass with per-stage events).  The stride-1 '
                        'layers execute a third fewer FLOPs for the same results since their interior-class cells share conv outputs with '
                        'their neighbours, so this utilisation figure fell (0.80 -> 0.6 of burst) while the time per scene fell by a fifth: '
                        'stride1_layers.unshared_equivalent_TFLOPs is the figure comparable with earlier lines; '
                        'per_patch_equivalent_tflops = the per-patch network FLOPs the same result would cost / whole-step time: it '
                        'exceeds the peak because the dense algorithm shares work between overlapping patches'}
        a = conv_fl / (conv_ms / 1e3) / 1e12
        util = {'achieved_TFLOPs': a, 'frac_of_sustained_peak': a / pk['bf16_tflops_sustained'], 'frac_of_burst_peak': a / pk['bf16_tflops'],
                'ncu_tensor_pipe_active_pct': 'profiles/r02_dense_ncu_summary.json'}
        return roof, util

    roofline, conv_util = (dense_roofline(handle, scene, r0, r1, W, wl_key) if args.mode == 'dense'
                           else patch_roofline(handle, scene, r0, r1, W))

    secondary = {'conv_tensor_util': conv_util, 'e2e_single_scene_latency_ms': e2e_latency_ms}
    if not args.no_secondary:
        # ---- K1 patch gather GB/s (BASELINE.json metric, second item) at the C5 sizes: p = 8 / 16 / 32, batch 8192
        wt = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def best_ms(fn, reps=5):
            for _ in range(3):
                fn()
            best = 1e9
            for _ in range(reps):
                flush.fill_(1)
                g0.record()
                fn()
                g1.record()
                torch.cuda.synchronize()
                best = min(best, g0.elapsed_time(g1))
            return best
        write_ceiling = (1 << 30) / best_ms(lambda: wt.zero_()) / 1e6
        del wt
        sweep = []
        for p_ in (8, 16, 32):
            sc_p = scene if p_ == P else dmf.Scene.from_raw(ms[:1000, :1000], pan[:4000, :4000], p_, dev)
            gidx = torch.randint(0, sc_p.H * sc_p.W, (8192,), device=dev)
            o_ms = torch.empty((8192, 4, p_, p_), device=dev)
            o_pan = torch.empty((8192, 1, 4 * p_, 4 * p_), device=dev)
            import ctypes as Cc
            call = lambda: dmf.check(dmf.lib.dmf_gather(sc_p._h, Cc.c_void_p(gidx.data_ptr()), 8192, Cc.c_void_p(o_ms.data_ptr()),
                                                        Cc.c_void_p(o_pan.data_ptr()), None, None,
                                                        Cc.c_void_p(torch.cuda.current_stream().cuda_stream)))
            g_ms = best_ms(call)
            gbs = 8192 * 20 * p_ * p_ * 4 / g_ms / 1e6
            ridx = gidx
            gidx = torch.arange(300 * sc_p.W + 17, 300 * sc_p.W + 17 + 8192, device=dev)       # loader order (test / colour loaders): consecutive pixels
            s_ms = best_ms(call)
            sgbs = 8192 * 20 * p_ * p_ * 4 / s_ms / 1e6
            gidx = ridx
            entry = {'patch_size': p_, 'batch': 8192, 'bytes_per_patch': 20 * p_ * p_ * 4,
                     'gather_GBs_written': sgbs, 'frac_of_hbm_copy_peak': sgbs / pk['hbm_gbs'],
                     'order': 'sequential pixels (the reference test / colour loaders: overlapping windows, source reads served by L2)',
                     'random_batch': {'gather_GBs_written': gbs, 'frac_of_hbm_copy_peak': gbs / pk['hbm_gbs'],
                                      'note': 'RandomSampler-style indices after an L2 flush: every window is first read from DRAM '
                                              '(compulsory reads ~ the bytes written), so the kernel is bound by read + write traffic'}}
            # conv tensor-pipe utilisation of the per-patch kernels at this patch size, batch 8192 (default-initialised weights)
            torch.manual_seed(0)
            net_p = Net({'Categories_Number': C, 'patch_size': p_, 'schedule': {'activate': 'Relu'}, 'b200': {'max_batch': 8192}}).to(dev).eval()
            hp = net_p.native()
            hp.set_dense(False)
            hp.forward_scene(sc_p, flat_idx=gidx, want_logits=False, want_pred=True)
            hp.set_timing(True)
            hp.forward_scene(sc_p, flat_idx=gidx, want_logits=False, want_pred=True)
            stp = hp.get_timing()
            hp.set_timing(False)
            cfl = 2 * conv(64, 128, 3, p_) + conv(32, 64, 3, 2 * p_) + conv(256, 128, 1, p_ // 2)
            cms = sum(stp[k] for k in ('conv_ms2', 'conv_pan2', 'conv_pan3', 'conv_fuse'))
            entry.update({'conv_TFLOPs': cfl * 8192 / (cms / 1e3) / 1e12, 'conv_frac_of_burst': cfl * 8192 / (cms / 1e3) / 1e12 / pk['bf16_tflops'],
                          'whole_net_TFLOPs': hp.flops_per_patch * 8192 / (stp['total'] / 1e3) / 1e12,
                          'patches_per_s': 8192 / (stp['total'] / 1e3), 'stage_ms': {k: round(v, 4) for k, v in stp.items()}})
            sweep.append(entry)
            hp.close()
            del net_p, o_ms, o_pan
            if sc_p is not scene:
                sc_p.close()
        p16 = sweep[1]
        secondary['patch_gather'] = {'GBs_written': p16['gather_GBs_written'],
                                     'frac_of_hbm_copy_peak': p16['frac_of_hbm_copy_peak'], 'batch': 8192, 'bytes_per_patch': p16['bytes_per_patch'],
                                     'order': p16['order'], 'random_batch': p16['random_batch'],
                                     'torch_fill_1GiB_GBs': write_ceiling,
                                     'note': 'write-dominated kernel (TMA load -> bulk store); denominator = the measured 6551 GB/s copy peak. A plain fill kernel '
                                             '(torch zero_ of 1 GiB, torch_fill_1GiB_GBs) writes slower than the bulk-store path, so it is not a ceiling'}
        secondary['c5_patch_size_sweep'] = {'workload': 'BASELINE.json configs[4]: p = 8 / 16 / 32 at batch 8192: K1 gather GB/s and conv tensor throughput of the per-patch kernels',
                                            'entries': sweep}
        if args.mode == 'dense':
            # the per-patch kernels on the same band, for comparison (one warm-up pass + one timed pass)
            handle.set_dense(False)
            handle.infer_scene(scene, r0, r1, pred_map=pred_map)
            pp0, pp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(1)
            pp0.record()
            handle.infer_scene(scene, r0, r1, pred_map=pred_map)
            pp1.record()
            torch.cuda.synchronize()
            pp_roof, pp_util = patch_roofline(handle, scene, r0, r1, W)
            secondary['per_patch_path'] = {'px_per_s_this_rank': n_local / pp0.elapsed_time(pp1) * 1e3, 'chunk_pixels': args.max_batch,
                                           'dominant_kernel': pp_roof['kernel'], 'achieved_TFLOPs': pp_roof['achieved'], 'frac_of_burst': pp_roof['frac_of_burst'],
                                           'frac_of_sustained': pp_roof['frac_of_sustained'], 'whole_net_tflops': pp_roof['whole_net_tflops'],
                                           'conv_tensor_util': pp_util, 'stage_ms': pp_roof['stage_ms']}
            handle.set_dense(True, args.band)
        if wl_key == 'c3' and world == 1:
            # BASELINE.json configs[1]: the C2 scene on one GPU
            ms2, pan2, lab2 = workload_scene('c2')
            net2 = fitted_product_net('c2', dev, max_batch=args.max_batch)
            h2 = net2.native()
            h2.set_dense(args.mode == 'dense', args.band)
            sc2 = dmf.Scene.from_raw(ms2, pan2, P, dev)
            sc2.set_labels(lab2)
            cm2 = torch.zeros((13, 13), dtype=torch.int64, device=dev)
            pm2 = torch.zeros((1000, 1000), dtype=torch.uint8, device=dev)
            for _ in range(3):
                h2.infer_scene(sc2, 0, 1000, pred_map=pm2, cm=cm2)
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for a, b in ev2:
                flush.fill_(1)
                cm2.zero_()
                a.record()
                h2.infer_scene(sc2, 0, 1000, pred_map=pm2, cm=cm2)
                b.record()
            torch.cuda.synchronize()
            ms2_step = sum(a.elapsed_time(b) for a, b in ev2) / len(ev2)
            with open(os.devnull, 'w') as null, _redirect(null):
                r2 = aa_oa(cm2.cpu().numpy().astype(np.float64))
            roof2, _ = dense_roofline(h2, sc2, 0, 1000, 1000, 'c2') if args.mode == 'dense' else patch_roofline(h2, sc2, 0, 1000, 1000)
            secondary['c2_single_gpu'] = {'workload': WORKLOADS['c2']['name'], 'px_per_s': 1e6 / (ms2_step / 1e3), 'ms_per_step': ms2_step, 'steps': 10,
                                          'OA_AA_Kappa': [float(r2[1]), float(r2[0]), float(r2[2])], 'stage_ms': roof2['stage_ms'],
                                          'dominant_kernel_TFLOPs': roof2['achieved'], 'frac_of_burst': roof2['frac_of_burst']}
            sc2.close()
            del net2, ms2, pan2
        if not args.no_train:
            secondary['train_step'] = train_step_metric(dev, world)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, kind, desc, _ = cpu_reference_run(wl_key, args.cpu_budget_s, scene=(ms, pan, label))
        cpu_baseline = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': desc}

    if rank == 0:
        details.update({'global_pixels': npix_total, 'row_band_rank0': [r0, r1], 'flops_per_pixel': handle.flops_per_patch,
                        'OA_AA_Kappa': [float(result[1]), float(result[0]), float(result[2])],
                        'classes_predicted': int((cm_host.sum(axis=1) > 0).sum()),
                        'e2e': 'dmf.ScenePipeline: 2-deep software pipeline over a stream of scenes (upload of scene i+1 on a copy stream while scene i is '
                               'classified); secondary.e2e_single_scene_latency_ms is the un-pipelined submit -> result time'})
        emit(({'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
               'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
               'dtype': 'fp16', 'data': 'synthetic', 'config': config, 'clocks': clk.summary(),
               'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_total, 'd2h_bytes_per_step': d2h_total},
               'gpu_launches': int(launches), 'roofline': roofline, 'secondary': secondary, 'cpu_baseline': cpu_baseline, 'details': details}))
    if world > 1:
        dist.destroy_process_group()


class _redirect:
    def __init__(self, f):
        self.f = f

    def __enter__(self):
        self.old = sys.stdout
        sys.stdout = self.f

    def __exit__(self, *a):
        sys.stdout = self.old


if __name__ == '__main__':
    main()
