"""HBM-bound kernels of the data path at the BASELINE.json sizes: CUDA-event GB/s (one JSON line each), or --once for ncu.

    python tools/datapath_probe.py                  # timed, L2 flushed between repetitions
    ncu --set full --clock-control none -k regex:'gather_tma|ihs_|argmax_confusion|paint|confusion_at|pan2ms' -c 40 \
        -o gpurun_out/r02_datapath python tools/datapath_probe.py --once

K1 gather at p = 8 / 16 / 32 x batch 8192, dual and tri (train/dataset.py:168-185, 259-279); K2 IHS_tran / pan2ms / the fused
IHS -> scene pass at C3 size (image_convert/IHS.py:6-54); K4 argmax + confusion on 1e6 x 13 logits (solver/mainsolver.py:139-141);
confusion_at and K5 paint at C3 size (solver/mainsolver.py:186-189).  Algorithmic bytes as in SURVEY.md 8(d) / DESIGN.md.
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np
import torch
import dmf
from oracle import dmf_oracle as orc

once = '--once' in sys.argv
dev = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
try:
    peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
except Exception:
    peaks = {'hbm_gbs': 6650.0}


def timed(fn, reps=7):
    if once:
        fn()
        torch.cuda.synchronize()
        return None
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(reps):
        flush.fill_(1)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def report(kernel, what, nbytes, ms, **extra):
    if ms is None:
        return
    gbs = nbytes / ms / 1e6
    print(json.dumps(dict(kernel=kernel, case=what, algorithmic_bytes=int(nbytes), ms=round(ms, 4), GBs=round(gbs, 1),
                          frac_of_copy_peak=round(gbs / peaks['hbm_gbs'], 3), **extra)), flush=True)


# the write-only ceiling of this GPU (the gather is a write-bound kernel; the copy peak counts read + write)
buf = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
ms_ = timed(lambda: buf.zero_())
write_peak = None
if ms_ is not None:
    write_peak = (1 << 30) / ms_ / 1e6
    report('memset (torch zero_)', '1 GiB pure write: the write-only ceiling', 1 << 30, ms_)
del buf

H, W = 1000, 1000
ms, pan, label = orc.synthetic_scene(H, W, 12, seed=0)
for p in (8, 16, 32):
    sc = dmf.Scene.from_raw(ms, pan, p, dev)
    sc.set_mspan(np.random.default_rng(1).random((sc.H4p, sc.W4p), dtype=np.float32))
    B = 8192
    idx = torch.randint(0, H * W, (B,), device=dev)
    o_ms = torch.empty((B, 4, p, p), device=dev)
    o_pan = torch.empty((B, 1, 4 * p, 4 * p), device=dev)
    o_msp = torch.empty_like(o_pan)
    import ctypes as C
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    seq = torch.arange(300 * W + 17, 300 * W + 17 + B, device=dev)          # loader order of the test / colour loaders: consecutive pixels
    for order, ii in (('random (RandomSampler batch; the windows are read from a cold L2: compulsory reads on top of the writes)', idx),
                      ('sequential (test / colour loader order: overlapping windows, reads served by L2)', seq)):
        for tri in (False, True):
            fn = lambda: dmf.check(dmf.lib.dmf_gather(sc._h, ptr(ii), B, ptr(o_ms), ptr(o_pan), ptr(o_msp if tri else None), ptr(None), st()))
            t = timed(fn)
            nbytes = B * (4 * p * p + 16 * p * p * (2 if tri else 1)) * 4
            extra = {'frac_of_write_ceiling': round(nbytes / t / 1e6 / write_peak, 3)} if t else {}
            report('gather_tma_kernel', 'p=%d batch %d %s, %s' % (p, B, 'tri' if tri else 'dual', order), nbytes, t, **extra)
    sc.close()

# K2 at C3 size
H3, W3 = 2001, 2101
rng = np.random.default_rng(3)
MS = torch.rand((H3, W3, 4), dtype=torch.float64, device=dev)
PAN = torch.rand((4 * H3, 4 * W3), dtype=torch.float64, device=dev)
offs = torch.randint(0, 4, (4, H3, W3, 2), dtype=torch.int8, device=dev)
out = torch.empty_like(PAN)
t = timed(lambda: dmf.check(dmf.lib.dmf_ihs_tran(ptr(MS), ptr(PAN), ptr(offs), ptr(out), H3, W3, st())))
report('ihs_tran_kernel', 'C3 2001x2101 fp64', H3 * W3 * 296, t)
pan16 = torch.randint(0, 2048, (4 * H3, 4 * W3), dtype=torch.int16, device=dev)
o4 = torch.empty((H3, W3, 4), dtype=torch.float64, device=dev)
t = timed(lambda: dmf.check(dmf.lib.dmf_pan2ms(ptr(pan16), dmf.U16, 4 * H3, 4 * W3, ptr(o4), st())))
report('pan2ms_kernel<u16>', 'C3 8004x8404 u16 -> f64 [H,W,4]', 16 * H3 * W3 * 2 + 4 * H3 * W3 * 8, t)
del MS, PAN, out, o4
ms16 = torch.randint(0, 2048, (H3, W3, 4), dtype=torch.int16, device=dev)
sc3 = dmf.Scene.from_raw(ms16, pan16, 16, dev)
t = timed(lambda: sc3.set_mspan_ihs(ms16, pan16, offs))
report('ihs_scene_kernel<u16,u16> (+ 2 min/max passes)', 'C3 raw u16 -> padded f32 MSPAN in the scene',
       H3 * W3 * (8 + 8 + 32) + sc3.H4p * sc3.W4p * 4, t)
sc3.close()
del ms16, pan16, offs

# K4 / K5
N, Cn = 1_000_000, 13
lg = torch.randn((N, Cn), device=dev)
tg = torch.randint(0, Cn, (N,), device=dev).float()
cm = torch.zeros((Cn, Cn), dtype=torch.int64, device=dev)
pred = torch.empty((N,), dtype=torch.int64, device=dev)
t = timed(lambda: dmf.check(dmf.lib.dmf_argmax_confusion(ptr(lg), ptr(tg), dmf.F32, N, Cn, ptr(None), ptr(cm), st())))
report('argmax_confusion_kernel', '1e6 x 13 fp32 logits, f32 targets, no pred output', N * (4 * Cn + 4), t)
t = timed(lambda: dmf.check(dmf.lib.dmf_argmax_confusion(ptr(lg), ptr(tg), dmf.F32, N, Cn, ptr(pred), ptr(cm), st())))
report('argmax_confusion_kernel', '1e6 x 13 fp32 logits, f32 targets, int64 pred output', N * (4 * Cn + 4 + 8), t)
npx = H3 * W3
pm = torch.randint(0, 12, (npx,), dtype=torch.uint8, device=dev)
lb = torch.randint(0, 12, (npx,), dtype=torch.uint8, device=dev)
cm12 = torch.zeros((12, 12), dtype=torch.int64, device=dev)
t = timed(lambda: dmf.check(dmf.lib.dmf_confusion_at(ptr(pm), ptr(lb), ptr(None), npx, 12, ptr(cm12), st())))
report('confusion_at_kernel', 'C3 whole scene, u8 maps', 2 * npx, t)
pal = np.ascontiguousarray(np.asarray([[(37 * i) % 256, (91 * i) % 256, (53 * i) % 256] for i in range(12)], dtype=np.uint8))
rgb = torch.empty((npx, 3), dtype=torch.uint8, device=dev)
t = timed(lambda: dmf.check(dmf.lib.dmf_paint_labels(ptr(pm), npx, pal.ctypes.data_as(C.c_void_p), 12, ptr(rgb), st())))
report('paint_kernel', 'C3 whole scene u8 -> RGB', 4 * npx, t)
