"""Quick device-side timing of the fused whole-scene path with per-stage breakdown (dev tool)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np, torch
import dmf
from oracle import dmf_oracle as orc
from oracle.gmfnet_ref import Net as RefNet

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
rows = int(sys.argv[4]) if len(sys.argv) > 4 else 100
p, C = 16, 13
ms, pan, label = orc.synthetic_scene(H, W, C - 1, seed=0, label_seed=1)
torch.manual_seed(3407)
net = RefNet({'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}}).eval()
h = dmf.NetHandle(p, C, max_batch=NB)
h.load_state_dict(net.state_dict())
t0 = time.perf_counter(); sc = dmf.Scene.from_raw(ms, pan, p); sc.set_labels(label); torch.cuda.synchronize()
print('scene create (H2D + normalise + pad): %.1f ms' % ((time.perf_counter() - t0) * 1e3))
h.infer_scene(sc, 0, 8); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pm, cm = h.infer_scene(sc, 0, rows); e1.record(); torch.cuda.synchronize()
ms_t = e0.elapsed_time(e1); npx = rows * W
fl = h.flops_per_patch
print('rows=%d px=%d: %.2f ms -> %.3f Mpx/s, %.1f TFLOP/s (%.1f%% of 1406.8 sustained)' % (rows, npx, ms_t, npx / ms_t / 1e3, npx * fl / ms_t / 1e9, npx * fl / ms_t / 1e9 / 1406.8 * 100))
h.set_timing(True); h.infer_scene(sc, 0, rows); t = h.get_timing(); h.set_timing(False)
print({k: round(v, 2) for k, v in t.items()})
# gather GB/s
idx = torch.randint(0, H * W, (8192,), device='cuda')
for _ in range(3): sc.gather(idx, want_target=False)
e0.record()
for _ in range(10): sc.gather(idx, want_target=False)
e1.record(); torch.cuda.synchronize()
g_ms = e0.elapsed_time(e1) / 10
print('gather 8192 patches p=16: %.3f ms -> %.0f GB/s written' % (g_ms, 8192 * 20480 / g_ms / 1e6))
