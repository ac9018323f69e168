"""BASELINE.json configs[4]: patch-size sweep (MS 8/16/32, PAN 32/64/128) at batch 8192 —
K1 gather GB/s written (fp32 patches) and tensor-pipe rate of the tcgen05 layers."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np, torch, dmf
from oracle import dmf_oracle as orc
from oracle.gmfnet_ref import Net as RefNet

B = 8192
peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(REPO, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops_sustained': 1400.0}
ms, pan, label = orc.synthetic_scene(1000, 1000, 12, seed=0, label_seed=1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rows = []
for p in (8, 16, 32):
    sc = dmf.Scene.from_raw(ms, pan, p)
    idx = torch.randint(0, 1000 * 1000, (B,), device='cuda')
    for _ in range(3): sc.gather(idx, want_target=False)
    best = 1e9
    for _ in range(5):
        e0.record(); sc.gather(idx, want_target=False); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    bytes_per_patch = 4 * p * p * 4 + 16 * p * p * 4
    gbs = B * bytes_per_patch / best / 1e6
    torch.manual_seed(0)
    h = dmf.NetHandle(p, 13, max_batch=B)
    h.load_state_dict(RefNet({'Categories_Number': 13, 'patch_size': p, 'schedule': {'activate': 'Relu'}}).state_dict())
    h.set_timing(True)
    for _ in range(2):
        h.set_timing(True); h.forward_scene(sc, flat_idx=idx, want_logits=False, want_pred=True)
    t = h.get_timing(); h.set_timing(False)
    conv = lambda cin, cout, k, s: 2 * cin * cout * k * k * s * s
    fl = {'conv_ms2': conv(64, 128, 3, p), 'conv_pan2': conv(32, 64, 3, 2 * p), 'conv_pan3': conv(64, 128, 3, p), 'conv_fuse': conv(256, 128, 1, p // 2)}
    row = {'p': p, 'gather_ms': round(best, 4), 'gather_GBs_written': round(gbs, 1), 'gather_frac_of_hbm_peak': round(gbs / peaks['hbm_gbs'], 3),
           'forward_ms_8192': round(t['total'], 3), 'px_per_s': round(B / t['total'] * 1e3)}
    for k, f in fl.items():
        row[k + '_TFLOPs'] = round(f * B / t[k] / 1e9, 1)
    row['net_TFLOPs'] = round(h.flops_per_patch * B / t['total'] / 1e9, 1)
    rows.append(row)
    print(json.dumps(row))
    sc.close(); h.close()
