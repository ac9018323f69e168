import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np, torch, dmf
from oracle import dmf_oracle as orc
ms, pan, _ = orc.synthetic_scene(1000, 1000, 12, seed=0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for p in (8, 16, 32):
    sc = dmf.Scene.from_raw(ms, pan, p)
    for B in (8192, 65536):
        if p == 32 and B > 8192: B = 32768
        idx = torch.randint(0, 10**6, (B,), device='cuda')
        for _ in range(3): sc.gather(idx, want_target=False)
        best = 1e9
        for _ in range(5):
            e0.record(); sc.gather(idx, want_target=False); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        print('p=%d B=%d: %.3f ms -> %.0f GB/s written (includes two torch.empty)' % (p, B, best, B * 20 * p * p * 4 / best / 1e6))
    sc.close()
