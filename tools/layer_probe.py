"""Time one tcgen05 layer in isolation, with diagnostic modes that remove the TMA loads and/or the
epilogue, to see which stage bounds the tile loop (dev tool)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import torch, dmf
from oracle.gmfnet_ref import Net as RefNet
p, C, N = 16, 13, 4096
torch.manual_seed(0)
h = dmf.NetHandle(p, C, max_batch=N)
h.load_state_dict(RefNet({'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}}).state_dict())
L = {0: ('ms2', 64, 16, (N, 32, 8, 8, 8), 37.75e6), 1: ('pan2', 32, 32, (N, 8, 16, 16, 8), 37.75e6),
     2: ('pan3', 64, 16, (N, 32, 8, 8, 8), 37.75e6), 3: ('fuse', 256, 8, (N, 16, 8, 8, 8), 4.19e6)}
import subprocess
print(subprocess.run(['nvidia-smi','--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu','--format=csv,noheader'],capture_output=True,text=True).stdout.strip())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for layer, (name, cin, S, oshape, fl) in L.items():
    x = torch.rand((N, cin // 8, S, S, 8), device='cuda').to(torch.bfloat16)
    for impl, tag in ((0, 'full'), (2, 'no-TMA'), (3, 'no-epilogue'), (4, "MMA only")):
        out = torch.zeros(oshape, dtype=torch.bfloat16, device='cuda')
        for _ in range(3): h.debug_layer(layer, impl, x, oshape, out)
        best = 1e9
        for _ in range(5):
            e0.record()
            for _ in range(4): h.debug_layer(layer, impl, x, oshape, out)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4)
        ms = best
        print('%-5s %-12s %.3f ms  %.0f TFLOP/s' % (name, tag, ms, fl * N / ms / 1e9))
