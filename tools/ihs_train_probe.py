"""BASELINE.json configs[3] (IHS-input training path) as it stands in round 1, plus K2 timing at scene size.
  K2 IHS_tran / pan2ms on the device (fp64, bit-exact)           -> GB/s against the HBM roofline (296 B per MS pixel)
  one training step, batch 512: K1 tri-gather (native) + forward/backward/Adam through torch autograd on the same
  Parameters (native backward kernels are the next scope row)."""
import json, os, random, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np, torch, dmf
from oracle import dmf_oracle as orc
from image_convert.IHS import draw_offsets
from model.gmfnet import Net
H, W, p, C = 2001, 2101, 16, 12
ms, pan, label = orc.synthetic_scene(H, W, C - 1, seed=0, label_seed=1)
MSn = (ms - ms.min()) / (ms.max() - ms.min()); PANn = (pan - pan.min()) / (pan.max() - pan.min())
random.seed(7); t0 = time.perf_counter(); offs = draw_offsets(H, W, 4, 4); t_off = time.perf_counter() - t0
d_ms, d_pan, d_off = torch.from_numpy(MSn).cuda(), torch.from_numpy(PANn).cuda(), torch.from_numpy(offs).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(4):
    e0.record(); out = dmf.ihs_tran(d_ms, d_pan, d_off); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print(json.dumps({'kernel': 'ihs_tran', 'scene': 'C3 2001x2101', 'ms': round(best, 3), 'GBs': round(H * W * 296 / best / 1e6, 1), 'host_offsets_s': round(t_off, 2),
                  'max_abs_minus_pan': float((out - d_pan).abs().max())}))
d_pan16 = torch.from_numpy(pan.view(np.int16)).cuda()
best = 1e9
for _ in range(4):
    e0.record(); o2 = dmf.pan2ms(pan); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print(json.dumps({'kernel': 'pan2ms (includes H2D of the uint16 raster)', 'ms': round(best, 3)}))
# training step, batch 512
Hs = Ws = 400
sc = dmf.Scene.from_raw(ms[:Hs, :Ws].copy(), pan[:4 * Hs, :4 * Ws].copy(), p)
sc.set_labels(label[:Hs, :Ws].copy())
sc.set_mspan(np.pad(out[:4 * Hs, :4 * Ws].cpu().numpy(), ((0, 4 * p - 1), (0, 4 * p - 1)), mode='reflect'))
torch.manual_seed(0)
net = Net({'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}}).cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3); loss_fn = torch.nn.CrossEntropyLoss()
idx = torch.randint(0, Hs * Ws, (512,))
def step():
    a, b, m, t = sc.gather(idx, tri=True)
    opt.zero_grad(); loss = loss_fn(net(a, m), t.long()); loss.backward(); opt.step(); return loss
for _ in range(5): step()
torch.cuda.synchronize(); e0.record()
for _ in range(20): l = step()
e1.record(); torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / 20
print(json.dumps({'config': 'C4 training step, batch 512, tri gather native + torch autograd fwd/bwd/Adam', 'ms_per_step': round(ms_step, 3), 'patches_per_s': round(512 / ms_step * 1e3), 'loss': float(l)}))
