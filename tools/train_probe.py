"""BASELINE.json configs[3]: IHS-input training step, batch 512 (default), p = 16, 12 classes, on one GPU.
Every step: K1 tri-gather of the batch from the device scene (PAN input = the IHS product's window) -> native forward
(train-mode BatchNorm) -> CrossEntropyLoss -> native backward -> FusedAdam, all inside libdmf_b200.
Prints one JSON line: ms per step (CUDA events), patches/s, algorithmic TFLOP/s (3 x forward FLOPs per patch)."""
import json, os, random, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np, torch, dmf
from oracle import dmf_oracle as orc
from image_convert.IHS import draw_offsets
from model.gmfnet import Net

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
p, C, Hs, Ws = 16, 12, 400, 400
ms, pan, label = orc.synthetic_scene(Hs, Ws, C - 1, seed=0, label_seed=1)
MSn = (ms - ms.min()) / (ms.max() - ms.min()); PANn = (pan - pan.min()) / (pan.max() - pan.min())
random.seed(7); offs = draw_offsets(Hs, Ws, 4, 4)
mspan = dmf.ihs_tran(torch.from_numpy(MSn).cuda(), torch.from_numpy(PANn).cuda(), torch.from_numpy(offs).cuda())
sc = dmf.Scene.from_raw(ms, pan, p)
sc.set_labels(label)
sc.set_mspan(np.pad(mspan.cpu().numpy(), ((0, 4 * p - 1), (0, 4 * p - 1)), mode='reflect'))
torch.manual_seed(0)
net = Net({'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}, 'b200': {'max_train_batch': B}}).cuda().train()
opt = dmf.FusedAdam(net.parameters(), lr=1e-3)
labelled = torch.from_numpy(np.flatnonzero(label.reshape(-1) != 0)).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
batches = [labelled[torch.randint(0, labelled.numel(), (B,), device='cuda', generator=g)] for _ in range(8)]
loss = torch.zeros((), device='cuda')
for i in range(5):
    loss = net.train_step_scene(sc, batches[i % 8], opt, use_mspan=True)
torch.cuda.synchronize()
l0 = dmf.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    loss = net.train_step_scene(sc, batches[i % 8], opt, use_mspan=True)
e1.record(); torch.cuda.synchronize()
ms_step = e0.elapsed_time(e1) / steps
flops = 3 * net.native().flops_per_patch * B
print(json.dumps({'config': 'C4 IHS-input training step, batch %d, p=16, 12 classes: tri-gather + native fwd/bwd + FusedAdam' % B,
                  'ms_per_step': round(ms_step, 4), 'patches_per_s': round(B / ms_step * 1e3), 'algorithmic_TFLOPs': round(flops / ms_step / 1e9, 1),
                  'kernels_per_step': (dmf.launch_count() - l0) / steps, 'loss': float(loss)}))
