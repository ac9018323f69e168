"""Where does the end-to-end (host rasters -> metrics) time go? (dev tool)"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np, torch, dmf
from oracle import dmf_oracle as orc
from model.gmfnet import Net
H = W = 1000; P = 16; C = 13
ms, pan, label = orc.synthetic_scene(H, W, C - 1, seed=0, label_seed=1)
torch.manual_seed(3407)
net = Net({'Categories_Number': C, 'patch_size': P, 'schedule': {'activate': 'Relu'}}).to('cuda:0').eval()
handle = net.native()
ms_pin = torch.from_numpy(ms.view(np.int16)).pin_memory(); pan_pin = torch.from_numpy(pan.view(np.int16)).pin_memory()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); sc = dmf.Scene.from_raw(ms_pin, pan_pin, P, 'cuda:0'); t1 = T()
    sc.set_labels(label); t2 = T()
    pm, m = handle.infer_scene(sc); t3 = T()
    a = pm.cpu(); b = m.cpu(); t4 = T()
    sc.close(); t5 = T()
    print('rep %d: from_raw %.1f ms, set_labels %.1f, infer %.1f, d2h %.1f, close %.1f' % (rep, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3))
