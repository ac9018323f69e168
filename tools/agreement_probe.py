"""Where does the bf16-path logit error come from?  Product (dense + per-patch) vs the fp32 oracle with (i) the fp32 checkpoint
and (ii) the same checkpoint with conv weights rounded to bf16 (what the kernels actually multiply with)."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np, torch
import dmf
from model.gmfnet import Net
from oracle import dmf_oracle as orc, fitted_net
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda:0'
for tag in sys.argv[1:] or ['c1', 'c2']:
    H, W, ncls, p = fitted_net.WORKLOADS[tag]
    ms, pan, label = fitted_net.scene(tag)
    sd = fitted_net.fitted_state(tag)
    net = Net(dict(fitted_net.cfg_for(tag), b200={})); net.load_state_dict(sd); net = net.to(dev).eval()
    scene = dmf.Scene.from_raw(ms, pan, p, dev); scene.set_labels(label)
    rng = np.random.default_rng(123)
    idx = np.arange(H * W) if H * W <= 20000 else np.sort(rng.choice(H * W, size=24000, replace=False))
    h = net.native()
    h.set_dense(True)
    _, _, lg_dense = h.infer_scene(scene, want_logits=True)
    lg_dense = lg_dense.cpu().numpy()[idx]
    lg_patch, _ = h.forward_scene(scene, flat_idx=torch.from_numpy(idx), want_logits=True)
    lg_patch = lg_patch.cpu().numpy()
    refs = {}
    for name in ('fp32 weights', 'fp16-rounded conv weights'):
        ref = fitted_net.base_net(tag)
        s2 = {k: v.clone() for k, v in sd.items()}
        if name.startswith('fp16'):
            for k in s2:
                if k.endswith('.0.weight') and k.split('.')[0] in ('ms2', 'pan2', 'pan3', 'fuse'):      # the stems run hi/lo-split (fp32-grade)
                    s2[k] = s2[k].to(torch.float16).float()
        ref.load_state_dict(s2); ref = ref.to(dev).eval()
        out = []
        with torch.no_grad():
            for i in range(0, idx.size, 2000):
                a, b, _ = scene.gather(torch.from_numpy(idx[i:i + 2000]), want_target=False)
                out.append(ref(a, b).cpu().numpy())
        refs[name] = np.concatenate(out)
    for name, want in refs.items():
        for pn, got in (('dense', lg_dense), ('per-patch', lg_patch)):
            err = np.abs(got - want)
            scale = np.abs(want).max(axis=1)
            agree = (got.argmax(1) == want.argmax(1)).mean()
            print(json.dumps({'workload': tag, 'oracle': name, 'path': pn, 'px': int(idx.size), 'max_abs_err': float(err.max()), 'mean_abs_err': float(err.mean()),
                              'max_err_over_row_scale': float((err.max(axis=1) / scale).max()), 'median_err_over_row_scale': float(np.median(err.max(axis=1) / scale)),
                              'argmax_agreement': float(agree), 'differ': int((got.argmax(1) != want.argmax(1)).sum())}), flush=True)
    a, b = refs['fp32 weights'], refs['fp16-rounded conv weights']
    print(json.dumps({'workload': tag, 'oracle fp32 vs oracle bf16-weights': float((a.argmax(1) == b.argmax(1)).mean()), 'max_abs': float(np.abs(a - b).max())}), flush=True)
