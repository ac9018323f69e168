"""Produce tests/golden/fitted_nets.npz: BatchNorm tensors and head weights of the "fitted" GMFNets (oracle/fitted_net.py).

    python tests/golden/make_fitted_nets.py            # authoring container, CPU, ~2 minutes

Per workload (c1, c2, c3, smoke): structured synthetic scene -> padded float64 rasters (oracle.data_padding) -> patches of
sampled labelled pixels (oracle.gather_dual) -> seed-3407 default-initialised oracle Net with seeded random BN gamma / beta ->
BN running statistics calibrated on one batch -> fc1 / fc2 fitted by full-batch Adam on the pooled features (fp32, CPU).
Only fp32 tensors are stored (state_dict names prefixed with the workload tag); the convolutions are rebuilt from the seed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import dmf_oracle as orc            # noqa: E402
from oracle import fitted_net as fn             # noqa: E402


def patches(MS, PAN, W, idx, p):
    a, b = orc.gather_dual(MS, PAN, idx // W, idx % W, p)
    return torch.from_numpy(a), torch.from_numpy(b)


def fit(tag):
    H, W, ncls, p = fn.WORKLOADS[tag]
    C = ncls + 1
    ms, pan, label = fn.scene(tag)
    MS, PAN = orc.data_padding(ms, p), orc.data_padding(pan, p)
    lab = label.reshape(-1)
    labelled = np.flatnonzero(lab != 0)
    rng = np.random.default_rng(7)
    n_fit = min(6000, labelled.size // 2)
    pick = rng.choice(labelled, size=min(labelled.size, n_fit + 3000), replace=False)
    tr, te = pick[:n_fit], pick[n_fit:]
    net = fn.base_net(tag)
    g = torch.Generator().manual_seed(11)
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.weight.data.copy_(torch.rand(m.num_features, generator=g) * 0.8 + 0.6)
        m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.2)
        m.momentum = 1.0                          # running statistics := the calibration batch's
    net.train()
    with torch.no_grad():
        net.features(*patches(MS, PAN, W, tr[:1024], p))
    for m in bns:
        m.momentum = 0.1
    net.eval()

    def feats(idx):
        out = []
        with torch.no_grad():
            for i in range(0, idx.size, 500):
                out.append(net.features(*patches(MS, PAN, W, idx[i:i + 500], p)))
        return torch.cat(out)
    Ftr, Fte = feats(tr), feats(te)
    ytr, yte = torch.from_numpy(lab[tr].astype(np.int64)), torch.from_numpy(lab[te].astype(np.int64))
    torch.manual_seed(5)
    head = torch.nn.Sequential(torch.nn.Linear(128, 64), torch.nn.ReLU(), torch.nn.Linear(64, C))
    opt = torch.optim.AdamW(head.parameters(), lr=3e-3, weight_decay=1e-2)        # an ordinarily regularised head, not a razor-sharp one
    for _ in range(1500):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(head(Ftr), ytr)
        loss.backward()
        opt.step()
    with torch.no_grad():
        pred = head(Fte).argmax(1).numpy()
    M = orc.confusion(pred, yte.numpy(), C)
    aa, oa, k, _ = orc.aa_oa(M)
    print('%s: fit loss %.4f, held-out %d px: OA %.4f AA %.4f Kappa %.4f, predicted classes %d' %
          (tag, float(loss), te.size, oa, aa, k, len(np.unique(pred))))
    net.fc1.load_state_dict(head[0].state_dict())
    net.fc2.load_state_dict(head[2].state_dict())
    out = {}
    for name, t in net.state_dict().items():
        if ('.1.' in name or name.startswith('fc')) and t.dtype == torch.float32:     # BatchNorm (index 1 of each block) + head
            out[tag + '/' + name] = t.numpy().copy()
    return out


if __name__ == '__main__':
    # `python make_fitted_nets.py p8 p32` fits only the named workloads and keeps the stored tensors of the others byte for byte
    # (a refit on another thread count can differ in the last bits, and tests/golden/solver_c1_fitted.npz was produced with the stored c1 net)
    only = sys.argv[1:]
    blob = {}
    if only:
        with np.load(fn.GOLDEN, allow_pickle=False) as z:
            blob = {k: z[k].copy() for k in z.files if k.split('/')[0] not in only}
    for tag in (only or fn.WORKLOADS):
        blob.update(fit(tag))
    np.savez_compressed(fn.GOLDEN, **blob)
    print(fn.GOLDEN, os.path.getsize(fn.GOLDEN), 'bytes,', len(blob), 'tensors')
