"""ORACLE / CPU BASELINE (test infrastructure): the reference's own evaluation path restated as it
runs there — a torch Dataset whose __getitem__ slices one patch at a time from the padded float64
rasters (train/dataset.py:168-185), DataLoader(batch_size=300, shuffle=False, num_workers=0)
(solver/basesolver.py:96-104), the fp32 Net on the CPU, ``output.data.max(1, keepdim=True)[1]`` and
the per-sample ``M[pred][target] += 1`` loop (solver/mainsolver.py:139-141, train/test.py:58-60),
then aa_oa (indicators/kappa.py:69-84).

The reference itself is Python and cannot travel to the GPU box, so this port (kind = "port") is
what bench.py times as ``cpu_baseline`` and as the ``--impl reference`` arm.  Its pieces are pinned
bit-for-bit against the reference in tests/test_oracle_golden.py.
"""
import time

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Subset

from . import dmf_oracle as orc
from .gmfnet_ref import Net


class RefDatasetDual(Dataset):
    def __init__(self, MS, PAN, xyl, p):
        self.MS, self.PAN = MS, PAN
        self.x, self.y, self.Label = xyl[0], xyl[1], xyl[2]
        self.p = p

    def __getitem__(self, index):
        p = self.p
        x, y = int(self.x[index]), int(self.y[index])
        ms = self.MS[x:x + p, y:y + p, :]
        pan = self.PAN[4 * x:4 * x + 4 * p, 4 * y:4 * y + 4 * p]
        label = torch.Tensor(self.Label[index]).squeeze()
        ms = torch.from_numpy(ms.transpose((2, 0, 1))).type(torch.FloatTensor)
        pan = torch.from_numpy(np.expand_dims(pan, axis=0)).type(torch.FloatTensor)
        return ms, pan, label, x, y

    def __len__(self):
        return len(self.x)


class RefPipeline:
    """Scene preparation + per-pixel classification exactly as the reference's Solver would do it."""

    def __init__(self, ms, pan, label, p, num_classes, net=None, seed=3407):
        self.H, self.W = label.shape
        self.p, self.C = p, num_classes
        t0 = time.perf_counter()
        self.MS = orc.data_padding(ms, p)
        self.PAN = orc.data_padding(pan, p)
        self.xyl, self.matrix_ = orc.split_data_old(label, [self.H, self.W, 4])
        self.prep_s = time.perf_counter() - t0
        self.dataset = RefDatasetDual(self.MS, self.PAN, self.xyl, p)
        if net is None:
            torch.manual_seed(seed)
            net = Net({'Categories_Number': num_classes, 'patch_size': p, 'schedule': {'activate': 'Relu'}})
        self.net = net.eval()

    def classify(self, indices, batch_size=300, budget_s=None):
        """Runs the reference loop over `indices` (stops early once `budget_s` seconds are spent).
        Returns (confusion float64 [C,C], label_map float64 [H,W], pixels done, seconds)."""
        M = np.zeros([self.C, self.C])
        label_map = np.zeros([self.H, self.W])
        loader = DataLoader(Subset(self.dataset, list(indices)), batch_size=batch_size, shuffle=False, num_workers=0)
        done = 0
        t0 = time.perf_counter()
        with torch.no_grad():
            for data1, data2, target, x, y in loader:
                output = self.net(data1, data2)
                pred = output.data.max(1, keepdim=True)[1]
                for i in range(len(target)):
                    M[int(pred[i].item())][int(target[i].item())] += 1
                    label_map[int(x[i])][int(y[i])] = int(pred[i])
                done += len(target)
                if budget_s is not None and time.perf_counter() - t0 > budget_s:
                    break
        return M, label_map, done, time.perf_counter() - t0
