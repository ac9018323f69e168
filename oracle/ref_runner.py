"""ORACLE / CPU BASELINE (test infrastructure): run the UNMODIFIED reference on the host cores.

`__graft_entry__.build()` copies the reference's Python files from /root/reference into the git-ignored oracle/_ref/ (it is
pure Python: "building" it is a file copy; the directory travels to the GPU box with the repo snapshot like a built .so and
never enters the history).  This module imports that copy exactly as tests/golden/make_golden.py imports /root/reference:
empty stand-ins for the uninstalled libtiff / h5py / openpyxl / matplotlib (none is on the hot path),
`solver.basesolver.read_tif` pointed at in-memory rasters, and a `model.gmfnet` module (the reference never shipped one,
solver/mainsolver.py:30-34) holding the fp32 oracle Net.  What is timed is then the reference's own objects:
BaseSolver.__init__ (to_tensor + data_padding + split_data_old, solver/basesolver.py:25-58), dataloader() (:63-105), the
DataLoader over dataset_dual.__getitem__ (train/dataset.py:168-185), the Net on the CPU, and the per-sample confusion /
label-map loops of solver/mainsolver.py:139-141, 171-173 (clean copy: train/test.py:58-60), then indicators.kappa.aa_oa.
"""
import contextlib
import io
import os
import sys
import tempfile
import time
import types

import numpy as np
import torch

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    return os.path.exists(os.path.join(REF_DIR, 'solver', 'mainsolver.py'))


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


class RefRun:
    """The reference's Solver on an in-memory scene with `net_factory(cfg) -> nn.Module` as model.gmfnet.Net."""

    def __init__(self, ms, pan, label, p, n_classes, net_factory):
        assert available(), 'oracle/_ref is missing (run __graft_entry__.build() where /root/reference exists)'
        for name in ['libtiff', 'h5py', 'openpyxl', 'matplotlib', 'matplotlib.pyplot']:
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules['libtiff'].TIFF = object
        sys.modules['openpyxl'].Workbook = object
        sys.modules['openpyxl'].load_workbook = lambda *a, **k: None
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
        # the reference's top-level package names (solver, train, utils, ...) collide with the product mirror's: give the
        # reference its own import context for the duration of the imports
        saved_path = list(sys.path)
        saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                      if k.split('.')[0] in ('solver', 'train', 'utils', 'function', 'indicators', 'image_convert', 'model')}
        # (some reference packages have no __init__.py: as namespace packages they would lose against the mirror's regular ones)
        sys.path[:] = [REF_DIR] + [q for q in sys.path if 'dual-modal-fusion_b200' not in q]
        try:
            import indicators.kappa as rk
            import solver.basesolver as rbs
            import solver.mainsolver as rms
            self.rk = rk
            mod = types.ModuleType('model.gmfnet')
            mod.Net = lambda args: net_factory(args)
            sys.modules['model'] = types.ModuleType('model')
            sys.modules['model.gmfnet'] = mod
            H, W = label.shape
            tmp = tempfile.mkdtemp() + '/'
            np.save(tmp + 'label.npy', label)
            colors = [[(37 * i) % 256, (91 * i) % 256, (53 * i) % 256] for i in range(n_classes + 1)]
            cfg = {'patch_size': p, 'data_city': 'synthetic', 'DATA_DICT': {'synthetic': {'size': [H, W, 4], 'color': colors}},
                   'task': 'classification', 'time': 1, 'index': 0, 'epoch': 1, 'device': 'cpu', 'gpu_mode': False, 'data_new': 0,
                   'data_address': tmp, 'use_h5': False, 'nohup': 0, 'model_name': 'gmfnet', 'batchsize': 256, 'test_batchsize': 300,
                   'color_batchsize': 300, 'train_rate': 0.02, 'verify_rate': 0.02, 'Categories_Number': n_classes + 1,
                   'schedule': {'loss': 'Criterion', 'optimizer': 'ADAM', 'if_scheduler': 0, 'scheduler': 'ExponentialLR',
                                'activate': 'Relu', 'lr': 1e-3, 'base_lr': 5e-4},
                   'train': {'index': 0, 'pretrained': 0, 'save_best': True}, 'test': {'index': 1, 'save_matrix': 1},
                   'color': {'index': 1, 'supervised': 1, 'unsupervised': 1}}
            rbs.read_tif = lambda c, mode: ms if mode == 'ms' else pan
            t0 = time.perf_counter()
            self.solver = _quiet(rms.Solver, cfg)
            _quiet(self.solver.dataloader)
            _quiet(self.solver.init_model)
            self.prep_s = time.perf_counter() - t0
        finally:
            sys.path[:] = saved_path
            self.ref_modules = {k: sys.modules.pop(k) for k in list(sys.modules)
                                if k.split('.')[0] in ('solver', 'train', 'utils', 'function', 'indicators', 'image_convert', 'model')}
            sys.modules.update(saved_mods)
        self.H, self.W, self.C = H, W, n_classes + 1
        self.net = self.solver.model.eval()
        self._iter = None

    def classify(self, budget_s):
        """The reference's colour loop (labelled pixels first, then unlabelled: color_loader1, color_loader2) with the clean
        confusion loop, resumed where the previous call stopped; stops once `budget_s` seconds are spent.
        Returns (confusion float64 [C,C], label_map float64 [H,W], pixels done, seconds)."""
        s = self.solver
        M = np.zeros([self.C, self.C])
        label_map = np.zeros([self.H, self.W])
        done = 0
        t0 = time.perf_counter()
        with torch.no_grad():
            while True:
                if self._iter is None:
                    self._iter = self._batches()
                try:
                    data1, data2, target, x, y = next(self._iter)
                except StopIteration:
                    self._iter = None
                    continue
                output = self.net(data1, data2)
                pred = output.data.max(1, keepdim=True)[1]
                for i in range(len(target)):
                    M[int(pred[i].item())][int(target[i].item())] += 1
                    label_map[int(x[i])][int(y[i])] = int(pred[i])
                done += len(target)
                if time.perf_counter() - t0 > budget_s:
                    break
        return M, label_map, done, time.perf_counter() - t0

    def _batches(self):
        for loader in (self.solver.color_loader1, self.solver.color_loader2):
            for batch in loader:
                yield batch

    def metrics(self, M):
        return _quiet(self.rk.aa_oa, M)
