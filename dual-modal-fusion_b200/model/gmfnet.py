"""model.gmfnet.Net — the plug-in the reference loads with
``importlib.import_module('model.' + cfg['model_name'].lower()).Net(args=cfg)``
(solver/mainsolver.py:30-34) but never shipped.  Contract kept: an nn.Module with real Parameters
(state_dict / load_state_dict / optimizers work unchanged),
``forward(ms[B,4,p,p], pan[B,1,4p,4p]) -> logits[B,C]``.

Inference (eval mode, no grad) runs on the hand-written sm_100a kernels of libdmf_b200: fp32
CUDA-core stems, tcgen05/TMEM implicit-GEMM convolutions fed by TMA, fused head.  The packed bf16
weights are rebuilt whenever a parameter changes.  ``infer_scene`` is the fused whole-scene path
(gather + network + argmax + confusion matrix without materialising patches).

Training mode builds the same graph from torch ops so that autograd / Adam work (Solver.train);
native backward kernels are the next scope row.  There is no CPU path: eval-mode
forward on CPU tensors raises.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import dmf

MS_BANDS = 4
WIDTHS = {'ms1': (MS_BANDS, 64), 'ms2': (64, 128), 'pan1': (1, 32), 'pan2': (32, 64), 'pan3': (64, 128)}
C_FUSE, C_HID = 128, 64


def _conv_bn(cin, cout, k):
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=k // 2, bias=True), nn.BatchNorm2d(cout))


class Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_classes = int(args['Categories_Number'])
        self.patch = int(args['patch_size'])
        if str(args.get('schedule', {}).get('activate', 'Relu')).lower() != 'relu':
            raise ValueError("gmfnet: schedule.activate must be Relu")
        self.max_batch = int(args.get('b200', {}).get('max_batch', 16384)) if isinstance(args.get('b200'), dict) else 16384
        for name, (cin, cout) in WIDTHS.items():
            setattr(self, name, _conv_bn(cin, cout, 3))
        self.fuse = _conv_bn(128 + 128, C_FUSE, 1)
        self.fc1 = nn.Linear(C_FUSE, C_HID)
        self.fc2 = nn.Linear(C_HID, self.num_classes)
        self._native = None
        self._native_key = None

    # ---------------------------------------------------------------- native (sm_100a) inference
    def _weights_key(self):
        ts = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in ts)

    def native(self):
        """NetHandle with weights matching the current parameters."""
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("gmfnet inference needs the model on a CUDA device (no CPU fallback); got %s" % dev)
        key = (str(dev),) + self._weights_key()
        if self._native is None or self._native.device != str(dev):
            self._native = dmf.NetHandle(self.patch, self.num_classes, self.max_batch, str(dev))
            self._native_key = None
        if key != self._native_key:
            self._native.load_state_dict(self.state_dict())
            self._native_key = key
        return self._native

    def infer_scene(self, scene, row0=0, row1=None, pred_map=None, cm=None):
        return self.native().infer_scene(scene, row0, row1, pred_map, cm)

    # ---------------------------------------------------------------- autograd graph for training
    def _graph(self, ms, pan):
        r, mp = F.relu, F.max_pool2d
        m = r(self.ms1(ms))
        m = mp(r(self.ms2(m)), 2)
        q = mp(r(self.pan1(pan)), 2)
        q = mp(r(self.pan2(q)), 2)
        q = mp(r(self.pan3(q)), 2)
        f = r(self.fuse(torch.cat([m, q], dim=1)))
        return self.fc2(r(self.fc1(f.mean(dim=(2, 3)))))

    def forward(self, ms, pan):
        if self.training or torch.is_grad_enabled():
            if not ms.is_cuda:
                raise RuntimeError("gmfnet: tensors must be on a CUDA device (no CPU path)")
            return self._graph(ms, pan)
        return self.native().forward_patches(ms.float(), pan.float())

    def _apply(self, fn, *a, **k):
        self._native_key = None
        return super()._apply(fn, *a, **k)
