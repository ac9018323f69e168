"""model.gmfnet.Net — the plug-in the reference loads with
``importlib.import_module('model.' + cfg['model_name'].lower()).Net(args=cfg)``
(solver/mainsolver.py:30-34) but never shipped.  Contract kept: an nn.Module with real Parameters
(state_dict / load_state_dict / optimizers work unchanged),
``forward(ms[B,4,p,p], pan[B,1,4p,4p]) -> logits[B,C]``.

Inference (eval mode, no grad) runs on the hand-written sm_100a kernels of libdmf_b200: fp32
CUDA-core stems, tcgen05/TMEM implicit-GEMM convolutions fed by TMA, fused head.  The packed bf16
weights are rebuilt whenever a parameter changes.  ``infer_scene`` is the fused whole-scene path
(gather + network + argmax + confusion matrix without materialising patches); by default it runs the
scene-dense evaluation (csrc/dense.cu: patches at stride 1 share every layer output that sits at the
same scene position with the same border class), ``b200.dense: false`` in the config selects the
per-patch kernels.

Training (train mode, or grad enabled) also runs on libdmf_b200 (csrc/train.cu): ``forward`` returns
logits attached to a torch.autograd.Function whose backward is the native backward pass (BatchNorm
batch statistics, max-pool routing, tcgen05 dgrad / wgrad), so the reference's loop
``loss = criterion(model(a, b), t.long()); loss.backward(); optimizer.step()``
(solver/mainsolver.py:49-55) works unchanged with any torch loss / optimizer.  ``train_step`` is the
fused fast path (forward + CrossEntropyLoss + backward in one library call, gradients written into one
flat buffer -> a single NCCL all-reduce under torch.distributed, Adam in one kernel).
There is no CPU path and no torch-op fallback: forward on CPU tensors raises.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

import dmf

MS_BANDS = 4
WIDTHS = {'ms1': (MS_BANDS, 64), 'ms2': (64, 128), 'pan1': (1, 32), 'pan2': (32, 64), 'pan3': (64, 128)}
C_FUSE, C_HID = 128, 64


def _conv_bn(cin, cout, k):
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=k // 2, bias=True), nn.BatchNorm2d(cout))


class _NativeTrainFn(torch.autograd.Function):
    """logits = net(ms, pan) in train mode; the parameters are inputs so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, net, ms, pan, *params):
        h = net.trainer()
        ctx.h, ctx.n_params = h, len(params)
        out = h.forward(ms, pan)
        h.forward_id += 1                             # the handle keeps ONE set of activations: those of the latest forward
        ctx.forward_id = h.forward_id
        return out

    @staticmethod
    def backward(ctx, dlogits):
        h = ctx.h
        if ctx.forward_id != h.forward_id:
            raise RuntimeError("gmfnet: backward through a forward whose activations were overwritten by a later forward of the "
                               "same model (the native training handle keeps one activation workspace); call backward before "
                               "the next forward, or use Net.train_step per micro-batch")
        # flat_grad is the storage of every p.grad (autograd ACCUMULATES what this function returns into it): run the native
        # backward on a zeroed buffer, hand out a copy, put the accumulated gradients back untouched
        keep = h.flat_grad.clone()
        h.flat_grad.zero_()
        h.backward(dlogits)
        g = h.flat_grad.clone()
        h.flat_grad.copy_(keep)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(g)                        # data-parallel: the ranks' gradients are averaged here too
            g.div_(dist.get_world_size())
        grads = tuple(g[off:off + k].view(q.shape) for q, off, k in h._views)
        return (None, None, None) + grads


class Net(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.num_classes = int(args['Categories_Number'])
        self.patch = int(args['patch_size'])
        if str(args.get('schedule', {}).get('activate', 'Relu')).lower() != 'relu':
            raise ValueError("gmfnet: schedule.activate must be Relu")
        b200 = args.get('b200') if isinstance(args.get('b200'), dict) else {}
        self.max_batch = int(b200.get('max_batch', 16384))
        self.max_train_batch = int(b200.get('max_train_batch', max(512, int(args.get('batchsize', 0) or 0))))
        # whole-scene inference: scene-dense maps (default) or the per-patch kernels; anchor rows per dense pass
        self.dense = bool(b200.get('dense', True))
        self.dense_band = int(b200.get('dense_band', 512))
        for name, (cin, cout) in WIDTHS.items():
            setattr(self, name, _conv_bn(cin, cout, 3))
        self.fuse = _conv_bn(128 + 128, C_FUSE, 1)
        self.fc1 = nn.Linear(C_FUSE, C_HID)
        self.fc2 = nn.Linear(C_HID, self.num_classes)
        self._native = None
        self._native_key = None
        self._trainer = None

    # ---------------------------------------------------------------- native (sm_100a) inference
    def _weights_key(self):
        ts = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _device(self):
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError("gmfnet needs the model on a CUDA device (no CPU fallback); got %s" % dev)
        return dev

    def native(self):
        """NetHandle with weights matching the current parameters."""
        dev = self._device()
        key = (str(dev),) + self._weights_key()
        if self._native is None or self._native.device != str(dev):
            self._native = dmf.NetHandle(self.patch, self.num_classes, self.max_batch, str(dev))
            self._native.set_dense(self.dense, self.dense_band)
            self._native_key = None
        if key != self._native_key:
            self._native.load_state_dict(self.state_dict())
            self._native_key = key
        return self._native

    def infer_scene(self, scene, row0=0, row1=None, pred_map=None, cm=None, use_mspan=False):
        """Whole-band inference; use_mspan feeds the scene's IHS product (Scene.set_mspan) as the PAN input, the evaluation
        counterpart of train_step_scene(..., use_mspan=True)."""
        h = self.native()
        h.set_pan_source(use_mspan)
        return h.infer_scene(scene, row0, row1, pred_map, cm)

    # ---------------------------------------------------------------- native training
    def trainer(self):
        """TrainHandle bound to this module's (flattened) parameters."""
        dev = self._device()
        if self._trainer is None or self._trainer.stale():
            if self._trainer is not None:
                self._trainer.close()
            self._trainer = dmf.TrainHandle(self, self.patch, self.num_classes, self.max_train_batch, str(dev))
        # the library updates parameters / BatchNorm buffers through raw pointers (no torch version bump):
        # whatever the inference handle packed is stale after any training call
        self._native_key = None
        return self._trainer

    def train_step(self, ms, pan, target, optimizer, global_batch=None):
        """One fused step of Solver.train(): zero_grad -> forward -> CrossEntropyLoss(mean) -> backward ->
        (all-reduce of the flat gradient under torch.distributed) -> optimizer.step().  Returns the loss tensor.
        global_batch: samples of this step over ALL ranks (PatchLoader.last_global_batch); None = every rank holds ms.shape[0]."""
        h = self.trainer()
        h.reseat_grads()
        loss = h.step_patches(ms, pan, target)
        self._sync_grads(h, ms.shape[0], global_batch)
        optimizer.step()
        return loss

    def train_step_scene(self, scene, flat_idx, optimizer, use_mspan=False, global_batch=None):
        """The same with the batch cropped from the device scene (K1 + IHS-product input fused in front)."""
        h = self.trainer()
        h.reseat_grads()
        loss = h.step_scene(scene, flat_idx, use_mspan)
        self._sync_grads(h, int(flat_idx.numel() if hasattr(flat_idx, 'numel') else len(flat_idx)), global_batch)
        optimizer.step()
        return loss

    @staticmethod
    def _sync_grads(h, n_local, n_global=None):
        """Data-parallel gradient of the GLOBAL batch: each rank's gradient is the mean over its n_local samples, so the global
        mean is sum_r (n_r / n_global) g_r — the ranks' sub-batches may differ by one sample.  n_global is known on every rank
        without communication (all ranks draw the same batch and slice it); one scale kernel + ONE all-reduce of the flat buffer."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            world = dist.get_world_size()
            n_global = n_local * world if n_global is None else n_global
            h.flat_grad.mul_(float(n_local) / float(n_global))
            dist.all_reduce(h.flat_grad)                          # one bucket: the whole model

    def forward(self, ms, pan, mspan=None):
        """forward(ms, pan) as solver/mainsolver.py:52 calls it; forward(ms, pan, mspan) is the reference's 3-input call
        (mode '3', train/train.py:44-52, fed by dataset_tri): an IHS-input model reads the IHS product in its PAN branch (the
        product replaces PAN's intensity: it equals PAN to 2e-16, image_convert/IHS.py:40-54), so `mspan` takes `pan`'s place."""
        if mspan is not None:
            pan = mspan
        if not ms.is_cuda:
            raise RuntimeError("gmfnet: tensors must be on a CUDA device (no CPU path)")
        if self.training:
            return _NativeTrainFn.apply(self, ms, pan, *[q for q in self.parameters()])
        if torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters()):
            raise RuntimeError("gmfnet: eval-mode forward with autograd enabled is not implemented natively; "
                               "wrap inference in torch.no_grad() or call .train()")
        return self.native().forward_patches(ms.float(), pan.float())

    def _apply(self, fn, *a, **k):
        self._native_key = None
        return super()._apply(fn, *a, **k)
