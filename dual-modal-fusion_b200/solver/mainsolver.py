"""Solver — train / test / color / run with the reference's names (solver/mainsolver.py:12-209).

test():  the reference's loop is argmax + ``M[pred][target] += 1`` per sample with two .item() syncs
         (solver/mainsolver.py:139-141; clean copy train/test.py:58-60).  Here each batch is one
         fused argmax+histogram kernel (K4) accumulating an int64 matrix on the device.  The shipped
         file evaluates only the first batch because of a debugging `break` (:142) and a t-SNE plot
         (:110-136); the full-loader semantics of train/test.py are the default and
         ``cfg['test']['first_batch_only']`` reproduces the shipped behaviour.
color(): the reference runs every pixel through the net in 300-patch batches and writes
         label_np[x][y] per sample, then paints in an HxW Python loop (:167-201).  Here one fused call
         classifies the rank's row band straight from the scene (no patches are materialised), the
         K5 kernel paints, and with torch.distributed the int64 confusion matrix is all-reduced and
         the label-map bands are gathered.
"""
import importlib
import os
import time

import numpy as np
import torch
import torch.distributed as dist
from PIL import Image

import dmf
from solver.basesolver import BaseSolver
from utils.utils import make_loss, make_optimizer, make_scheduler, save_checkpoint

try:
    from tqdm import tqdm
except ImportError:                       # progress bars are cosmetic
    def tqdm(it, **k):
        return it


def row_band(H, rank, world):
    """Contiguous row band of rank `rank`: ceil(H/world) rows each, the last ranks one short."""
    base, extra = divmod(H, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def indices_in_band(flat_idx, W, r0, r1):
    """The flat pixel indices (row*W + col) whose row lies in [r0, r1): a rank's share of a sample set under row-band sharding.
    The shares of all ranks partition the set, so the all-reduced confusion matrix counts every sample exactly once."""
    flat_idx = np.asarray(flat_idx, dtype=np.int64)
    rows = flat_idx // W
    return flat_idx[(rows >= r0) & (rows < r1)]


class Solver(BaseSolver):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.model = self.cur_model = None
        self.matrix = None
        if self.cfg['train']['pretrained']:
            self.init_model()

    def init_model(self):
        lib = importlib.import_module('model.' + self.cfg['model_name'].lower())
        self.model = lib.Net(args=self.cfg)
        self.optimizer = make_optimizer(self.cfg, self.model.parameters())
        self.loss = make_loss(self.cfg['schedule']['loss'], self.cfg)
        self.scheduler = make_scheduler(self.optimizer, self.cfg)

    def _bar(self, loader):
        return loader if self.cfg['nohup'] else tqdm(loader, leave=True)

    def _weights(self, kind):
        return self.cfg['RESULT_output'] + str(self.time) + kind

    # ------------------------------------------------------------------ train
    def train(self):
        t0 = time.time()
        cfg = self.cfg
        save_best = cfg['train']['save_best']
        best_loss, best_epoch = (float('inf'), 0) if save_best else (None, None)
        if not cfg['train']['pretrained']:
            self.init_model()
        self.cur_model = self.model.to(self.DEVICE)
        rank, world = self.dist_info()
        # Data-parallel training (torchrun): every rank holds the same initial weights (same seed, test.py:8), takes its slice of
        # every batch (PatchLoader) and the flat gradient is averaged (Net.train_step), so the parameters stay identical on all
        # ranks.  BatchNorm normalises with the rank's own sub-batch statistics; the running statistics are averaged over the
        # ranks at the end of every epoch.  Files are written by rank 0.
        if rank == 0:
            os.makedirs(cfg['RESULT_output'], exist_ok=True)
        self.val_losses = []
        while self.epoch < self.EPOCH:
            self.cur_model.train()
            bar = self._bar(self.train_loader)
            # CrossEntropyLoss + a native model: one library call does zero_grad / forward / loss / backward and the
            # optimizer one kernel (model.train_step); any other loss runs the reference's loop text, with the
            # network's backward still native (autograd.Function in model/gmfnet.py)
            fused = isinstance(self.loss, torch.nn.CrossEntropyLoss) and hasattr(self.cur_model, 'train_step') \
                and not cfg.get('b200', {}).get('autograd_loop', False)
            for data1, data2, target, _, _ in bar:
                data1, data2, target = data1.to(self.DEVICE), data2.to(self.DEVICE), target.to(self.DEVICE)
                if fused:
                    loss = self.cur_model.train_step(data1, data2, target, self.optimizer,
                                                     global_batch=getattr(self.train_loader, 'last_global_batch', None))
                else:
                    self.optimizer.zero_grad()
                    loss = self.loss(self.cur_model(data1, data2), target.long())
                    loss.backward()
                    self.optimizer.step()
                if cfg['nohup']:
                    print("{} times {}th epoch is trained".format(self.time, self.epoch))
                else:
                    bar.set_postfix(ls=loss.item(), b_ep=best_epoch, ep=self.epoch, tm=self.time, m='train', d=cfg['device'])
            if cfg['schedule']['if_scheduler']:
                self.scheduler.step()
            if world > 1:
                for b in self.cur_model.buffers():
                    if torch.is_floating_point(b):
                        dist.all_reduce(b)
                        b.div_(world)
            if save_best:
                self.cur_model.eval()
                val_loss = torch.zeros((), device=self.DEVICE)
                with torch.no_grad():
                    for data1, data2, target, _, _ in self.valid_loader:
                        out = self.cur_model(data1, data2)
                        val_loss += self.loss(out, target.long()) * data1.size(0)     # accumulated on the device
                val_loss = float(val_loss)
                self.val_losses.append(val_loss)
                if val_loss < best_loss:
                    best_loss, best_epoch = val_loss, self.epoch
                    if rank == 0:
                        torch.save(self.cur_model.state_dict(), self._weights('_weights.pth'))
            if rank == 0:
                save_checkpoint(self.cur_model, self.optimizer, self._weights('_curweights.pth'))
            self.epoch += 1
        self.best_epoch, self.best_loss = best_epoch, best_loss
        if world > 1:
            dist.barrier()                # the checkpoints rank 0 wrote are complete before any rank reads them
        self.train_time = time.time() - t0
        self.epoch = 0

    def _load_for_eval(self):
        if self.model is None:
            self.init_model()
        if self.cur_model is None:
            self.cur_model = self.model.to(self.DEVICE)
        best, cur = self._weights('_weights.pth'), self._weights('_curweights.pth')
        # the reference's choice (solver/mainsolver.py:95-98): the best-validation weights when save_best, else the last epoch's
        # (its `_curweights.pth` holds {state_dict, optimizer}, utils/utils.py:82-88).  A missing file raises, as torch.load does
        # there: an untrained network must not produce plausible-looking metrics.  `test.allow_random_init` (tests, benchmarks
        # of the inference path) evaluates the model as it stands instead.
        if self.cfg['train']['save_best']:
            path, pick = best, (lambda ck: ck)
        else:
            path, pick = cur, (lambda ck: ck['state_dict'])
        if os.path.exists(path):
            self.cur_model.load_state_dict(pick(torch.load(path, map_location=self.DEVICE)))
        elif not self.cfg['test'].get('allow_random_init', False):
            raise FileNotFoundError('no checkpoint %s (train first, or set test.allow_random_init)' % path)
        self.cur_model.eval()

    # ------------------------------------------------------------------ test
    def test(self):
        """Confusion matrix of the test loader's sample set (solver/mainsolver.py:104-141).  Large sample sets are taken out of
        one scene-dense pass over the whole scene (the counts do not depend on the order or batching of the samples); small
        ones, and the reference's first-batch-only behaviour, run the loader batches through the per-patch kernels."""
        t0 = time.time()
        self._load_for_eval()
        C = self.cfg['Categories_Number']
        cm = torch.zeros((C, C), dtype=torch.int64, device=self.DEVICE)
        first_only = bool(self.cfg['test'].get('first_batch_only', False))
        rank, world = self.dist_info()
        sharded = world > 1 and not first_only and self.cfg.get('shard_test', True)
        idx = getattr(self.test_loader, 'indices', None)
        H, W = self.scene.H, self.scene.W
        flat = self.dataset.flat_index(idx) if idx is not None else None        # dataset indices -> flat pixel indices row*W + col
        # dense whole-scene pass ~ 10x cheaper per pixel than a per-patch evaluation
        dense = (not first_only and flat is not None and getattr(self.cur_model, 'dense', False) and 8 * len(flat) >= H * W)
        with torch.no_grad():
            if dense:
                # under torch.distributed every rank classifies its row band and counts the samples that fall into it
                r0, r1 = row_band(H, rank, world) if sharded else (0, H)
                mine = indices_in_band(flat, W, r0, r1)
                if flat.size and (flat.min() < 0 or flat.max() >= H * W):
                    raise IndexError('test sample outside the %d x %d scene' % (H, W))
                pred_map, _ = self.cur_model.infer_scene(self.scene, r0, r1, cm=torch.zeros_like(cm))
                dmf.confusion_at(pred_map, self.scene_labels(), torch.from_numpy(mine), C, cm=cm)
            else:
                loader = self.test_loader
                if sharded:                       # each rank runs its slice of every batch
                    loader = self._loader(idx, self.cfg['test_batchsize'], shard=True)
                for data1, data2, target, _, _ in self._bar(loader):
                    out = self.cur_model(data1, data2)
                    dmf.argmax_confusion(out, target, C, cm=cm, want_pred=False)
                    if first_only:
                        break
        if sharded:
            dist.all_reduce(cm)
        self.test_time = time.time() - t0
        self.test_matrix = cm.cpu().numpy().astype(np.float64)
        self.indicator()

    def scene_labels(self):
        """the label map on the device, uint8 [H, W] (the scene's own grid)"""
        if getattr(self, '_label_dev', None) is None:
            self._label_dev = torch.from_numpy(np.ascontiguousarray(np.asarray(self.label_np)[:self.scene.H, :self.scene.W]).astype(np.uint8)).to(self.DEVICE)
        return self._label_dev

    # ------------------------------------------------------------------ whole-scene inference
    def classify_scene(self):
        """Every pixel of the scene: (pred_map u8 [H,W] on the device, confusion matrix float64).
        Row-band sharded when torch.distributed is initialised (one int64 all-reduce of C*C)."""
        self._load_for_eval()
        H = self.scene.H
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        r0, r1 = row_band(H, rank, world)
        pred_map, cm = self.cur_model.infer_scene(self.scene, r0, r1)
        if world > 1:
            dist.all_reduce(cm)
            dist.all_reduce(pred_map)          # bands are disjoint and zero elsewhere: sum == gather
        self.scene_matrix = cm.cpu().numpy().astype(np.float64)
        return pred_map, self.scene_matrix

    def color(self):
        pred_map, _ = self.classify_scene()
        colors = self.cfg['DATA_DICT'][self.cfg['data_city']]['color']
        labelled = torch.from_numpy(np.asarray(self.label_np) != 0).to(pred_map.device)
        map1 = torch.where(labelled, pred_map, torch.zeros_like(pred_map)) if self.cfg['color']['supervised'] else torch.zeros_like(pred_map)
        map2 = pred_map if self.cfg['color']['unsupervised'] else map1
        self.label_np1, self.label_np2 = map1.cpu().numpy().astype(np.float64), map2.cpu().numpy().astype(np.float64)
        if self.dist_info()[0] != 0:
            return                                # every rank holds the whole label map after the all-reduce; rank 0 paints and writes
        os.makedirs(self.cfg['RESULT_output'], exist_ok=True)
        for tag, m in (('_pic_1.png', map1), ('_pic_2.png', map2)):
            rgb = dmf.paint_labels(m, colors).cpu().numpy()
            if self.cfg['color']['supervised']:
                Image.fromarray(rgb).save(self.cfg['RESULT_output'] + str(self.time) + tag)

    def run(self):
        while self.time < self.TIME:
            self.dataloader()
            self.train() if self.cfg['train']['index'] else None
            self.test() if self.cfg['test']['index'] else None
            self.color() if self.cfg['color']['index'] else None
            self.time += 1
