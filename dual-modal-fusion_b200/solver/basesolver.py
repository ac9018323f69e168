"""BaseSolver — scene loading, preparation, index lists, datasets and loaders, with the reference's
attribute and method names (solver/basesolver.py:8-140).

What changed underneath: the rasters go to the GPU once (dmf.Scene: normalise + reflect-pad in HBM)
instead of becoming 546 MB float64 host arrays; the index lists are built vectorised; the five
loaders are PatchLoaders (torch samplers on the host, one gather kernel per batch).  `MS` / `PAN`
(the padded float64 arrays of the reference) are materialised only if somebody reads them.

Synthetic / in-memory scenes: put ``cfg['rasters'] = {'ms': ..., 'pan': ..., 'label': ...}`` and no
file is read.  Multi-GPU: when torch.distributed is initialised every rank holds the whole scene
(171 MB of uint16 at Hohhot size) and works on its own row band.
"""
import os
import time

import numpy as np
import torch

import torch.distributed as dist

import dmf
from function.function import data_padding, data_show, label_mat2np, read_tif, split_data, split_data_old
from indicators.kappa import aa_oa, expo_result
from train.dataset import PatchLoader, dataset_dual


class BaseSolver:
    def __init__(self, cfg):
        self.cfg = cfg
        self.task = cfg['task']
        self.TIME = cfg['time']
        self.time = cfg['index']
        self.EPOCH = cfg['epoch']
        self.epoch = 0
        self.DEVICE = cfg['device']
        self.timestamp = int(time.time())
        self.num_workers = cfg['threads'] if cfg.get('gpu_mode') else 0      # kept for interface parity; unused
        self.train_time = self.test_time = 0

        mem = cfg.get('rasters')
        self.ms = mem['ms'] if mem else read_tif(cfg, 'ms')
        self.pan = mem['pan'] if mem else read_tif(cfg, 'pan')
        if cfg['data_new'] == 1:
            self.train_label = np.load(cfg['data_address'] + 'train.npy')
            self.test_label = np.load(cfg['data_address'] + 'test.npy')

        self.scene = dmf.Scene.from_raw(np.asarray(self.ms), np.asarray(self.pan), cfg['patch_size'], self.DEVICE)
        self._MS = self._PAN = None

        if mem:
            label_np = np.asarray(mem['label'])
        else:
            if not os.path.exists(cfg['data_address'] + 'label.npy'):
                label_mat2np(cfg)
            label_np = np.load(cfg['data_address'] + 'label.npy', encoding='bytes', allow_pickle=True)
        data_show(label_np)
        self.label_np = label_np
        size = cfg['DATA_DICT'][cfg['data_city']]['size']
        if tuple(label_np.shape[:2]) != (self.scene.H, self.scene.W) or (int(size[0]), int(size[1])) != (self.scene.H, self.scene.W):
            # the index lists (split_data_old over cfg size), the label map and the rasters must describe the same grid: flat
            # pixel indices row*W + col are shared by all three
            raise ValueError('scene is %d x %d but label.npy is %s and DATA_DICT[%s].size is %s' %
                             (self.scene.H, self.scene.W, tuple(label_np.shape), cfg['data_city'], size))
        self.scene.set_labels(label_np.astype(np.uint8))
        if cfg['data_new'] == 1:
            xyl_matrix, self.traintest_index = split_data(self.train_label, self.test_label, label_np, cfg)
            _, self.matrix_ = split_data_old(label_np, cfg)
        else:
            xyl_matrix, self.matrix_ = split_data_old(label_np, cfg)
        if cfg['use_h5']:
            raise AttributeError("not finished")      # as in the reference (solver/basesolver.py:45-46)
        self.dataset = dataset_dual(self.scene, None, xyl_matrix, cfg)
        print('All dataset size:', len(self.dataset))
        self.records = {'Epoch': [], 'PSNR': [], 'SSIM': [], 'Loss': []}

    # the reference's padded float64 arrays, on demand (function/function.py:99-117)
    @property
    def MS(self):
        if self._MS is None:
            self._MS = data_padding(self.ms, self.cfg, 'ms')
        return self._MS

    @property
    def PAN(self):
        if self._PAN is None:
            self._PAN = data_padding(self.pan, self.cfg, 'pan')
        return self._PAN

    @staticmethod
    def dist_info():
        """(rank, world) of the default process group, (0, 1) without torch.distributed."""
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _loader(self, indices, batch_size, shuffle=False, shard=False):
        rank, world = self.dist_info() if shard else (0, 1)
        return PatchLoader(self.dataset, indices, batch_size, shuffle, rank, world)

    def dataloader(self):
        """Same subsets, same split arithmetic and same consumption of the global torch RNG as
        solver/basesolver.py:63-105 (random_split, then one RandomSampler per epoch)."""
        cfg = self.cfg
        if cfg['data_new'] == 1:
            train_idx = np.asarray(self.traintest_index[1], dtype=np.int64)
            test_all = np.asarray(self.traintest_index[2], dtype=np.int64)
            valid_size = int(cfg['verify_rate'] * len(test_all))
            parts = torch.utils.data.random_split(range(len(test_all)), [len(test_all) - valid_size, valid_size])
            test_idx, valid_idx = test_all[parts[0].indices], test_all[parts[1].indices]
            color1 = np.asarray(self.matrix_[1], dtype=np.int64)
        else:
            labelled = np.asarray(self.matrix_[1], dtype=np.int64)
            train_size = int(cfg['train_rate'] * len(labelled))
            valid_size = int(cfg['verify_rate'] * len(labelled))
            parts = torch.utils.data.random_split(range(len(labelled)),
                                                  [train_size, len(labelled) - train_size - valid_size, valid_size])
            train_idx, test_idx, valid_idx = (labelled[p.indices] for p in parts)
            color1 = labelled
        # data-parallel training: every rank takes its slice of each (identically drawn) batch; gradients are averaged
        self.train_loader = self._loader(train_idx, cfg['batchsize'], shuffle=True, shard=True)
        self.test_loader = self._loader(test_idx, cfg['test_batchsize'])
        self.valid_loader = self._loader(valid_idx, cfg['color_batchsize'])
        self.color_loader1 = self._loader(color1, cfg['test_batchsize'])
        self.color_loader2 = self._loader(np.asarray(self.matrix_[0], dtype=np.int64), cfg['test_batchsize'])

    def load_checkpoint(self, model_path):
        if not os.path.exists(model_path):
            raise FileNotFoundError
        ckpt = torch.load(model_path)
        self.epoch, self.records = ckpt['epoch'], ckpt['records']

    def save_checkpoint(self):
        self.ckp = {'epoch': self.epoch, 'records': self.records}

    def indicator(self):
        writer = self.dist_info()[0] == 0                 # files are written by rank 0 only (every rank holds the same matrix)
        if self.cfg['test']['save_matrix'] and writer:
            os.makedirs(self.cfg['RESULT_output'], exist_ok=True)
            np.save(self.cfg['RESULT_output'] + str(self.time) + "_matrix.npy", self.test_matrix)
        self.result = aa_oa(self.test_matrix)
        if self.cfg.get('RESULT_excel') and writer:
            expo_result(self.result, self.cfg, [self.train_time, self.test_time], self.time)
        return self.result

    def train(self):
        raise NotImplementedError

    def eval(self):
        raise NotImplementedError

    def run(self):
        while self.time < self.TIME:
            self.train()
            self.time += 1
