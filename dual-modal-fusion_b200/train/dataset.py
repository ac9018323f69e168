"""Patch datasets with the reference's names and item contract (train/dataset.py:158-188, 248-282),
backed by the K1 gather kernel, plus PatchLoader: a DataLoader look-alike that keeps torch's own
samplers on the host (bit-identical index order and RNG consumption) and produces whole batches on
the device with one kernel launch instead of batch_size Python __getitem__ calls."""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

import dmf


def _scene_of(ms, pan, p, device):
    if isinstance(ms, dmf.Scene):
        return ms
    return dmf.Scene.from_padded(np.asarray(ms), np.asarray(pan), p, device)


class _PatchDataset(Dataset):
    tri = False

    def _setup(self, scene, label, x, y):
        self.scene = scene
        self.Label, self.x, self.y = label, x, y
        self.ms_size, self.pan_size = scene.p, scene.p * 4
        lab = np.asarray(label).reshape(-1)
        if not scene.has_labels and lab.size == scene.H * scene.W:
            scene.set_labels(lab.reshape(scene.H, scene.W).astype(np.uint8))
        self._flat = None

    def flat_index(self, index):
        """dataset index -> flat pixel index row*W + col (identity for the reference's xyl lists)."""
        if self._flat is None:
            xs = np.asarray(self.x).reshape(-1).astype(np.int64)
            ys = np.asarray(self.y).reshape(-1).astype(np.int64)
            self._flat = xs * self.scene.W + ys
        return self._flat[np.asarray(index, dtype=np.int64)]

    def gather_batch(self, indices):
        """indices: 1-D int64 (tensor/array) of dataset indices -> the collated batch, data on the
        device, x / y on the host exactly as default_collate would return them."""
        idx = np.asarray(indices, dtype=np.int64).reshape(-1)
        flat = self.flat_index(idx)
        out = self.scene.gather(flat, tri=self.tri, want_target=True)
        x = torch.from_numpy(flat // self.scene.W)
        y = torch.from_numpy(flat % self.scene.W)
        return tuple(out) + (x, y)

    def __getitem__(self, index):
        b = self.gather_batch([int(index)])
        data = [t[0].cpu() for t in b[:-2]]
        return tuple(data) + (int(b[-2][0]), int(b[-1][0]))

    def __len__(self):
        return len(self.x)


class dataset_dual(_PatchDataset):
    """dataset_dual(ms, pan, xyl, cfg): ms/pan are data_padding() outputs (or a dmf.Scene as `ms`)."""

    def __init__(self, ms, pan, xyl, cfg):
        scene = _scene_of(ms, pan, cfg['patch_size'], cfg.get('device', 'cuda:0'))
        self._setup(scene, xyl[2], xyl[0], xyl[1])


class dataset_tri(_PatchDataset):
    """dataset_tri(ms, pan, mspan, label, x, y, size): third raster on the PAN grid (IHS product)."""
    tri = True

    def __init__(self, ms, pan, mspan, label, x, y, size, device='cuda:0'):
        scene = _scene_of(ms, pan, size, device)
        if mspan is not None and not scene.has_mspan:
            scene.set_mspan(np.asarray(mspan))
        self._setup(scene, label, x, y)


class _IndexOnly(Dataset):
    def __init__(self, indices):
        self.indices = np.asarray(indices, dtype=np.int64)

    def __getitem__(self, i):
        return int(self.indices[i])

    def __len__(self):
        return len(self.indices)


class PatchLoader:
    """for (data1, data2, target, x, y) in PatchLoader(dataset, indices, batch_size, shuffle): ...

    Index order comes from a real torch DataLoader over the index list, so RandomSampler /
    SequentialSampler / BatchSampler behave (and consume the global RNG) exactly as in
    solver/basesolver.py:94-104."""

    def __init__(self, dataset, indices, batch_size, shuffle=False, rank=0, world=1):
        """rank / world: data-parallel sharding of every batch (torch.distributed): all ranks draw the SAME batches from the
        same sampler (same global seed, test.py:8) and rank r keeps elements r, r + world, ... of each, so the union of the
        ranks' sub-batches is exactly the reference's batch and every sample is visited once per epoch."""
        self.dataset = dataset
        self.indices = np.asarray(indices, dtype=np.int64)
        self.batch_size = batch_size
        self.rank, self.world = int(rank), int(world)
        self.last_global_batch = 0
        self._index_loader = DataLoader(_IndexOnly(self.indices), batch_size=batch_size, shuffle=shuffle, num_workers=0)

    def __iter__(self):
        for idx in self._index_loader:
            idx = idx.numpy()
            self.last_global_batch = int(idx.size)      # samples of this step over all ranks (the gradient mean's denominator)
            if self.world > 1:
                if idx.size < self.world:          # a tail batch that cannot feed every rank is dropped by ALL ranks alike
                    continue
                idx = idx[self.rank::self.world]
            yield self.dataset.gather_batch(idx)

    def __len__(self):
        return len(self._index_loader)
