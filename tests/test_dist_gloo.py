"""world_size-2 gloo test of the multi-GPU host logic on CPU: each rank classifies its row band
(here with the oracle standing in for the device call), the int64 confusion matrices are all-reduced
and the disjoint label-map bands summed; the result must equal the single-process answer bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (REPO, os.path.join(REPO, 'dual-modal-fusion_b200')):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import dmf_oracle as orc
    from solver.mainsolver import row_band
    H, W, C = 37, 29, 6
    rng = np.random.default_rng(0)
    pred = rng.integers(0, C, (H, W))                  # what the device would have predicted
    label = rng.integers(0, C, (H, W))
    r0, r1 = row_band(H, rank, world)
    cm = torch.from_numpy(orc.confusion(pred[r0:r1].reshape(-1), label[r0:r1].reshape(-1), C).astype(np.int64))
    pm = torch.zeros((H, W), dtype=torch.uint8)
    pm[r0:r1] = torch.from_numpy(pred[r0:r1].astype(np.uint8))
    dist.all_reduce(cm)
    dist.all_reduce(pm)
    if rank == 0:
        want = orc.confusion(pred.reshape(-1), label.reshape(-1), C)
        out.put((np.array_equal(cm.numpy().astype(np.float64), want), np.array_equal(pm.numpy(), pred.astype(np.uint8))))
    dist.destroy_process_group()


def test_band_sharded_confusion_matrix_allreduce_gloo():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(timeout=5) == (True, True)
