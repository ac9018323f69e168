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


def _worker_shards(rank, world, port, out):
    """Host logic of the data-parallel paths under a real process group (gloo, CPU): PatchLoader batch sharding, the row-band
    share of a test sample set + all-reduced matrix (Solver.test), and the sample-weighted gradient average (Net._sync_grads)."""
    for p in (REPO, os.path.join(REPO, 'dual-modal-fusion_b200')):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import dmf_oracle as orc
    from solver.mainsolver import indices_in_band, row_band
    from train.dataset import PatchLoader
    from model.gmfnet import Net

    class FakeDataset:                                   # gather_batch would launch K1; the index flow is what is tested
        def gather_batch(self, idx):
            return np.asarray(idx)

    H, W, C = 41, 23, 5
    rng = np.random.default_rng(0)
    pred = rng.integers(0, C, (H, W))
    label = rng.integers(0, C, (H, W))
    test_idx = np.sort(rng.choice(H * W, size=301, replace=False))
    # (1) Solver.test under torch.distributed: every sample is counted exactly once
    r0, r1 = row_band(H, rank, world)
    mine = indices_in_band(test_idx, W, r0, r1)
    cm = torch.from_numpy(orc.confusion(pred.reshape(-1)[mine], label.reshape(-1)[mine], C).astype(np.int64))
    dist.all_reduce(cm)
    ok_cm = int(cm.sum()) == len(test_idx) and np.array_equal(
        cm.numpy().astype(np.float64), orc.confusion(pred.reshape(-1)[test_idx], label.reshape(-1)[test_idx], C))
    # (2) PatchLoader sharding: same batches on every rank (same seed), rank r keeps elements r::world, short tails are dropped by all
    train_idx = rng.permutation(H * W)[:101]
    torch.manual_seed(3407)
    got = [b for b in PatchLoader(FakeDataset(), train_idx, 8, shuffle=True, rank=rank, world=world)]
    torch.manual_seed(3407)
    full = [b for b in PatchLoader(FakeDataset(), train_idx, 8, shuffle=True)]
    full = [b for b in full if len(b) >= world]
    ok_loader = len(got) == len(full) and all(np.array_equal(g, f[rank::world]) for g, f in zip(got, full))
    steps = torch.tensor([len(got)])
    lo, hi = steps.clone(), steps.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok_loader = ok_loader and int(lo) == int(hi)          # every rank runs the same number of steps (no collective is left hanging)
    # (3) gradient of the global batch = sum_r n_r g_r / sum_r n_r, one collective
    class H_:
        pass
    h = H_()
    h.flat_grad = torch.zeros(6, dtype=torch.float32)
    n_local = 3 + rank                                     # unequal sub-batches
    g_local = torch.arange(6, dtype=torch.float32) * (rank + 1)
    h.flat_grad.copy_(g_local)
    Net._sync_grads(h, n_local, sum(3 + r for r in range(world)))
    want = sum((3 + r) * torch.arange(6, dtype=torch.float32) * (r + 1) for r in range(world)) / sum(3 + r for r in range(world))
    ok_grad = torch.allclose(h.flat_grad, want, rtol=1e-6)
    if rank == 0:
        out.put((bool(ok_cm), bool(ok_loader), bool(ok_grad)))
    dist.destroy_process_group()


def test_data_parallel_host_logic_gloo():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_shards, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(timeout=5) == (True, True, True)
