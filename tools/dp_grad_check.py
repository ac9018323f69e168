"""Data-parallel gradient parity on 2+ GPUs (launch with torchrun): the all-reduced flat gradient of Net.train_step's
data-parallel path equals the sample-weighted mean of the per-rank gradients that ONE rank computes for the same sub-batches
(BatchNorm statistics are per sub-batch in both: DP-local BN).  Replaces the step of solver/mainsolver.py:49-55 run
under torch.distributed.  Prints one JSON line on rank 0; exit code 1 on mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_grad_check.py
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = 'cuda:%d' % local
dist.init_process_group('nccl', device_id=torch.device(dev))
import dmf
from model.gmfnet import Net
from oracle import dmf_oracle as orc

C, p, H, W = 8, 16, 96, 96
ms, pan, label = orc.synthetic_scene_structured(H, W, C - 1, seed=0, label_seed=1)
scene = dmf.Scene.from_raw(ms, pan, p, dev)
scene.set_labels(label)
labelled = np.flatnonzero(label.reshape(-1) != 0)
rng = np.random.default_rng(5)
batch = rng.choice(labelled, size=96 + 1, replace=False)              # odd: the sub-batches differ by one sample
subs = [batch[r::world] for r in range(world)]


def fresh():
    torch.manual_seed(3407)
    return Net({'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}, 'b200': {'max_train_batch': 128}}).to(dev).train()


class NoStep:                       # the check is on the gradient: no parameter update
    def step(self):
        pass


net = fresh()
net.train_step_scene(scene, torch.from_numpy(subs[rank]).to(dev), NoStep(), global_batch=len(batch))
g_dp = net.trainer().flat_grad.clone()
ok, worst = True, 0.0
if rank == 0:
    def solo_grad(r):
        solo = fresh()
        h = solo.trainer()
        h.reseat_grads()
        h.step_scene(scene, torch.from_numpy(subs[r]).to(dev))          # no collective: the raw local gradient of sub-batch r
        return h.flat_grad.clone()
    solos = [solo_grad(r) for r in range(world)]
    n = sum(len(s_) for s_ in subs)
    want = sum(g * len(s_) for g, s_ in zip(solos, subs)) / n
    plain = sum(solos) / world                                          # the UNWEIGHTED mean: must explain the result worse
    scale = want.abs().max().item()
    err = (g_dp - want).abs().max().item()
    err_plain = (g_dp - plain).abs().max().item()
    # The step is not bit-reproducible (fp32 / fp64 atomics in the weight-gradient and BatchNorm reductions, bf16 storage of Z / dZ: a
    # last-bit change of a sum can move a stored value by one bf16 ulp): the same sub-batch twice gives the noise floor of this check.
    noise = max((solo_grad(r) - solos[r]).abs().max().item() for r in range(world))
    worst = err / scale
    # Measured over repeated runs: rel is ~4e-5 most of the time and jumps to 5e-4 .. 9e-4 when one stored bf16 value flips (the rerun
    # noise shows the same two levels); the unweighted mean is off by 7e-3.  So: within 2e-3, and at least 3x closer than the unweighted mean.
    ok = worst <= 2e-3 and 3.0 * err < err_plain
    print(json.dumps({'check': 'data-parallel flat gradient == sample-weighted mean of the single-rank gradients of the same sub-batches',
                      'world': world, 'sub_batches': [len(s) for s in subs], 'max_abs_err': err, 'max_abs_grad': scale, 'rel': worst,
                      'rerun_noise_rel': noise / scale, 'unweighted_mean_rel': err_plain / scale, 'ok': ok}))
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if int(flag) else 0)
