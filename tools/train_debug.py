"""Stage-by-stage comparison of the native training step (csrc/train.cu) with an fp32 torch autograd run of the
oracle network on the same batch: pre-BatchNorm tensors Z, activations, logits, loss, every parameter gradient and
the running statistics.  Diagnostic tool (GPU); the pass/fail version is tests/test_gpu_train.py."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'dual-modal-fusion_b200'))
import torch
import torch.nn.functional as F

import dmf
from model.gmfnet import Net
from oracle.gmfnet_ref import Net as RefNet

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def planar_to_nchw(x, N, C, S):
    return x.view(N, C // 8, S, S, 8).permute(0, 1, 4, 2, 3).reshape(N, C, S, S).float()


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(x):
    """bf16 rounding with a straight-through gradient (the value the kernels store)"""
    return x + (x.bfloat16().float() - x).detach()


def rnd_grad(x):
    """identity whose gradient is rounded to bf16 (dZ / dA are stored as bf16)"""
    if x.requires_grad:
        x.register_hook(lambda g: g.bfloat16().float())
    return x


def run(p, C, N, swap, seed=0, emulate=False):
    cfg = {'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}, 'b200': {'max_train_batch': N}}
    torch.manual_seed(seed)
    ref = RefNet(cfg).cuda().train()
    # non-trivial BatchNorm affine parameters so their gradients matter
    with torch.no_grad():
        for blk in ('ms1', 'ms2', 'pan1', 'pan2', 'pan3', 'fuse'):
            bn = getattr(ref, blk)[1]
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
    net = Net(cfg)
    net.load_state_dict(ref.state_dict())
    net = net.cuda().train()
    g = torch.Generator(device='cuda').manual_seed(seed + 1)
    ms = torch.rand((N, 4, p, p), device='cuda', generator=g)
    pan = torch.rand((N, 1, 4 * p, 4 * p), device='cuda', generator=g)
    tgt = torch.randint(1, C, (N,), device='cuda', generator=g)

    # ---- fp32 reference with retained intermediates
    z = {}
    def block(name, x, pool):
        conv, bn = getattr(ref, name)[0], getattr(ref, name)[1]
        if not emulate:
            zz = F.conv2d(x, conv.weight, None, padding=conv.padding)        # bias cancels inside train-mode BN
            z[name] = zz
            y = torch.relu(bn(zz + conv.bias.view(1, -1, 1, 1)))
            return F.max_pool2d(y, 2) if pool else y
        # the kernels' precision: bf16 operands (the two stems see fp32-grade inputs), fp32 accumulation and batch
        # statistics, Z / activations / dZ / dA stored as bf16
        stem = name in ('ms1', 'pan1')
        w = conv.weight if stem else rnd(conv.weight)
        z32 = rnd_grad(F.conv2d(rnd_grad(x), w, None, padding=conv.padding))
        z[name] = z32
        mean, var = z32.mean((0, 2, 3)), z32.var((0, 2, 3), unbiased=False)
        zb = rnd(z32)
        y = (zb - mean.view(1, -1, 1, 1)) * (torch.rsqrt(var + bn.eps) * bn.weight).view(1, -1, 1, 1) + bn.bias.view(1, -1, 1, 1)
        y = torch.relu(y)
        if name == 'fuse':
            return y
        y = rnd(y)
        return F.max_pool2d(y, 2) if pool else y
    a1 = block('ms1', ms, False)
    m = block('ms2', a1, True)
    q = block('pan1', pan, True)
    q2 = block('pan2', q, True)
    q3 = block('pan3', q2, True)
    cat = torch.cat([m, q3], 1)
    f = block('fuse', cat, False)
    gvec = f.mean(dim=(2, 3))
    logits_ref = ref.fc2(torch.relu(ref.fc1(gvec)))
    loss_ref = F.cross_entropy(logits_ref, tgt)
    loss_ref.backward()

    # ---- native
    h = net.trainer()
    h.set_debug(swap)
    loss = h.step_patches(ms, pan, tgt)
    torch.cuda.synchronize()
    out = {'p': p, 'C': C, 'N': N, 'swap_lbo_sbo': swap, 'ref': 'bf16-emulating torch' if emulate else 'fp32 torch', 'loss': float(loss), 'loss_ref': float(loss_ref.detach())}
    S = {'ms1': p, 'ms2': p, 'pan1': 4 * p, 'pan2': 2 * p, 'pan3': p, 'fuse': p // 2}
    Cc = {'ms1': 64, 'ms2': 128, 'pan1': 32, 'pan2': 64, 'pan3': 128, 'fuse': 128}
    for k in S:
        zn = planar_to_nchw(h.buffer('Z_' + k, torch.bfloat16, (N, Cc[k] // 8, S[k], S[k], 8)), N, Cc[k], S[k])
        out['Z_' + k] = rel(zn, z[k])
    out['A1'] = rel(planar_to_nchw(h.buffer('A1', torch.bfloat16, (N, 8, p, p, 8)), N, 64, p), a1)
    out['B1'] = rel(planar_to_nchw(h.buffer('B1', torch.bfloat16, (N, 4, 2 * p, 2 * p, 8)), N, 32, 2 * p), q)
    out['B2'] = rel(planar_to_nchw(h.buffer('B2', torch.bfloat16, (N, 8, p, p, 8)), N, 64, p), q2)
    out['CAT'] = rel(planar_to_nchw(h.buffer('CAT', torch.bfloat16, (N, 32, p // 2, p // 2, 8)), N, 256, p // 2), cat)
    out['g'] = rel(h.buffer('g', torch.float32, (N, 128)), gvec)
    out['logits'] = rel(h.buffer('logits', torch.float32, (N, C)), logits_ref)
    grads = {}
    for (name, qn), (_, qr) in zip(net.named_parameters(), ref.named_parameters()):
        if name.endswith('.0.bias') and not name.startswith('fc'):
            grads[name] = 'abs %.1e (ref abs %.1e)' % (float(qn.grad.abs().max()), float(qr.grad.abs().max()) if qr.grad is not None else 0.0)
        else:
            grads[name] = round(rel(qn.grad, qr.grad), 5)
    out['grad_rel_err'] = grads
    rs = {}
    if not emulate:
        for (name, bn), (_, br) in zip(net.named_buffers(), ref.named_buffers()):
            rs[name] = round(rel(bn.float(), br.float()), 6) if bn.dtype.is_floating_point else (int(bn), int(br))
        out['running'] = rs
    return out


if __name__ == '__main__':
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    for emulate in (False, True):
        print(json.dumps(run(p, 12, N, 0, emulate=emulate), indent=None))
