"""GPU parity of the native training step (csrc/train.cu, wgrad_tc.cuh) — SURVEY.md 8 row a10, the loop of
solver/mainsolver.py:49-55 with utils/utils.py:12,28-29.

The reference trains in fp32 (cuDNN); the kernels use bf16 operands with fp32 accumulation.  Tolerances, stated once:

  * single kernels on exactly representable (small-integer) data — forward conv, dgrad, wgrad, PAN-stem wgrad — must be
    BIT-EXACT against a float64 torch computation (fp32 accumulation of small integers is exact, so any deviation is
    an indexing / layout bug, not rounding);
  * softmax cross-entropy and Adam against torch in fp32: 1e-6 relative;
  * one whole step against a torch autograd run that rounds to bf16 at the same storage points (EMU): loss 1e-4,
    logits 2e-3, every parameter gradient within GRAD_EMU relative L2 error.  The residual is not accumulated rounding:
    one-ulp differences in a bf16 activation flip a ReLU / max-pool decision for ~1e-3 of the elements, and a flipped
    fraction f shows up as ~sqrt(f) in a gradient norm;
  * against the pure fp32 oracle: loss 1e-3 relative, logits 1e-2, gradient cosine similarity >= GRAD_COS_FP32;
  * a 30-step run from the same initialisation must track the fp32 torch loss curve (final loss within 10 %).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.gmfnet_ref import Net as RefNet

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GRAD_EMU = 8e-2
GRAD_COS_FP32 = 0.95
BLOCKS = ('ms1', 'ms2', 'pan1', 'pan2', 'pan3', 'fuse')


@pytest.fixture(scope='module')
def dmf():
    import dmf as m
    return m


@pytest.fixture(autouse=True)
def strict_fp32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def cfg(p, C, nb):
    return {'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}, 'b200': {'max_train_batch': nb}}


def to_c8(x):
    N, C, H, W = x.shape
    return x.view(N, C // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16)


def from_c8(y):
    N, Cc, H, W, _ = y.shape
    return y.float().permute(0, 1, 4, 2, 3).reshape(N, Cc * 8, H, W)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def make_pair(p, C, nb, seed=0):
    """(oracle fp32 net in train mode, native Net) with identical parameters and non-trivial BN affine terms."""
    from model.gmfnet import Net
    torch.manual_seed(seed)
    ref = RefNet(cfg(p, C, nb)).to(DEV).train()
    with torch.no_grad():
        for blk in BLOCKS:
            bn = getattr(ref, blk)[1]
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
    net = Net(cfg(p, C, nb))
    net.load_state_dict(ref.state_dict())
    return ref, net.to(DEV).train()


def batch(p, C, N, seed=1):
    g = torch.Generator(device=DEV).manual_seed(seed)
    ms = torch.rand((N, 4, p, p), device=DEV, generator=g)
    pan = torch.rand((N, 1, 4 * p, 4 * p), device=DEV, generator=g)
    tgt = torch.randint(1, C, (N,), device=DEV, generator=g)
    return ms, pan, tgt


# ---------------------------------------------------------------------------------------------- single kernels, exact
GEOM = {  # layer -> (input buffer, cin, cout, S as a function of p, taps)
    'ms2': ('A1', 64, 128, lambda p: p, 9), 'pan2': ('B1', 32, 64, lambda p: 2 * p, 9),
    'pan3': ('B2', 64, 128, lambda p: p, 9), 'fuse': ('CAT', 256, 128, lambda p: p // 2, 1),
}


def int_tensor(shape, lo, hi, seed, density=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randint(lo, hi + 1, shape, device=DEV, generator=g).float()
    if density < 1.0:
        x = x * (torch.rand(shape, device=DEV, generator=g) < density)
    return x


@pytest.mark.parametrize('p,N', [(16, 5), (8, 7), (32, 3)])
@pytest.mark.parametrize('layer', ['ms2', 'pan2', 'pan3', 'fuse'])
def test_conv_wgrad_dgrad_fwd_exact_on_integers(dmf, layer, p, N):
    C = 6
    _, net = make_pair(p, C, 8)
    h = net.trainer()
    buf, cin, cout, Sf, taps = GEOM[layer]
    S = Sf(p)
    k = 3 if taps == 9 else 1
    conv = getattr(net, layer)[0]
    # weights: multiples of 1/8 in [-1, 1] (exact in bf16); activations / gradients: small integers
    with torch.no_grad():
        conv.weight.copy_(int_tensor(conv.weight.shape, -8, 8, 11) / 8)
    a = int_tensor((N, cin, S, S), -2, 2, 12, density=0.5)
    dz = int_tensor((N, cout, S, S), -2, 2, 13, density=0.5)
    h.buffer(buf, torch.bfloat16, (N, cin // 8, S, S, 8), alias=True).copy_(to_c8(a))
    h.buffer('dZ', torch.bfloat16, (N, cout // 8, S, S, 8), alias=True).copy_(to_c8(dz))
    h.debug_op('pack', layer, N)
    # forward (raw Z + batch statistics are checked through the BN-dependent tests below)
    h.debug_op('fwd', layer, N)
    z = from_c8(h.buffer('Z_' + layer, torch.bfloat16, (N, cout // 8, S, S, 8)))
    want = F.conv2d(a.double(), conv.weight.detach().double(), None, padding=k // 2)
    assert torch.equal(z.double(), want.float().bfloat16().double()), 'forward conv %s differs' % layer
    # wgrad: accumulated into the bound gradient (zero first)
    h.flat_grad.zero_()
    h.debug_op('wgrad', layer, N)
    torch.cuda.synchronize()
    dw = torch.nn.grad.conv2d_weight(a.double(), conv.weight.shape, dz.double(), padding=k // 2)
    assert torch.equal(conv.weight.grad.double(), dw), 'wgrad %s: max |d| = %g' % (layer, float((conv.weight.grad.double() - dw).abs().max()))
    # dgrad
    h.debug_op('dgrad', layer, N)
    name, ch = ('dCAT', cin // 8) if layer == 'fuse' else ('dA', cin // 8)
    da = from_c8(h.buffer(name, torch.bfloat16, (N, ch, S, S, 8)))
    want = torch.nn.grad.conv2d_input(a.shape, conv.weight.detach().double(), dz.double(), padding=k // 2)
    assert torch.equal(da.double(), want.float().bfloat16().double()), 'dgrad %s differs' % layer


@pytest.mark.parametrize('p,N', [(16, 5), (8, 7)])
def test_stem_wgrads_exact_on_integers(dmf, p, N):
    _, net = make_pair(p, 6, 8)
    h = net.trainer()
    # MS stem: the hi/lo-split 16-channel input tensor X0 = [hi0-3, lo0-3 | hi0-3, 0]; integers have lo = 0
    x = int_tensor((N, 4, p, p), -3, 3, 21)
    x16 = torch.zeros((N, 16, p, p), device=DEV)
    x16[:, 0:4] = x
    x16[:, 8:12] = x
    dz = int_tensor((N, 64, p, p), -2, 2, 22, density=0.5)
    # poison the patch slot behind the batch: 8x8 maps put 2 patches in one tile, and a batch that ends mid-tile must read zeros
    # there (the tensor maps are encoded for exactly N patches), not whatever the workspace held
    h.buffer('X0', torch.bfloat16, (N + 1, 2, p, p, 8), alias=True)[N].fill_(float('nan'))
    h.buffer('dZ', torch.bfloat16, (N + 1, 8, p, p, 8), alias=True)[N].fill_(float('nan'))
    h.buffer('X0', torch.bfloat16, (N, 2, p, p, 8), alias=True).copy_(to_c8(x16))
    h.buffer('dZ', torch.bfloat16, (N, 8, p, p, 8), alias=True).copy_(to_c8(dz))
    h.flat_grad.zero_()
    h.debug_op('wgrad', 'ms1', N)
    dw = torch.nn.grad.conv2d_weight(x.double(), net.ms1[0].weight.shape, dz.double(), padding=1)
    assert torch.equal(net.ms1[0].weight.grad.double(), dw)
    # PAN stem (CUDA cores, fp32 input)
    S = 4 * p
    xp = int_tensor((N, 1, S, S), -3, 3, 23)
    dz = int_tensor((N, 32, S, S), -2, 2, 24, density=0.5)
    h.buffer('in_pan', torch.float32, (N, 1, S, S), alias=True).copy_(xp)
    h.buffer('dZ', torch.bfloat16, (N, 4, S, S, 8), alias=True).copy_(to_c8(dz))
    h.flat_grad.zero_()
    h.debug_op('wgrad', 'pan1', N)
    dw = torch.nn.grad.conv2d_weight(xp.double(), net.pan1[0].weight.shape, dz.double(), padding=1)
    assert torch.equal(net.pan1[0].weight.grad.double(), dw)
    # PAN stem forward: integer weights too
    with torch.no_grad():
        net.pan1[0].weight.copy_(int_tensor(net.pan1[0].weight.shape, -4, 4, 25) / 4)
    h.debug_op('fwd', 'pan1', N)
    z = from_c8(h.buffer('Z_pan1', torch.bfloat16, (N, 4, S, S, 8)))
    want = F.conv2d(xp.double(), net.pan1[0].weight.detach().double(), None, padding=1)
    assert torch.equal(z.double(), want.float().bfloat16().double())


# ---------------------------------------------------------------------------------------------- loss and optimizer
@pytest.mark.parametrize('C,N', [(8, 1), (13, 300), (40, 77)])
def test_softmax_ce_matches_torch(dmf, C, N):
    g = torch.Generator(device=DEV).manual_seed(5)
    logits = torch.randn((N, C), device=DEV, generator=g) * 3
    tgt = torch.randint(0, C, (N,), device=DEV, generator=g)
    lr = logits.clone().requires_grad_(True)
    want = F.cross_entropy(lr, tgt)
    want.backward()
    for t in (tgt, tgt.float()):              # int64 after .long(), or the loaders' float labels
        loss, dl = dmf.softmax_ce(logits, t)
        assert abs(float(loss) - float(want.detach())) <= 1e-6 * max(1.0, abs(float(want.detach()))) * 4
        assert torch.allclose(dl, lr.grad, rtol=1e-5, atol=1e-8)


def test_fused_adam_matches_torch_adam(dmf):
    g = torch.Generator(device=DEV).manual_seed(6)
    shapes = [(64, 4, 3, 3), (64,), (12, 64), (7,)]
    flat = torch.randn(sum(int(np.prod(s)) for s in shapes), device=DEV, generator=g)
    flat_g = torch.zeros_like(flat)
    ps, qs, off = [], [], 0
    for s in shapes:
        n = int(np.prod(s))
        q = torch.nn.Parameter(flat[off:off + n].view(s))
        q.grad = flat_g[off:off + n].view(s)
        ps.append(q)
        qs.append(torch.nn.Parameter(q.detach().clone()))
        off += n
    mine = dmf.FusedAdam(ps, lr=1e-3)
    theirs = torch.optim.Adam(qs, lr=1e-3)
    for step in range(5):
        gr = torch.randn(flat.shape, device=DEV, generator=g) * (10.0 ** (step - 2))
        flat_g.copy_(gr)
        off = 0
        for q in qs:
            q.grad = gr[off:off + q.numel()].view(q.shape).clone()
            off += q.numel()
        mine.step()
        theirs.step()
        got = torch.cat([q.detach().reshape(-1) for q in ps])
        want = torch.cat([q.detach().reshape(-1) for q in qs])
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), (step, float((got - want).abs().max()))
    assert mine._flat is not None                         # the flat (single-launch) path was taken
    mine.zero_grad()
    assert float(flat_g.abs().max()) == 0.0 and ps[0].grad is not None


# ---------------------------------------------------------------------------------------------- whole step
def rnd(x):
    return x + (x.bfloat16().float() - x).detach()


def rnd_grad(x):
    if x.requires_grad:
        x.register_hook(lambda g: g.bfloat16().float())
    return x


def torch_step(ref, ms, pan, tgt, emulate):
    """loss.backward() of the oracle network; emulate=True rounds to bf16 wherever the kernels store bf16."""
    def block(name, x, pool):
        conv, bn = getattr(ref, name)[0], getattr(ref, name)[1]
        if not emulate:
            y = torch.relu(bn(conv(x)))
            return F.max_pool2d(y, 2) if pool else y
        w = conv.weight if name in ('ms1', 'pan1') else rnd(conv.weight)
        z32 = rnd_grad(F.conv2d(rnd_grad(x), w, None, padding=conv.padding))
        mean, var = z32.mean((0, 2, 3)), z32.var((0, 2, 3), unbiased=False)
        y = (rnd(z32) - mean.view(1, -1, 1, 1)) * (torch.rsqrt(var + bn.eps) * bn.weight).view(1, -1, 1, 1) + bn.bias.view(1, -1, 1, 1)
        y = torch.relu(y)
        if name == 'fuse':
            return y
        y = rnd(y)
        return F.max_pool2d(y, 2) if pool else y
    m = block('ms2', block('ms1', ms, False), True)
    q = block('pan3', block('pan2', block('pan1', pan, True), True), True)
    f = block('fuse', torch.cat([m, q], 1), False)
    logits = ref.fc2(torch.relu(ref.fc1(f.mean(dim=(2, 3)))))
    loss = F.cross_entropy(logits, tgt)
    ref.zero_grad()
    loss.backward()
    return loss.detach(), logits.detach()


def grads_of(net):
    return {n: (q.grad.detach().clone() if q.grad is not None else torch.zeros_like(q)) for n, q in net.named_parameters()}


@pytest.mark.parametrize('p,N', [(16, 64), (8, 33), (32, 16)])
def test_train_step_matches_torch(dmf, p, N):
    C = 12
    ref, net = make_pair(p, C, N)
    ms, pan, tgt = batch(p, C, N)
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    loss32, logits32 = torch_step(ref, ms, pan, tgt, emulate=False)
    g32, bufs32 = grads_of(ref), {k: v.clone() for k, v in ref.named_buffers()}
    ref.load_state_dict(sd0)
    loss_e, logits_e = torch_step(ref, ms, pan, tgt, emulate=True)
    ge = grads_of(ref)
    h = net.trainer()
    loss = h.step_patches(ms, pan, tgt)
    logits = h.buffer('logits', torch.float32, (N, C))
    gn = grads_of(net)
    # forward
    assert abs(float(loss) - float(loss_e)) <= 1e-4 * float(loss_e)
    assert abs(float(loss) - float(loss32)) <= 1e-3 * float(loss32)
    assert rel(logits, logits_e) <= 2e-3 and rel(logits, logits32) <= 1e-2
    # running statistics (torch semantics: momentum 0.1, unbiased variance, conv bias included in the mean)
    for k, v in net.named_buffers():
        if v.dtype.is_floating_point:
            assert rel(v, bufs32[k]) <= 5e-3, k
        else:
            assert int(v) == int(bufs32[k]) == 1, k
    # gradients: per-tensor table (relative L2 error vs the bf16-emulated torch step and vs pure fp32, cosine vs fp32) -> gpurun_out/
    import json
    import os
    table = {k: {'rel_err_vs_bf16_emulation': rel(gn[k], ge[k]), 'rel_err_vs_fp32': rel(gn[k], g32[k]),
                 'cosine_vs_fp32': float(F.cosine_similarity(gn[k].reshape(1, -1).double(), g32[k].reshape(1, -1).double())),
                 'numel': gn[k].numel()} for k in gn if not (k.endswith('.0.bias') and not k.startswith('fc'))}
    print(json.dumps({'p': p, 'N': N, 'gradients': table}))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out):
        json.dump({'test': 'tests/test_gpu_train.py::test_train_step_matches_torch', 'p': p, 'batch': N, 'bounds': {'rel_err_vs_bf16_emulation': GRAD_EMU,
                   'cosine_vs_fp32': GRAD_COS_FP32}, 'gradients': table}, open(os.path.join(out, 'grad_table_p%d.json' % p), 'w'), indent=1)
    for k in gn:
        if k.endswith('.0.bias') and not k.startswith('fc'):
            # a bias in front of a train-mode BatchNorm has zero gradient (torch returns rounding noise)
            assert float(gn[k].abs().max()) <= 1e-6 and float(g32[k].abs().max()) <= 1e-6, k
            continue
        assert rel(gn[k], ge[k]) <= GRAD_EMU, (k, rel(gn[k], ge[k]))
        cos = float(F.cosine_similarity(gn[k].reshape(1, -1).double(), g32[k].reshape(1, -1).double()))
        assert cos >= GRAD_COS_FP32, (k, cos)


def test_autograd_function_and_loop_track_fp32_training(dmf):
    """The reference's own loop text (zero_grad / criterion(model(a, b), t.long()) / backward / step) through the
    autograd.Function, and the fused train_step, both against fp32 torch training from the same initialisation."""
    p, C, N, steps = 16, 6, 64, 30
    ref, net = make_pair(p, C, N, seed=3)
    from model.gmfnet import Net
    net2 = Net(cfg(p, C, N))
    net2.load_state_dict(ref.state_dict())
    net2 = net2.to(DEV).train()
    # a learnable task: the class decides the mean level of both rasters
    g = torch.Generator(device=DEV).manual_seed(9)
    tgt = torch.randint(1, C, (N,), device=DEV, generator=g)
    lvl = (tgt.float() / C).view(N, 1, 1, 1)
    ms = (torch.rand((N, 4, p, p), device=DEV, generator=g) * 0.5 + lvl * 0.5)
    pan = (torch.rand((N, 1, 4 * p, 4 * p), device=DEV, generator=g) * 0.5 + lvl * 0.5)
    crit = torch.nn.CrossEntropyLoss()
    o_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    o_a = torch.optim.Adam(net.parameters(), lr=1e-3)                 # stock torch optimizer on the autograd path
    o_b = dmf.FusedAdam(net2.parameters(), lr=1e-3)                   # fused path
    l_ref, l_a, l_b = [], [], []
    for _ in range(steps):
        o_ref.zero_grad(); l = crit(ref(ms, pan), tgt.float().long()); l.backward(); o_ref.step(); l_ref.append(float(l))
        o_a.zero_grad(); l = crit(net(ms, pan), tgt.float().long()); l.backward(); o_a.step(); l_a.append(float(l))
        l_b.append(float(net2.train_step(ms, pan, tgt.float(), o_b)))
    assert abs(l_a[0] - l_ref[0]) <= 1e-3 * l_ref[0] and abs(l_b[0] - l_ref[0]) <= 1e-3 * l_ref[0]
    assert l_ref[-1] < 0.5 * l_ref[0], 'the fp32 run did not learn: the task is broken'
    for curve in (l_a, l_b):
        assert abs(curve[-1] - l_ref[-1]) <= 0.1 * l_ref[-1] + 0.02, (curve[-1], l_ref[-1])
        assert max(abs(a - b) for a, b in zip(curve, l_ref)) <= 0.15 * l_ref[0]
    # eval-mode inference of the trained module runs on the inference kernels with the UPDATED weights
    chk = RefNet(cfg(p, C, N)).to(DEV)
    chk.load_state_dict(net2.state_dict())
    chk.eval(); net2.eval()
    with torch.no_grad():
        want, got = chk(ms, pan), net2(ms, pan)
    assert torch.allclose(got, want, rtol=2e-2, atol=5e-3), float((got - want).abs().max())


def test_train_step_scene_equals_step_on_gathered_patches(dmf):
    """K1 fused in front of the step (and the IHS-product window as PAN input) = gather + step_patches."""
    from oracle import dmf_oracle as orc
    p, C, N = 16, 8, 48
    ms, pan, label = orc.synthetic_scene(40, 44, C - 1, seed=2, label_seed=3, blocky=True)
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc.set_labels(label)
    mspan = np.random.default_rng(4).random((sc.H4p, sc.W4p)).astype(np.float32)
    sc.set_mspan(mspan)
    idx = torch.from_numpy(np.random.default_rng(5).choice(40 * 44, N, replace=False)).to(DEV)
    for use_mspan in (False, True):
        ref, net = make_pair(p, C, N, seed=7)
        _, net_b = make_pair(p, C, N, seed=7)
        a, b, m, t = sc.gather(idx, tri=True)
        l1 = net.trainer().step_patches(a, m if use_mspan else b, t)
        l2 = net_b.trainer().step_scene(sc, idx, use_mspan=use_mspan)
        assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l1))
        ga, gb = net.trainer().flat_grad, net_b.trainer().flat_grad
        # not bit-equal: the batch statistics and weight gradients are summed with fp32/fp64 atomics, whose order differs
        # from run to run; a last-bit change of a statistic moves some bf16 activations by one ulp (see the module docstring)
        assert rel(gb, ga) <= 1e-2


def test_errors_are_loud(dmf):
    _, net = make_pair(16, 6, 8)
    h = net.trainer()
    ms, pan, tgt = batch(16, 6, 9)
    with pytest.raises(RuntimeError, match='batch must be'):
        h.step_patches(ms, pan, tgt)
    with pytest.raises(RuntimeError, match='no CPU'):
        net.cpu()(ms.cpu(), pan.cpu())


def test_data_parallel_gradient_parity_nccl():
    """2 ranks over NCCL: averaged flat gradient == single-rank gradient of the same two sub-batches (tools/dp_grad_check.py).
    Skipped on a box with fewer than 2 GPUs (the driver's GPU-test box has one; run under `gpurun --gpus 2`)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29611', os.path.join(repo, 'tools', 'dp_grad_check.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
