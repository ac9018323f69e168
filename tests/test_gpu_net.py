"""GPU parity of the network path (K3 + fused K1/K4) against the fp32 PyTorch oracle Net.

Tolerances (stated once, used below):
  * single tensor-core layer vs the same layer in fp32 on identical fp16 inputs and weights (the inference kernels run
    on fp16 operands, conv_tc.cuh): the two differ only in fp32 accumulation order and one fp16 rounding -> <= 2 fp16
    ulps (2^-10 relative each) + 1e-3 absolute;
  * logits vs the fp32 oracle: fp16 weights/activations, fp32 accumulate -> |dlogit| <= LOGIT_ATOL
    + LOGIT_RTOL * |logit| (per pixel: the largest |logit| of the row, see test_gpu_parity_fitted.py);
  * argmax agreement >= 99.9 % on the default-init network (BASELINE.json north_star), and on a
    sharpened network every pixel whose fp32 top-2 margin exceeds 2*LOGIT_ATOL must agree.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dmf_oracle as orc
from oracle.gmfnet_ref import Net as RefNet

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
LOGIT_ATOL, LOGIT_RTOL = 2e-3, 2e-2


@pytest.fixture(scope='module')
def dmf():
    import dmf as m
    return m


def cfg(p, C):
    return {'Categories_Number': C, 'patch_size': p, 'schedule': {'activate': 'Relu'}}


def make_ref(p, C, seed=3407, randomize_bn=True):
    torch.manual_seed(seed)
    net = RefNet(cfg(p, C)).eval()
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return net


def to_c8(x):
    """NCHW float -> [N][C/8][H][W][8] fp16 (the inference kernels' activation layout)."""
    N, C, H, W = x.shape
    return x.view(N, C // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.float16)


def from_c8(y):
    N, Cc, H, W, _ = y.shape
    return y.float().permute(0, 1, 4, 2, 3).reshape(N, Cc * 8, H, W)


def rb(x):
    """round to the inference storage format (fp16: activations and tensor-core weights)"""
    return x.to(torch.float16).float()


def ref_block(blk, x, pool, quant_w=True):
    conv, bn = blk[0], blk[1]
    w = rb(conv.weight) if quant_w else conv.weight
    y = F.conv2d(x, w, None, padding=conv.padding)
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    b = bn.bias + (conv.bias - bn.running_mean) * s
    y = rb(torch.relu(y * s[None, :, None, None] + b[None, :, None, None]))
    return F.max_pool2d(y, 2) if pool else y


def close_bf16(a, b, ulps=2):
    """within `ulps` fp16 ulps (2^-10 relative each, as an upper bound over the binade) + 1e-3"""
    tol = ulps * 2.0 ** -10 * torch.maximum(a.abs(), b.abs()) + 1e-3
    bad = (a - b).abs() > tol
    assert not bad.any(), 'mismatch: %d of %d, max abs err %g' % (int(bad.sum()), bad.numel(), float((a - b).abs().max()))


@pytest.mark.parametrize('p', [16, 8, 32])
def test_stems(dmf, p):
    net = make_ref(p, 8)
    h = dmf.NetHandle(p, 8, max_batch=64, device=DEV)
    h.load_state_dict(net.state_dict())
    g = torch.Generator().manual_seed(1)
    N = 5
    ms = torch.rand((N, 4, p, p), generator=g)
    pan = torch.rand((N, 1, 4 * p, 4 * p), generator=g)
    with torch.no_grad():
        want_ms = ref_block(net.ms1, ms, False, quant_w=False)
        want_pan = ref_block(net.pan1, pan, True, quant_w=False)
    got_ms = from_c8(h.debug_stem(0, ms.to(DEV), (N, 8, p, p, 8))).cpu()
    got_pan = from_c8(h.debug_stem(1, pan.to(DEV), (N, 4, 2 * p, 2 * p, 8))).cpu()
    close_bf16(got_ms, want_ms)
    close_bf16(got_pan, want_pan)


LAYERS = {0: ('ms2', 64, 128, 1, True), 1: ('pan2', 32, 64, 2, True), 2: ('pan3', 64, 128, 1, True),
          3: ('fuse', 256, 128, 0.5, False)}


@pytest.mark.parametrize('p', [16, 8, 32])
@pytest.mark.parametrize('layer', [0, 1, 2, 3])
def test_tensor_core_layer(dmf, p, layer):
    """One tcgen05 layer against (a) the fp32 torch layer on the same fp16 operands and (b) the
    CUDA-core debug convolution on the device."""
    name, cin, cout, smul, pool = LAYERS[layer]
    S = int(p * smul)
    net = make_ref(p, 8)
    h = dmf.NetHandle(p, 8, max_batch=64, device=DEV)
    h.load_state_dict(net.state_dict())
    g = torch.Generator().manual_seed(layer)
    N = 7                                            # odd: exercises the partially filled last tile
    x = rb(torch.rand((N, cin, S, S), generator=g) * 2 - 0.5)
    with torch.no_grad():
        want = ref_block(getattr(net, name), x, pool)
    So = S // 2 if pool else S
    och = 32 if layer in (0, 2) else cout // 8          # ms2 / pan3 write into the 256-channel concat buffer
    oc0 = 16 if layer == 2 else 0
    out_shape = (N, och, So, So, 8)
    got_tc = h.debug_layer(layer, 0, to_c8(x).to(DEV), out_shape)
    torch.cuda.synchronize()
    sl = slice(oc0, oc0 + cout // 8)
    close_bf16(from_c8(got_tc[:, sl]).cpu(), want)
    try:                                             # the CUDA-core debug conv cannot read row-pair packed weights
        got_dc = h.debug_layer(layer, 1, to_c8(x).to(DEV), out_shape)
        torch.cuda.synchronize()
        close_bf16(from_c8(got_dc[:, sl]).cpu(), want)
    except RuntimeError as e:
        assert 'row-pair' in str(e)
    if och != cout // 8:                             # chunks owned by the other branch stay untouched
        other = torch.ones(och, dtype=torch.bool)
        other[sl] = False
        assert float(got_tc[:, other].float().abs().max()) == 0.0


def run_both(dmf, p, C, net, ms_u16, pan_u16, idx, max_batch=256):
    W = ms_u16.shape[1]
    MS, PAN = orc.data_padding(ms_u16, p), orc.data_padding(pan_u16, p)
    a, b = orc.gather_dual(MS, PAN, idx // W, idx % W, p)
    with torch.no_grad():
        want = net(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    h = dmf.NetHandle(p, C, max_batch=max_batch, device=DEV)
    h.load_state_dict(net.state_dict())
    sc = dmf.Scene.from_raw(ms_u16, pan_u16, p, DEV)
    got_scene, _ = h.forward_scene(sc, flat_idx=idx)
    got_patch = h.forward_patches(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV))
    torch.cuda.synchronize()
    return want, got_scene.cpu().numpy(), got_patch.cpu().numpy()


@pytest.mark.parametrize('p', [16, 8, 32])
def test_logits_vs_fp32_oracle(dmf, p):
    C = 8
    net = make_ref(p, C)
    ms, pan, _ = orc.synthetic_scene(40, 44, 7, seed=5, blocky=True)
    idx = np.random.default_rng(0).integers(0, 40 * 44, 300 if p < 32 else 120)
    want, got_scene, got_patch = run_both(dmf, p, C, net, ms, pan, idx, max_batch=128)   # several chunks
    assert np.array_equal(got_scene, got_patch), 'scene-fused and patch-fed paths must be identical'
    err = np.abs(got_scene - want)
    tol = LOGIT_ATOL + LOGIT_RTOL * np.abs(want)
    assert (err <= tol).all(), 'max |dlogit| %g (tol %g)' % (err.max(), tol.flat[err.argmax()])
    assert (got_scene.argmax(1) == want.argmax(1)).mean() >= 0.999 or margin_rule(want, got_scene)


def margin_rule(want, got):
    top2 = np.sort(want, axis=1)[:, -2:]
    confident = (top2[:, 1] - top2[:, 0]) > 2 * (LOGIT_ATOL + LOGIT_RTOL * np.abs(top2[:, 1]))
    return bool((got.argmax(1)[confident] == want.argmax(1)[confident]).all())


def test_c1_whole_scene_matches_reference_run(dmf, golden):
    """BASELINE.json configs[0]: the reference's own Solver objects classified this synthetic scene
    with the default-init (seed 3407) oracle Net; the fused CUDA band path must reproduce its label
    map, confusion matrix and OA/AA/Kappa."""
    g = golden('solver_c1')
    H = W = 128
    p, C = 16, 8
    ms, pan, label = orc.synthetic_scene(H, W, 7, seed=0, label_seed=1)
    torch.manual_seed(3407)
    n = len(orc.split_data_old(label, [H, W, 4])[1][1])
    tr = int(0.02 * n)
    parts = torch.utils.data.random_split(range(n), [tr, n - 2 * tr, tr])       # replay the reference's RNG use
    next(iter(torch.utils.data.DataLoader(parts[0], batch_size=256, shuffle=True)))
    net = RefNet(cfg(p, C)).eval()
    h = dmf.NetHandle(p, C, max_batch=4096, device=DEV)
    h.load_state_dict(net.state_dict())
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc.set_labels(label)
    logits, _ = h.forward_scene(sc, first=0, count=512)
    np.testing.assert_allclose(logits.cpu().numpy(), g['logits_first512'], rtol=LOGIT_RTOL, atol=LOGIT_ATOL)
    pred_map, cm = h.infer_scene(sc)
    pm = pred_map.cpu().numpy()
    agree = (pm == g['label_map']).mean()
    assert agree >= 0.999, agree
    M = cm.cpu().numpy().astype(np.float64)
    assert M.sum() == H * W
    # bit-exact confusion matrix given identical predictions
    assert np.array_equal(M, orc.confusion(pm.reshape(-1), label.reshape(-1), C))
    if agree == 1.0:
        assert np.array_equal(M, g['M'])
        aa, oa, k, _ = orc.aa_oa(M)
        assert np.array_equal(np.array([aa, oa, k]), g['aa_oa_k'], equal_nan=True)
    # row-band sharding: two half-scene calls add up to the whole-scene matrix (integer sums)
    cm2 = torch.zeros_like(cm)
    pm2 = torch.zeros_like(pred_map)
    h.infer_scene(sc, 0, 61, pred_map=pm2, cm=cm2)
    h.infer_scene(sc, 61, H, pred_map=pm2, cm=cm2)
    assert torch.equal(cm2, cm) and torch.equal(pm2, pred_map)


def test_sharpened_net_margin_rule(dmf):
    """A network whose predictions actually vary (default init predicts one class everywhere):
    centre the logits over the scene so that the argmax is decided by the per-pixel signal."""
    p, C = 16, 12
    net = make_ref(p, C, seed=7)
    ms, pan, _ = orc.synthetic_scene(48, 48, 11, seed=9, blocky=True)
    idx = np.arange(48 * 48)[::3]
    MS, PAN = orc.data_padding(ms, p), orc.data_padding(pan, p)
    a, b = orc.gather_dual(MS, PAN, idx // 48, idx % 48, p)
    with torch.no_grad():
        base = net(torch.from_numpy(a), torch.from_numpy(b))
        net.fc2.bias.data -= base.mean(0)
    want, got, _ = run_both(dmf, p, C, net, ms, pan, idx)
    assert len(np.unique(want.argmax(1))) >= 3
    assert margin_rule(want, got)
    print('sharpened-net argmax agreement: %.4f' % (got.argmax(1) == want.argmax(1)).mean())


@pytest.mark.parametrize('H,W,ncls', [(1000, 1000, 12), (2001, 2101, 11)])
def test_full_size_scene_properties(dmf, H, W, ncls):
    """BASELINE.json configs[1] / [2] at full size (1 M and 4.2 M pixels): size-independent properties of
    the fused band path — the matrix covers every pixel, equals the histogram of (label map, labels) bit
    for bit, row bands add up to the whole scene, and a random sample of pixels agrees with the fp32 oracle."""
    p, C = 16, ncls + 1
    ms, pan, label = orc.synthetic_scene(H, W, ncls, seed=0, label_seed=1)
    net = make_ref(p, C, seed=3407, randomize_bn=False)
    h = dmf.NetHandle(p, C, device=DEV)
    h.load_state_dict(net.state_dict())
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc.set_labels(label)
    pred_map, cm = h.infer_scene(sc)
    assert int(cm.sum()) == H * W
    lab = torch.from_numpy(label).to(DEV).long()
    want = torch.bincount((pred_map.long() * C + lab).reshape(-1), minlength=C * C).reshape(C, C)
    assert torch.equal(cm, want)
    cm2, pm2 = torch.zeros_like(cm), torch.zeros_like(pred_map)
    cut = [0, H // 3, H // 3 + 1, H]                  # uneven bands, one of them a single row
    for a, b in zip(cut[:-1], cut[1:]):
        h.infer_scene(sc, a, b, pred_map=pm2, cm=cm2)
    assert torch.equal(cm2, cm) and torch.equal(pm2, pred_map)
    idx = np.random.default_rng(1).integers(0, H * W, 400)
    idx[:4] = [0, W - 1, (H - 1) * W, H * W - 1]      # scene corners: windows made of reflect padding
    logits, pred = h.forward_scene(sc, flat_idx=idx, want_pred=True)
    assert torch.equal(pred.cpu(), pred_map.reshape(-1)[torch.from_numpy(idx).to(DEV)].cpu())
    MS = orc.data_padding(ms, p)
    PAN = orc.data_padding(pan, p)
    a, b = orc.gather_dual(MS, PAN, idx // W, idx % W, p)
    with torch.no_grad():
        ref = net(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    got = logits.cpu().numpy()
    assert (np.abs(got - ref) <= LOGIT_ATOL + LOGIT_RTOL * np.abs(ref)).all()
    assert (got.argmax(1) == ref.argmax(1)).mean() >= 0.999
    aa, oa, k, _ = orc.aa_oa(cm.cpu().numpy().astype(np.float64))
    assert 0.0 <= oa <= 1.0 and np.isfinite(k)


def test_bad_arguments_raise_not_crash(dmf):
    with pytest.raises(RuntimeError, match='patch_size'):
        dmf.NetHandle(12, 8, device=DEV)
    h = dmf.NetHandle(16, 8, max_batch=64, device=DEV)
    with pytest.raises(RuntimeError, match='finalize'):
        h.forward_patches(torch.zeros(1, 4, 16, 16, device=DEV), torch.zeros(1, 1, 64, 64, device=DEV))
    net = make_ref(16, 8)
    sd = net.state_dict()
    sd.pop('fc2.bias')
    with pytest.raises(RuntimeError, match='fc2.bias'):
        h.load_state_dict(sd)
    h.load_state_dict(net.state_dict())
    ms, pan, label = orc.synthetic_scene(20, 20, 7, seed=1)
    sc8 = dmf.Scene.from_raw(ms, pan, 8, DEV)
    with pytest.raises(RuntimeError, match='patch size'):
        h.forward_scene(sc8, first=0, count=4)
    sc = dmf.Scene.from_raw(ms, pan, 16, DEV)
    with pytest.raises(RuntimeError, match='labels'):
        h.infer_scene(sc, cm=torch.zeros((8, 8), dtype=torch.int64, device=DEV))
    with pytest.raises(RuntimeError, match='outside'):
        h.forward_scene(sc, first=390, count=20)
    with pytest.raises(RuntimeError, match='tri mode'):
        sc.gather([0, 1], tri=True)
