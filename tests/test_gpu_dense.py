"""GPU parity of the scene-dense whole-scene path (csrc/dense.cu, csrc/dense_tc.cuh).

The dense path evaluates every layer once per scene position and border class instead of once per patch
(Solver.color()/test() visit patches at stride 1, solver/mainsolver.py:167-185).  Checks:
  * layer by layer, in isolation: for sampled anchors the per-patch tensor gathered from the dense maps of layer L-1
    is pushed through the fp32 torch definition of layer L (same fp16 rounding points) and compared with the tensor
    gathered from the dense maps of layer L: <= 2 fp16 ulps + 1e-3, the tolerance of the single-layer tests in
    test_gpu_net.py;
  * whole scene, several bands: logits vs the per-patch kernels (fp32 summation-order noise only) and vs the fp32
    oracle (|d| <= LOGIT_ATOL + LOGIT_RTOL*|logit|), argmax agreement >= 99.9 %, label map == argmax of the returned
    logits, confusion matrix == oracle.confusion(pred, label) bit for bit.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dmf_oracle as orc
from test_gpu_net import LOGIT_ATOL, LOGIT_RTOL, cfg, make_ref, rb, ref_block

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False          # the torch reference layers run on the GPU here: keep them fp32
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope='module')
def dmf():
    import dmf as m
    return m


def c3(i, S):
    return 0 if i == 0 else (2 if i == S - 1 else 1)


def gather_patches(buf, planes_chunks, dims, anchors, S, a, b, chunk0=0, chunks=None):
    """Per-patch tensors [N][C][S][S] (float) out of a dense map [9][planes_chunks][rows][cols][8]: patch-relative position
    (i, j) of the patch anchored at band-local (xl, y) sits at (a*xl + b*i, a*y + b*j) of plane (c3(i), c3(j))."""
    rows, cols = dims
    chunks = planes_chunks if chunks is None else chunks
    m = buf[:9 * planes_chunks * rows * cols * 8].view(9, planes_chunks, rows, cols, 8)
    ii = torch.arange(S, device=buf.device)
    cls = torch.tensor([c3(i, S) for i in range(S)], device=buf.device)
    plane = cls[:, None] * 3 + cls[None, :]                                   # [S][S]
    out = []
    for xl, y in anchors:
        r = (a * xl + b * ii)[:, None].expand(S, S)
        c = (a * y + b * ii)[None, :].expand(S, S)
        v = m[plane, chunk0:chunk0 + chunks, r, c]                            # [S][S][chunks][8]
        out.append(v.permute(2, 3, 0, 1).reshape(chunks * 8, S, S))
    return torch.stack(out).float()


def gather_patches_b1(buf, dims, anchors, p):
    """The PAN stem maps are phase-separated ([variant][row phase * 2 + col phase][chunk][rows][cols][8]): pooled-once position
    (U, V) = (2*xl + u, 2*y + v) sits at (U >> 1, V >> 1) of phase plane (U & 1, V & 1)."""
    rows, cols = dims
    S = 2 * p
    m = buf[:9 * 4 * 4 * rows * cols * 8].view(9, 4, 4, rows, cols, 8)
    ii = torch.arange(S, device=buf.device)
    cls = torch.tensor([c3(i, S) for i in range(S)], device=buf.device)
    variant = cls[:, None] * 3 + cls[None, :]
    out = []
    for xl, y in anchors:
        r = (2 * xl + ii)[:, None].expand(S, S)
        c = (2 * y + ii)[None, :].expand(S, S)
        v = m[variant, (r & 1) * 2 + (c & 1), :, r >> 1, c >> 1]               # [S][S][4][8]
        out.append(v.permute(2, 3, 0, 1).reshape(32, S, S))
    return torch.stack(out).float()


def assert_close_bf16(got, want, what, ulps=2, max_bad_frac=0.0):
    tol = ulps * 2.0 ** -10 * torch.maximum(got.abs(), want.abs()) + 1e-3          # fp16 ulps
    bad = (got - want).abs() > tol
    frac = float(bad.float().mean())
    assert frac <= max_bad_frac, '%s: %d of %d outside tolerance (max abs err %g, max |want| %g)' % (
        what, int(bad.sum()), bad.numel(), float((got - want).abs().max()), float(want.abs().max()))


def scene_and_net(dmf, p, H, W, C, seed=0):
    ms, pan, label = orc.synthetic_scene(H, W, C - 1, seed=seed, label_seed=seed + 1)
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc.set_labels(label)
    ref = make_ref(p, C)
    with torch.no_grad():                       # negative BatchNorm scales: the dense kernels fold sign(scale) into the weights
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight[::3] *= -1.0
    h = dmf.NetHandle(p, C, max_batch=2048, device=DEV)
    h.load_state_dict(ref.state_dict())
    return ms, pan, label, sc, ref, h


@pytest.mark.parametrize('p,H,W,row0,nb', [(16, 21, 37, 2, 19), (8, 30, 41, 0, 30), (32, 9, 20, 1, 7)])
def test_dense_layers_in_isolation(dmf, p, H, W, row0, nb):
    C = 8
    ms, pan, label, sc, ref, h = scene_and_net(dmf, p, H, W, C)
    h.set_dense(True, band_rows=nb)
    h.infer_scene(sc, row0, row0 + nb)
    torch.cuda.synchronize()
    ref = ref.to(DEV)
    g = np.random.default_rng(5)
    anchors = [(0, 0), (nb - 1, W - 1), (0, W - 1), (nb - 1, 0)] + [(int(g.integers(nb)), int(g.integers(W))) for _ in range(12)]
    idx = torch.tensor([(row0 + xl) * W + y for xl, y in anchors], device=DEV)
    pm, pp, _ = sc.gather(idx, want_target=False)
    A, dims = h.dense_buffer('A')
    B1, _ = h.dense_buffer('B1')
    B2, _ = h.dense_buffer('B2')
    CAT, _ = h.dense_buffer('CAT')
    Sm, _ = h.dense_buffer('S')
    R, Cc = dims
    with torch.no_grad():
        # stems: fp32 weights (CUDA cores), fp16-rounded output
        got = gather_patches(A, 8, dims, anchors, p, 1, 1)
        assert_close_bf16(got, ref_block(ref.ms1, pm, False, quant_w=False), 'ms stem maps')
        got_b1 = gather_patches_b1(B1, dims, anchors, p)
        assert_close_bf16(got_b1, ref_block(ref.pan1, pp, True, quant_w=False), 'pan stem maps')
        # tensor-core layers, each fed with the dense path's own input
        got_ms2 = gather_patches(CAT, 32, dims, anchors, p // 2, 1, 2, 0, 16)
        assert_close_bf16(got_ms2, ref_block(ref.ms2, got, True), 'ms2 (conv_pool4, stride-1 pool)')
        got_b2 = gather_patches(B2, 8, dims, anchors, p, 1, 1)
        assert_close_bf16(got_b2, ref_block(ref.pan2, got_b1, True), 'pan2 (conv_pool4: conv + aligned 2x2 max fused)')
        got_p3 = gather_patches(CAT, 32, dims, anchors, p // 2, 1, 2, 16, 16)
        assert_close_bf16(got_p3, ref_block(ref.pan3, got_b2, True), 'pan3 (conv_pool4, stride-1 pool)')
        # fusion conv + row sums (fuse_rowsum_kernel): F stays on chip; S[a][ch][X][y][8] = sum_l F[a, cls(l)][X][y + 2l] in fp32.
        # The patch sum of the reference fusion output must equal sum_k S[cls(k)][xl + 2k][y].
        want_f = ref_block(ref.fuse, torch.cat([got_ms2, got_p3], 1), False)           # [N][128][P2][P2], fp16-rounded
        P2, rows_b = p // 2, nb + p - 1
        S5 = Sm[:3 * 16 * rows_b * W * 8].view(3, 16, rows_b, W, 8)
        for i, (xl, y) in enumerate(anchors):
            got_sum = P2 * sum(S5[c3(k, P2), :, xl + 2 * k, y, :].float() for k in range(P2)).reshape(128)      # S holds row MEANS (fp16)
            want_sum = want_f[i].sum(dim=(1, 2))
            tol = (2 * 2.0 ** -8 * want_f[i].abs() + 1e-3).sum(dim=(1, 2))
            assert bool(((got_sum - want_sum).abs() <= tol).all()), 'fuse + row sums: max err %g' % float((got_sum - want_sum).abs().max())
    h.close()


@pytest.mark.parametrize('p,H,W,band', [(16, 40, 52, 16), (8, 33, 70, 128), (32, 12, 40, 5)])
def test_dense_scene_matches_patch_path_and_oracle(dmf, p, H, W, band):
    C = 8
    ms, pan, label, sc, ref, h = scene_and_net(dmf, p, H, W, C, seed=3)
    h.set_dense(True, band_rows=band)
    pm_d, cm_d, lg_d = h.infer_scene(sc, want_logits=True)
    h.set_dense(False)
    pm_p, cm_p, lg_p = h.infer_scene(sc, want_logits=True)
    torch.cuda.synchronize()
    # dense vs per-patch kernels: same fp16 rounding points, different fp32 summation order and stem arithmetic
    d = (lg_d - lg_p).abs()
    assert float(d.max()) <= LOGIT_ATOL + LOGIT_RTOL * float(lg_p.abs().max()), 'dense vs per-patch logits: max |d| = %g' % float(d.max())
    # label map = first-maximum argmax of the logits it returned; matrix = confusion of that map
    pred = lg_d.max(1)[1].to(torch.uint8).view(H, W)
    assert torch.equal(pred, pm_d)
    want_cm = orc.confusion(pm_d.cpu().numpy().reshape(-1), label.reshape(-1), C)
    assert np.array_equal(cm_d.cpu().numpy().astype(np.float64), want_cm)
    agree = float((pm_d == pm_p).float().mean())
    assert agree >= 0.999, 'dense vs per-patch argmax agreement %.5f' % agree
    if torch.equal(pm_d, pm_p):
        assert torch.equal(cm_d, cm_p)
    # vs the fp32 oracle on a sample of pixels
    g = np.random.default_rng(11)
    idx = torch.from_numpy(g.choice(H * W, size=min(H * W, 600), replace=False)).to(DEV)
    a, b, _ = sc.gather(idx, want_target=False)
    with torch.no_grad():
        want = ref.to(DEV)(a, b)
    got = lg_d[idx]
    tol = LOGIT_ATOL + LOGIT_RTOL * want.abs()
    assert bool(((got - want).abs() <= tol).all()), 'dense vs fp32 oracle: max |d| = %g' % float((got - want).abs().max())
    margin = want.topk(2, dim=1)[0]
    sure = (margin[:, 0] - margin[:, 1]) > 2 * (LOGIT_ATOL + LOGIT_RTOL * want.abs().max(1)[0])
    assert bool((got.max(1)[1] == want.max(1)[1])[sure].all())
    h.close()


def test_dense_row_bands_add_up(dmf):
    """Two band calls (what two ranks do) == one whole-scene call, bit for bit."""
    p, H, W, C = 16, 37, 45, 6
    ms, pan, label, sc, ref, h = scene_and_net(dmf, p, H, W, C, seed=7)
    h.set_dense(True, band_rows=16)
    pm, cm = h.infer_scene(sc)
    pm2 = torch.zeros_like(pm)
    cm2 = torch.zeros_like(cm)
    h.infer_scene(sc, 0, 19, pred_map=pm2, cm=cm2)
    h.infer_scene(sc, 19, H, pred_map=pm2, cm=cm2)
    torch.cuda.synchronize()
    assert torch.equal(pm, pm2) and torch.equal(cm, cm2)
    assert int(cm.sum()) == H * W
    h.close()


@pytest.mark.parametrize('H,world', [(45, 3), (20, 2), (33, 1)])
def test_band_scene_equals_whole_scene(dmf, H, world):
    """A rank that uploads only its band (+ p-1 halo rows) and normalises with the all-reduced scene range sees the same
    windows and produces the same label rows / matrix as the whole-scene object (bench.py e2e arm at N > 1)."""
    from solver.mainsolver import row_band
    p, W, C = 16, 30, 6
    ms, pan, label = orc.synthetic_scene(H, W, C - 1, seed=9, label_seed=10)
    whole = dmf.Scene.from_raw(ms, pan, p, DEV)
    whole.set_labels(label)
    ref = make_ref(p, C)
    h = dmf.NetHandle(p, C, max_batch=1024, device=DEV)
    h.load_state_dict(ref.state_dict())
    pm_w, cm_w = h.infer_scene(whole)
    ms_t = torch.from_numpy(ms.view(np.int16)).to(DEV)
    pan_t = torch.from_numpy(pan.view(np.int16)).to(DEV)
    bands = [row_band(H, r, world) for r in range(world)]
    # what the MIN all-reduce of {lo, -hi} yields
    rm = torch.stack([dmf.raster_minmax(ms_t[a:b].contiguous()) for a, b in bands])
    rp = torch.stack([dmf.raster_minmax(pan_t[4 * a:4 * b].contiguous()) for a, b in bands])
    ms_rng = torch.stack([rm[:, 0].min(), rm[:, 1].max()])
    pan_rng = torch.stack([rp[:, 0].min(), rp[:, 1].max()])
    assert float(ms_rng[0]) == ms.min() and float(ms_rng[1]) == ms.max() and float(pan_rng[1]) == pan.max()
    cm_sum = torch.zeros_like(cm_w)
    for r0, r1 in bands:
        s0, s1 = dmf.band_slice(H, p, r0, r1)
        a, b = ms_t[s0:s1].contiguous(), pan_t[4 * s0:4 * s1].contiguous()
        band = dmf.Scene.from_raw(a, b, p, DEV)
        band.update_raw(a, b, ms_rng, pan_rng)
        band.set_labels(label[s0:s1])
        idx_w = torch.arange(r0 * W, r1 * W, device=DEV)
        idx_b = idx_w - s0 * W
        gw, gb = whole.gather(idx_w, want_target=False), band.gather(idx_b, want_target=False)
        assert torch.equal(gw[0], gb[0]) and torch.equal(gw[1], gb[1]), 'band windows differ from the whole scene'
        pm_b, cm_b = h.infer_scene(band, r0 - s0, r1 - s0)
        assert torch.equal(pm_b[r0 - s0:r1 - s0], pm_w[r0:r1])
        cm_sum += cm_b
    assert torch.equal(cm_sum, cm_w)
    h.close()


@pytest.mark.parametrize('p,H,W', [(8, 1, 1), (16, 3, 200), (16, 130, 5)])
def test_dense_degenerate_scene_shapes(dmf, p, H, W):
    """Scenes smaller than a tile / a head segment, single rows and columns: same label map as the per-patch kernels."""
    C = 5
    ms, pan, label, sc, ref, h = scene_and_net(dmf, p, H, W, C, seed=21)
    h.set_dense(True, band_rows=64)
    pm_d, cm_d, lg_d = h.infer_scene(sc, want_logits=True)
    h.set_dense(False)
    pm_p, cm_p, lg_p = h.infer_scene(sc, want_logits=True)
    torch.cuda.synchronize()
    assert int(cm_d.sum()) == H * W
    assert float((lg_d - lg_p).abs().max()) <= LOGIT_ATOL + LOGIT_RTOL * float(lg_p.abs().max())
    assert float((pm_d == pm_p).float().mean()) >= 0.999
    h.close()


@pytest.mark.parametrize('C', [2, 37, 64])
def test_dense_class_counts(dmf, C):
    """Class counts from the minimum to the maximum the C-ABI accepts (the head's class loop, logits layout, histogram size)."""
    p, H, W = 8, 20, 150
    ms, pan, label, sc, ref, h = scene_and_net(dmf, p, H, W, C, seed=31)
    h.set_dense(True, band_rows=64)
    pm_d, cm_d, lg_d = h.infer_scene(sc, want_logits=True)
    h.set_dense(False)
    pm_p, cm_p, lg_p = h.infer_scene(sc, want_logits=True)
    torch.cuda.synchronize()
    assert int(cm_d.sum()) == H * W and int(pm_d.max()) < C
    assert float((lg_d - lg_p).abs().max()) <= LOGIT_ATOL + LOGIT_RTOL * float(lg_p.abs().max())
    assert torch.equal(lg_d.max(1)[1].to(torch.uint8).view(H, W), pm_d)
    assert np.array_equal(cm_d.cpu().numpy().astype(np.float64), orc.confusion(pm_d.cpu().numpy().reshape(-1), label.reshape(-1), C))
    h.close()


def test_ihs_product_as_pan_input(dmf):
    """set_pan_source(True): scene inference reads the scene's IHS product (dataset_tri's third raster, train/dataset.py:259-279)
    in place of PAN — same logits as a scene whose PAN raster IS that product, on the dense path and on the per-patch kernels."""
    p, H, W, C = 16, 24, 40, 6
    ms, pan, label, sc, ref, h = scene_and_net(dmf, p, H, W, C, seed=41)
    g = np.random.default_rng(5)
    mspan_pad = g.random((4 * H + 4 * p - 1, 4 * W + 4 * p - 1)).astype(np.float32)      # stands for data_padding(IHS_tran(...))
    sc.set_mspan(mspan_pad)
    ms_pad = sc.export(0).cpu().numpy()
    twin = dmf.Scene.from_padded(ms_pad, mspan_pad, p, DEV)                              # PAN := the product
    twin.set_labels(label)
    for dense in (True, False):
        h.set_dense(dense, band_rows=16)
        h.set_pan_source(True)
        pm_a, cm_a, lg_a = h.infer_scene(sc, want_logits=True)
        h.set_pan_source(False)
        pm_b, cm_b, lg_b = h.infer_scene(twin, want_logits=True)
        pm_c, _, lg_c = h.infer_scene(sc, want_logits=True)                              # plain PAN input: must differ
        torch.cuda.synchronize()
        assert torch.equal(lg_a, lg_b) and torch.equal(pm_a, pm_b) and torch.equal(cm_a, cm_b)
        assert not torch.equal(lg_a, lg_c)
    bare = dmf.Scene.from_raw(ms, pan, p, DEV)
    h.set_pan_source(True)
    with pytest.raises(RuntimeError, match='IHS product'):
        h.infer_scene(bare)
    h.close()


_SHARE_PROBE = r'''
import os, sys
import numpy as np
import torch
sys.path[:0] = [%(tests)r, %(repo)r, %(pkg)r]
import dmf
import test_gpu_dense as T
out = {}
for p, H, W, band in ((16, 45, 61, 24), (8, 40, 33, 64), (32, 20, 37, 9)):
    ms, pan, label, sc, ref, h = T.scene_and_net(dmf, p, H, W, 8, seed=5)
    h.set_dense(True, band_rows=band)
    pm, cm, lg = h.infer_scene(sc, want_logits=True)
    torch.cuda.synchronize()
    out['pm%%d' %% p], out['cm%%d' %% p], out['lg%%d' %% p] = pm.cpu().numpy(), cm.cpu().numpy(), lg.cpu().numpy()
    h.close()
np.savez(sys.argv[1], **out)
'''


def test_shared_sub_positions_equal_the_unshared_evaluation(dmf, tmp_path):
    """The stride-1 conv + pool layers let interior-class cells take sub-position 1 from their neighbour (dense_tc.cuh, SHARE);
    DMF_DENSE_SHARE=0 makes every cell evaluate all four itself.  Both are the same conv outputs (other fp32 summation order of the
    column taps at most), so the label maps must agree and the logits differ by rounding only.  The switch is read once per process."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    code = _SHARE_PROBE % {'tests': here, 'repo': repo, 'pkg': os.path.join(repo, 'dual-modal-fusion_b200')}
    res = {}
    for tag, val in (('shared', '1'), ('unshared', '0')):
        path = str(tmp_path / (tag + '.npz'))
        env = dict(os.environ, DMF_DENSE_SHARE=val)
        r = subprocess.run([sys.executable, '-c', code, path], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[tag] = np.load(path)
    for p in (16, 8, 32):
        a, b = res['shared']['lg%d' % p], res['unshared']['lg%d' % p]
        tol = LOGIT_ATOL + LOGIT_RTOL * np.abs(b).max()
        assert np.abs(a - b).max() <= 0.1 * tol, 'p=%d: shared vs unshared logits differ by %g (rounding-level bound %g)' % (p, np.abs(a - b).max(), 0.1 * tol)
        pa, pb = res['shared']['pm%d' % p], res['unshared']['pm%d' % p]
        assert (pa == pb).mean() >= 0.999, 'p=%d: label maps agree on %.5f' % (p, (pa == pb).mean())
        if np.array_equal(pa, pb):
            assert np.array_equal(res['shared']['cm%d' % p], res['unshared']['cm%d' % p])
