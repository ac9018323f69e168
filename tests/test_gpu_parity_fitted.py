"""Non-degenerate parity of the network path: argmax agreement and confusion matrices on networks whose predictions VARY.

The default-initialised GMFNet predicts one class for every pixel (Kappa = 0), so the ">= 99.9 % argmax agreement" of
BASELINE.json's north_star is trivially met on it.  These tests use the fitted nets of oracle/fitted_net.py (seed-3407
convolutions, calibrated BatchNorm, head fitted on the structured synthetic scene: >= 5 predicted classes, Kappa > 0.1):

  * C1 (128 x 128, all 16 384 pixels): the product path against the run of the REFERENCE's own Solver / DataLoader /
    dataset_dual objects with the same fp32 network on the CPU (tests/golden/solver_c1_fitted.npz, produced by
    tests/golden/make_golden.py) — solver/mainsolver.py:139-141, 167-185; train/test.py:58-60;
  * C2 (1000 x 1000) and C3 (2001 x 2101): >= 20 000 sampled pixels against the fp32 oracle Net evaluated by torch on the GPU
    with TF32 off, on patches cut by the (bit-exact) K1 gather.

Logit tolerance: bf16 operands (8-bit mantissa) through five convolution layers with fp32 accumulation leave every pooled
feature with an error of a few 1e-3 of the FEATURE scale, and a logit is a sum over those features: its error scales with the
magnitude of the pixel's logit vector, not with the individual logit (a logit near 0 next to one of 35 is off by the same
~0.3).  So |dlogit| <= LOGIT_ATOL + LOGIT_RTOL * max_c |logit[c]| per pixel (the same constants as test_gpu_net.py, where the
default-initialised net has logits of one magnitude).  Measured on the fitted nets: max |dlogit| 0.5 at logits up to 35.

Asserted: argmax agreement >= 0.999; every disagreeing pixel has an fp32 top-2 margin below twice that tolerance; the confusion matrix equals
oracle.confusion(pred_map, label) bit for bit, and the reference run's matrix / OA / AA / Kappa when the predictions agree
everywhere.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import dmf_oracle as orc
from oracle import fitted_net
from test_gpu_net import LOGIT_ATOL, LOGIT_RTOL

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def product_net(tag, **b200):
    from model.gmfnet import Net
    cfg = dict(fitted_net.cfg_for(tag), b200=b200)
    net = Net(cfg)
    net.load_state_dict(fitted_net.fitted_state(tag))
    return net.to(DEV).eval()


def logit_tol(want, extra_rtol=0.0):
    """per-pixel tolerance, broadcast over the classes"""
    return (LOGIT_ATOL + (LOGIT_RTOL + extra_rtol) * np.abs(want).max(axis=1))[:, None]


def check_disagreements(pm_flat, want_logits, what):
    """pm_flat: predictions of the product path, want_logits: fp32 logits of the same pixels.  Returns the agreement."""
    want = want_logits.argmax(1)
    agree = float((pm_flat == want).mean())
    srt = np.sort(want_logits, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    bad = np.flatnonzero(pm_flat != want)
    tol = logit_tol(want_logits)[:, 0]
    assert np.all(margin[bad] <= 2 * tol[bad]), '%s: a pixel with a clear fp32 margin (%.4g) was classified differently' % (
        what, float(margin[bad].max()))
    assert agree >= 0.999, '%s: argmax agreement %.5f < 0.999 (%d of %d differ)' % (what, agree, bad.size, want.size)
    return agree, margin


def test_c1_fitted_whole_scene_matches_reference_run(golden):
    import dmf
    g = golden('solver_c1_fitted')
    H = W = 128
    C = 8
    ms, pan, label = fitted_net.scene('c1')
    assert len(np.unique(g['label_map'])) >= 5 and g['aa_oa_k'][2] > 0.1          # the reference run itself is non-degenerate
    net = product_net('c1')
    scene = dmf.Scene.from_raw(ms, pan, 16, DEV)
    scene.set_labels(label)
    want_logits = g['logits'].astype(np.float32)
    for dense in (True, False):
        h = net.native()
        h.set_dense(dense)
        pred_map, cm, logits = h.infer_scene(scene, want_logits=True)
        pm, M, lg = pred_map.cpu().numpy(), cm.cpu().numpy().astype(np.float64), logits.cpu().numpy()
        what = 'C1 fitted, %s path' % ('dense' if dense else 'per-patch')
        # logits: exact fp32 goldens for the first 512 pixels, a float16 copy for the rest (2^-11 relative on top of the tolerance)
        agree = float((pm == g['label_map']).mean())
        bad = np.flatnonzero(pm.reshape(-1) != g['label_map'].reshape(-1))
        err = np.abs(lg - want_logits)
        print('%s: max |dlogit| %.4f at max |logit| %.2f; agreement %.5f (%d px differ)' % (what, err.max(), np.abs(want_logits).max(), agree, bad.size))
        err512 = np.abs(lg[:512] - g['logits_first512'])
        assert np.all(err512 <= logit_tol(g['logits_first512'])), '%s: logits off by %g' % (what, err512.max())
        assert np.all(err <= logit_tol(want_logits, 2.0 ** -10)), '%s: logits off by %g' % (what, err.max())
        assert agree >= 0.999, '%s: argmax agreement with the reference run %.5f' % (what, agree)
        assert np.all(g['margin_top2'][bad] <= 2 * logit_tol(want_logits[bad])[:, 0]) if bad.size else True
        assert np.array_equal(pm.reshape(-1), lg.argmax(1)), what + ': label map is not the first-max argmax of the logits'
        assert np.array_equal(M, orc.confusion(pm.reshape(-1), label.reshape(-1), C)), what + ': confusion matrix'
        if bad.size == 0:
            assert np.array_equal(M, g['M'])
            aa, oa, k, _ = orc.aa_oa(M)
            assert np.array_equal(np.array([aa, oa, k]), g['aa_oa_k'])
        else:                                        # each differing pixel moves one count between two rows of its column
            assert np.abs(M - g['M']).sum() == 2 * bad.size
        # Solver.test()'s sample set out of the same maps
        Mt = orc.confusion(pm.reshape(-1)[g['test_idx']], label.reshape(-1)[g['test_idx']], C)
        if bad.size == 0:
            assert np.array_equal(Mt, g['M_test'])
        print('%s: agreement %.5f, %d classes predicted, Kappa %.4f' % (what, agree, len(np.unique(pm)), orc.aa_oa(M)[2]))


@pytest.mark.parametrize('tag', ['c2', 'c3', 'p8', 'p32'])
def test_full_scale_fitted_sampled_pixels_vs_fp32_oracle(tag):
    """>= 20 000 sampled pixels of the C2 / C3 scene: whole-scene dense inference vs the fp32 oracle on the same patches.
    p8 / p32: the other patch sizes of BASELINE.json configs[4], small scenes, EVERY pixel."""
    import dmf
    H, W, ncls, p = fitted_net.WORKLOADS[tag]
    C = ncls + 1
    ms, pan, label = fitted_net.scene(tag)
    net = product_net(tag)
    scene = dmf.Scene.from_raw(ms, pan, p, DEV)
    scene.set_labels(label)
    pred_map, cm = net.infer_scene(scene)
    torch.cuda.synchronize()
    pm = pred_map.cpu().numpy().reshape(-1)
    M = cm.cpu().numpy().astype(np.float64)
    assert M.sum() == H * W
    assert np.array_equal(M, orc.confusion(pm, label.reshape(-1), C)), 'confusion matrix is not the histogram of (pred, label)'
    aa, oa, k, _ = orc.aa_oa(M)
    assert len(np.unique(pm)) >= 5 and k > 0.1, 'the fitted net must be non-degenerate: %d classes, Kappa %.3f' % (len(np.unique(pm)), k)
    ref = fitted_net.fitted_net(tag).to(DEV)
    rng = np.random.default_rng(123)
    n = min(24000, H * W)
    idx = np.sort(rng.choice(H * W, size=n, replace=False))
    if n < H * W:   # corners and edges too (reflect padding on the bottom / right, the band seams of the dense path every 512 rows)
        idx[:8] = [0, W - 1, (H - 1) * W, H * W - 1, 511 * W + 5, 512 * W + 5, (H - 1) * W + W // 2, (H // 2) * W + W - 1]
    want = []
    with torch.no_grad():
        for i in range(0, n, 2000):
            a, b, _ = scene.gather(torch.from_numpy(idx[i:i + 2000]), want_target=False)
            want.append(ref(a, b).float().cpu().numpy())
    want = np.concatenate(want)
    # the per-patch product path on the same pixels: logits within tolerance of the oracle's
    h = net.native()
    logits, _ = h.forward_scene(scene, flat_idx=torch.from_numpy(idx), want_logits=True)
    err = np.abs(logits.cpu().numpy() - want)
    assert np.all(err <= logit_tol(want)), '%s: per-patch logits off by %g' % (tag, err.max())
    agree, margin = check_disagreements(pm[idx], want, tag.upper() + ' fitted, dense path')
    hist, edges = np.histogram(margin, bins=[0, 1e-3, 3e-3, 1e-2, 3e-2, 0.1, 0.3, 1, 3, 10, 1e9])
    rec = {'workload': tag, 'sampled_px': int(n), 'argmax_agreement': agree, 'classes_predicted': int(len(np.unique(pm))),
           'OA': float(oa), 'AA': float(aa), 'Kappa': float(k), 'max_abs_dlogit_per_patch_path': float(err.max()),
           'fp32_top2_margin_hist': {'bin_edges': [float(e) for e in edges[:-1]] + ['inf'], 'counts': [int(c) for c in hist]},
           'min_margin': float(margin.min())}
    print(json.dumps(rec))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out):
        with open(os.path.join(out, 'margin_hist_%s.json' % tag), 'w') as f:
            json.dump(rec, f, indent=1)
