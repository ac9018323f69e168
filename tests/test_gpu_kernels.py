"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through the
C-ABI (dmf -> ctypes -> libdmf_b200.so) and checks it against the CPU oracle or the committed
golden vectors.  Integer / byte / fp64 paths must be bit-exact; the network has stated tolerances."""
import numpy as np
import pytest
import torch

from oracle import dmf_oracle as orc

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope='module')
def dmf():
    import dmf as m
    return m


# ------------------------------------------------------------------ a1/a2: normalise + pad
def test_normalize_pad_matches_reference_golden(dmf, golden):
    g = golden('prep_gather')
    eq(dmf.normalize_pad(g['ms_u16'], 8).cpu().numpy(), g['MS_pad'])
    eq(dmf.normalize_pad(g['pan_u16'], 32).cpu().numpy(), g['PAN_pad'])
    for tag in ('u8', 'f32', 'f64'):
        want = g['pad_' + tag]
        got = dmf.normalize_pad(g['raw_' + tag], 4).cpu().numpy()
        eq(got.astype(want.dtype), want)            # float32 rasters stay float32 in the reference
        want2 = g['pad2d_' + tag]
        got2 = dmf.normalize_pad(g['raw_' + tag][:, :, 0].copy(), 16).cpu().numpy()
        eq(got2.astype(want2.dtype), want2)


# ------------------------------------------------------------------ a4/a5: K1 gather
def test_gather_golden_dual_and_tri(dmf, golden):
    g = golden('prep_gather')
    sc = dmf.Scene.from_raw(g['ms_u16'], g['pan_u16'], 8, DEV)
    sc.set_labels(g['label'])
    sc.set_mspan(g['MSPAN_pad'])
    ms, pan, tgt = sc.gather(g['pick'])
    eq(ms.cpu().numpy(), g['dual_ms'])
    eq(pan.cpu().numpy(), g['dual_pan'])
    eq(tgt.cpu().numpy(), g['dual_target'])
    ms, pan, mspan, tgt = sc.gather(g['pick'], tri=True)
    eq(ms.cpu().numpy(), g['tri_ms'])
    eq(pan.cpu().numpy(), g['tri_pan'])
    eq(mspan.cpu().numpy(), g['tri_mspan'])
    # the same scene built from data_padding()'s float64 output (drop-in path of dataset_dual)
    sc2 = dmf.Scene.from_padded(g['MS_pad'], g['PAN_pad'], 8, DEV)
    ms2, pan2, _ = sc2.gather(g['pick'], want_target=False)
    eq(ms2.cpu().numpy(), g['dual_ms'])
    eq(pan2.cpu().numpy(), g['dual_pan'])
    eq(sc.export(0).cpu().numpy(), g['MS_pad'].astype(np.float32))
    eq(sc.export(1).cpu().numpy(), g['PAN_pad'].astype(np.float32))


@pytest.mark.parametrize('p', [8, 16, 32, 6])
def test_gather_vs_oracle_random(dmf, p):
    H, W = 37, 45
    ms, pan, label = orc.synthetic_scene(H, W, 5, seed=3, label_seed=4)
    MS, PAN = orc.data_padding(ms, p), orc.data_padding(pan, p)
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc.set_labels(label)
    rng = np.random.default_rng(p)
    idx = np.concatenate([rng.integers(0, H * W, 300), [0, W - 1, (H - 1) * W, H * W - 1]])
    a, b, t = sc.gather(idx)
    ra, rb = orc.gather_dual(MS, PAN, idx // W, idx % W, p)
    eq(a.cpu().numpy(), ra)
    eq(b.cpu().numpy(), rb)
    eq(t.cpu().numpy(), label.reshape(-1)[idx].astype(np.float32))
    e0, e1, _ = sc.gather(np.zeros(0, dtype=np.int64))
    assert e0.shape[0] == 0 and e1.shape[0] == 0


def test_gather_full_size_roundtrip_property(dmf):
    """C2-sized scene: every gathered patch element equals the padded raster at its source position
    (checked through a checksum identity instead of the slow CPU gather)."""
    H = W = 1000
    p = 16
    ms, pan, _ = orc.synthetic_scene(H, W, 12, seed=0)
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    MSp, PANp = sc.export(0), sc.export(1)
    rng = np.random.default_rng(0)
    idx = torch.as_tensor(rng.integers(0, H * W, 4096))
    a, b, _ = sc.gather(idx, want_target=False)
    x, y = (idx // W).to(DEV), (idx % W).to(DEV)
    r = torch.arange(p, device=DEV)
    ref_ms = MSp[(x[:, None] + r)[:, :, None], (y[:, None] + r)[:, None, :], :].permute(0, 3, 1, 2)
    R = torch.arange(4 * p, device=DEV)
    ref_pan = PANp[(4 * x[:, None] + R)[:, :, None], (4 * y[:, None] + R)[:, None, :]][:, None]
    assert torch.equal(a, ref_ms) and torch.equal(b, ref_pan)
    # last pixel of the scene: the window is reflect padding on both axes
    k = H * W - 1
    a1, b1, _ = sc.gather([k], want_target=False)
    MS = orc.data_padding(ms[-40:, -40:], p)   # local min/max differ -> compare structure via the scene export instead
    assert torch.equal(a1[0], MSp[H - 1:H - 1 + p, W - 1:W - 1 + p].permute(2, 0, 1))
    assert torch.equal(b1[0, 0], PANp[4 * (H - 1):4 * (H - 1) + 4 * p, 4 * (W - 1):4 * (W - 1) + 4 * p])


# ------------------------------------------------------------------ a7/a8: K2 IHS
def test_ihs_and_pan2ms_golden(dmf, golden):
    g = golden('ihs')
    eq(dmf.ihs_tran(g['MS'], g['PAN'], g['offsets'], DEV).cpu().numpy(), g['MSPAN'])
    eq(dmf.pan2ms(g['pan_u16'], DEV).cpu().numpy(), g['pan2ms_u16'])
    eq(dmf.pan2ms(g['PAN'], DEV).cpu().numpy(), g['pan2ms_f64'])
    eq(dmf.pan2ms(g['PAN'].astype(np.float32), DEV).cpu().numpy(), g['pan2ms_f32'])


def test_ihs_vs_oracle_larger(dmf):
    rng = np.random.default_rng(8)
    H, W = 61, 53
    MS, PAN = rng.random((H, W, 4)), rng.random((4 * H, 4 * W))
    offs = rng.integers(0, 4, (4, H, W, 2)).astype(np.int8)
    eq(dmf.ihs_tran(MS, PAN, offs, DEV).cpu().numpy(), orc.ihs_tran_from_offsets(MS, PAN, offs))
    pan16 = rng.integers(0, 2048, (4 * H, 4 * W), dtype=np.uint16)
    eq(dmf.pan2ms(pan16, DEV).cpu().numpy(), orc.pan2ms(pan16, [H, W, 4]))


# ------------------------------------------------------------------ a11..a14: K4 / K5
def test_argmax_confusion_golden(dmf, golden):
    g = golden('metrics')
    for tag, C in (('c8', 8), ('c13', 13)):
        lg = torch.from_numpy(g[tag + '_logits']).to(DEV)
        tg = torch.from_numpy(g[tag + '_target']).to(DEV)
        pred, cm = dmf.argmax_confusion(lg, tg, C)
        eq(pred.cpu().numpy(), g[tag + '_pred'])
        M = cm.cpu().numpy().astype(np.float64)
        eq(M, g[tag + '_M'])
        aa, oa, k, _ = orc.aa_oa(M)
        eq(np.array([aa, oa, k]), g[tag + '_aa_oa_k'])
        _, cm8 = dmf.argmax_confusion(lg, tg.to(torch.uint8), C)
        eq(cm8.cpu().numpy(), cm.cpu().numpy())


def test_argmax_confusion_large_and_accumulating(dmf):
    rng = np.random.default_rng(2)
    N, C = 1_000_003, 12
    lg = rng.normal(0, 1, (N, C)).astype(np.float32)
    lg[::5, 3] = lg[::5, 7] = 5.0
    tg = rng.integers(0, C, N).astype(np.float32)
    d_lg, d_tg = torch.from_numpy(lg).to(DEV), torch.from_numpy(tg).to(DEV)
    pred, cm = dmf.argmax_confusion(d_lg, d_tg, C)
    rp = orc.argmax_first(lg)
    eq(pred.cpu().numpy(), rp)
    eq(cm.cpu().numpy().astype(np.float64), orc.confusion(rp, tg, C))
    _, cm = dmf.argmax_confusion(d_lg[:1000], d_tg[:1000], C, cm=cm, want_pred=False)     # accumulates in place
    eq(cm.cpu().numpy().astype(np.float64), orc.confusion(rp[:1000], tg[:1000], C, orc.confusion(rp, tg, C)))
    assert int(cm.sum()) == N + 1000


def test_scatter_and_paint(dmf):
    rng = np.random.default_rng(3)
    H, W, C = 67, 93, 12
    colors = [[(37 * i) % 256, (91 * i) % 256, (53 * i) % 256] for i in range(C)]
    idx = rng.permutation(H * W)[:4000]
    pred = rng.integers(0, C, idx.size)
    lm = torch.zeros((H, W), dtype=torch.uint8, device=DEV)
    dmf.scatter_labels(lm, idx // W, idx % W, pred)
    ref = orc.scatter_labels(np.zeros((H, W)), idx // W, idx % W, pred)
    eq(lm.cpu().numpy(), ref.astype(np.uint8))
    eq(dmf.paint_labels(lm, colors).cpu().numpy(), orc.paint(ref, colors))


def test_scene_update_in_place_matches_fresh_scene(dmf):
    """dmf_scene_update_raw re-fills an existing scene (new min/max, new padding) without reallocating."""
    p, H, W = 16, 40, 44
    ms1, pan1, _ = orc.synthetic_scene(H, W, 5, seed=1)
    ms2, pan2, _ = orc.synthetic_scene(H, W, 5, seed=2, blocky=True)
    sc = dmf.Scene.from_raw(ms1, pan1, p, DEV)
    sc.update_raw(torch.from_numpy(ms2.view(np.int16)).pin_memory(), torch.from_numpy(pan2.view(np.int16)).pin_memory())
    torch.cuda.synchronize()
    eq(sc.export(0).cpu().numpy(), orc.data_padding(ms2, p).astype(np.float32))
    eq(sc.export(1).cpu().numpy(), orc.data_padding(pan2, p).astype(np.float32))


def test_confusion_at_matches_oracle(dmf):
    """cm[pred_map[k]][label_map[k]] over an index list == oracle.confusion on the same samples, bit for bit."""
    g = np.random.default_rng(3)
    H, W, C = 57, 91, 13
    pm = g.integers(0, C, (H, W), dtype=np.uint8)
    lab = g.integers(0, C, (H, W), dtype=np.uint8)
    idx = g.choice(H * W, size=3000, replace=False)
    pm_d, lab_d = torch.from_numpy(pm).to(DEV), torch.from_numpy(lab).to(DEV)
    cm = dmf.confusion_at(pm_d, lab_d, torch.from_numpy(idx), C)
    assert np.array_equal(cm.cpu().numpy().astype(np.float64), orc.confusion(pm.reshape(-1)[idx], lab.reshape(-1)[idx], C))
    cm_all = dmf.confusion_at(pm_d, lab_d, None, C)
    assert np.array_equal(cm_all.cpu().numpy().astype(np.float64), orc.confusion(pm.reshape(-1), lab.reshape(-1), C))
    cm2 = dmf.confusion_at(pm_d, lab_d, torch.from_numpy(idx[:0]), C, cm=cm.clone())      # empty list: unchanged
    assert torch.equal(cm2, cm)


# ------------------------------------------------------------------ r02: TMA gather sizes, device IHS -> scene
@pytest.mark.parametrize('p', [4, 12, 20, 64])
def test_gather_tma_chunking_other_patch_sizes(dmf, p):
    """p = 4 (1 KB windows), 12 / 20 (PAN chunks that must divide the window: 48 rows, 40 + 40), 64 (16 chunks of 16 rows); dual and tri."""
    H, W = 29, 31
    ms, pan, label = orc.synthetic_scene(H, W, 5, seed=13, label_seed=14)
    MS, PAN = orc.data_padding(ms, p), orc.data_padding(pan, p)
    MSPAN = np.random.default_rng(p).random(PAN.shape)
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc.set_labels(label)
    sc.set_mspan(MSPAN)
    rng = np.random.default_rng(100 + p)
    idx = np.concatenate([rng.integers(0, H * W, 700 if p < 64 else 60), [0, W - 1, (H - 1) * W, H * W - 1]])
    a, b, c, t = sc.gather(idx, tri=True)
    ra, rb, rc = orc.gather_tri(MS, PAN, MSPAN, idx // W, idx % W, p)
    eq(a.cpu().numpy(), ra)
    eq(b.cpu().numpy(), rb)
    eq(c.cpu().numpy(), rc)
    eq(t.cpu().numpy(), label.reshape(-1)[idx].astype(np.float32))
    a2, b2, _ = sc.gather(idx, want_target=False)
    eq(a2.cpu().numpy(), ra)
    eq(b2.cpu().numpy(), rb)


def test_gather_rejects_out_of_range_host_indices(dmf):
    ms, pan, _ = orc.synthetic_scene(20, 20, 3, seed=1)
    sc = dmf.Scene.from_raw(ms, pan, 8, DEV)
    with pytest.raises(IndexError):
        sc.gather([0, 400], want_target=False)
    with pytest.raises(IndexError):
        sc.gather(np.array([-1]), want_target=False)


def test_gather_after_update_raw_sees_the_new_rasters(dmf):
    """the planar MS copy / tensor maps of K1 follow dmf_scene_update_raw"""
    p, H, W = 16, 33, 38
    ms1, pan1, _ = orc.synthetic_scene(H, W, 5, seed=1)
    ms2, pan2, _ = orc.synthetic_scene(H, W, 5, seed=2, blocky=True)
    sc = dmf.Scene.from_raw(ms1, pan1, p, DEV)
    idx = np.arange(0, H * W, 7)
    sc.gather(idx, want_target=False)
    sc.update_raw(ms2, pan2)
    a, b, _ = sc.gather(idx, want_target=False)
    ra, rb = orc.gather_dual(orc.data_padding(ms2, p), orc.data_padding(pan2, p), idx // W, idx % W, p)
    eq(a.cpu().numpy(), ra)
    eq(b.cpu().numpy(), rb)


@pytest.mark.parametrize('dt', ['u16', 'u8', 'f32', 'f64'])
def test_scene_mspan_from_ihs_on_device_matches_oracle(dmf, dt):
    """dmf_scene_set_mspan_ihs == float32(reflect-pad(IHS_tran(to_tensor(ms), to_tensor(pan)))) bit for bit
    (image_convert/IHS.py:40-54 on function/function.py:99-124), then the tri gather reads it."""
    rng = np.random.default_rng(17)
    H, W, p = 23, 27, 8
    if dt == 'u16':
        ms, pan = rng.integers(0, 2048, (H, W, 4), dtype=np.uint16), rng.integers(0, 2048, (4 * H, 4 * W), dtype=np.uint16)
    elif dt == 'u8':
        ms, pan = rng.integers(2, 251, (H, W, 4), dtype=np.uint8), rng.integers(0, 256, (4 * H, 4 * W), dtype=np.uint8)
    else:
        t = np.float32 if dt == 'f32' else np.float64
        ms, pan = rng.normal(100, 30, (H, W, 4)).astype(t), rng.normal(90, 25, (4 * H, 4 * W)).astype(t)
    offs = rng.integers(0, 4, (4, H, W, 2)).astype(np.int8)
    prod = orc.ihs_tran_from_offsets(orc.to_tensor(ms), orc.to_tensor(pan), offs)
    rows = orc.reflect101_index(np.arange(4 * H + 4 * p - 1), 4 * H)
    cols = orc.reflect101_index(np.arange(4 * W + 4 * p - 1), 4 * W)
    want = prod[rows][:, cols].astype(np.float32)
    sc = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc.set_mspan_ihs(ms, pan, offs)
    eq(sc.export(2).cpu().numpy(), want)
    # device-resident inputs give the same raster; so does the host-padded route it replaces
    as_t = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(DEV)
    sc2 = dmf.Scene.from_raw(ms, pan, p, DEV)
    sc2.set_mspan_ihs(as_t(ms), as_t(pan), torch.from_numpy(offs).to(DEV))
    eq(sc2.export(2).cpu().numpy(), want)
    idx = np.array([0, W - 1, (H - 1) * W, H * W - 1, 5 * W + 3])
    _, _, c, _ = sc.gather(idx, tri=True, want_target=False)
    _, rc = orc.gather_dual(orc.data_padding(ms, p), prod[rows][:, cols], idx // W, idx % W, p)
    eq(c.cpu().numpy(), rc)
