"""CPU tests of the host-side mirror (no GPU compute): interface names, index lists, RNG replay,
config rendering, metrics, band partition, C-ABI symbol table."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle import dmf_oracle as orc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, 'include', 'dmf_b200.h')).read()
    declared = set(re.findall(r'\b(dmf_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'dmf_status', 'dmf_dtype'}
    import dmf._lib as L
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dmf_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import dmf
    from model.gmfnet import Net
    with pytest.raises(RuntimeError):
        dmf.Scene.from_raw(np.zeros((4, 4, 4), np.uint16), np.zeros((16, 16), np.uint16), 8)
    net = Net({'Categories_Number': 8, 'patch_size': 16, 'schedule': {'activate': 'Relu'}}).eval()
    with torch.no_grad(), pytest.raises(RuntimeError):
        net(torch.zeros(1, 4, 16, 16), torch.zeros(1, 1, 64, 64))


def test_product_tree_never_imports_the_oracle():
    pkg = os.path.join(REPO, 'dual-modal-fusion_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(root, f)).read()
                assert 'oracle' not in src.replace('debug oracle', '').replace('device-side debug', ''), os.path.join(root, f)


def test_split_data_old_matches_golden(golden):
    from function.function import split_data_old, split_data
    g = golden('prep_gather')
    H, W = g['label'].shape
    cfg = {'data_city': 'x', 'DATA_DICT': {'x': {'size': [H, W, 4]}}}
    xyl, mat_ = split_data_old(g['label'], cfg)
    assert np.array_equal(np.stack([m[:, 0] for m in xyl]), g['xyl'])
    assert mat_[0] == g['idx_unlabelled'].tolist() and mat_[1] == g['idx_labelled'].tolist()
    assert all(m.shape == (H * W, 1) and m.dtype == np.float64 for m in xyl)
    tr = (g['label'] == 1).astype(np.uint8)
    te = (g['label'] >= 1).astype(np.uint8)
    _, m3 = split_data(tr, te, g['label'], cfg)
    flat = g['label'].reshape(-1)
    assert m3[1] == np.flatnonzero(flat == 1).tolist() and m3[2] == np.flatnonzero(flat > 1).tolist()
    assert m3[0] == np.flatnonzero(flat == 0).tolist()


def test_offsets_replay_python_mersenne_twister():
    from image_convert.IHS import draw_offsets, unpooling
    for seed, (H, W) in [(1234, (6, 9)), (5, (23, 31))]:
        random.seed(seed)
        want = orc.draw_unpooling_offsets(H, W, 4, 4)
        state_want = random.getstate()
        random.seed(seed)
        got = draw_offsets(H, W, 4, 4)
        assert np.array_equal(got, want) and random.getstate() == state_want


def test_unpooling_matches_reference_golden(golden):
    from image_convert.IHS import unpooling
    g = golden('ihs')
    random.seed(99)
    assert np.array_equal(unpooling(g['MS'], 4), g['unpooled_seed99'])


def test_metrics_match_golden(golden):
    from indicators.kappa import aa_oa, kappa
    g = golden('metrics')
    for tag in ('c8', 'c13'):
        aa, oa, k, rows = aa_oa(g[tag + '_M'])
        assert np.array_equal(np.array([aa, oa, k]), g[tag + '_aa_oa_k'], equal_nan=True)
        assert np.array_equal(np.asarray(rows, dtype=np.float64), g[tag + '_rows'], equal_nan=True)
        assert np.float64(kappa(g[tag + '_M'])) == g[tag + '_kappa']


def test_config_renders_and_adds_dqtl(tmp_path, monkeypatch):
    from utils.config import get_render_config
    work = tmp_path / 'a' / 'b'
    work.mkdir(parents=True)
    monkeypatch.chdir(work)
    cfg = get_render_config(os.path.join(REPO, 'dual-modal-fusion_b200', 'config.yml'))
    assert cfg['Categories_Number'] == 12 and cfg['patch_size'] == 16 and cfg['model_name'] == 'gmfnet'
    assert cfg['schedule']['lr'] == 1e-3 and isinstance(cfg['dqtl']['epsilon'], float)
    assert cfg['RESULT_output'].endswith('gmfnet__0_output/') and os.path.isdir(cfg['RESULT_output'])
    assert cfg['parameters'] == 'image6_tr0.02_ep50_bs256'


def test_row_bands_partition_the_scene():
    from solver.mainsolver import row_band
    for H in (128, 1000, 2001):
        for world in (1, 2, 4, 8):
            bands = [row_band(H, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in bands]
            assert max(sizes) - min(sizes) <= 1


def test_state_dict_is_interchangeable_with_the_oracle_net():
    from model.gmfnet import Net
    from oracle.gmfnet_ref import Net as RefNet
    cfg = {'Categories_Number': 12, 'patch_size': 16, 'schedule': {'activate': 'Relu'}}
    a, b = Net(cfg), RefNet(cfg)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    a.load_state_dict(b.state_dict())
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))


def test_patch_loader_index_order_matches_reference_run(golden):
    """BaseSolver.dataloader() semantics on the host only: same random_split / RandomSampler draws
    under torch.manual_seed(3407) as the reference's Solver produced (golden solver_c1)."""
    from train.dataset import PatchLoader
    g = golden('solver_c1')
    H = W = 128
    _, _, label = orc.synthetic_scene(H, W, 7, seed=0, label_seed=1)
    labelled = np.flatnonzero(label.reshape(-1) != 0)
    torch.manual_seed(3407)
    n = len(labelled)
    tr = int(0.02 * n)
    parts = torch.utils.data.random_split(range(n), [tr, n - 2 * tr, tr])
    assert np.array_equal(labelled[parts[0].indices], g['train_idx'])

    class FakeDataset:
        def gather_batch(self, idx):
            return idx
    loader = PatchLoader(FakeDataset(), labelled[parts[0].indices], 256, shuffle=True)
    first = next(iter(loader))
    assert np.array_equal(first // W, g['train_batch0_x']) and np.array_equal(first % W, g['train_batch0_y'])
    assert len(loader) == 2


def test_band_slice_reproduces_whole_scene_windows():
    """A rank that holds only rows band_slice(H, p, r0, r1) of the scene (dmf.band_slice, bench.py e2e arm at N > 1) sees, after
    the reference's reflect padding of ITS rows, the same windows as the whole padded scene for every anchor of [r0, r1)."""
    from dmf import band_slice
    rng = np.random.default_rng(0)
    for H, p, world in [(45, 16, 3), (20, 16, 2), (33, 16, 1), (10, 16, 2), (64, 8, 5), (100, 32, 3), (17, 16, 17)]:
        W = 5
        scene = rng.random((H, W))
        whole = np.pad(scene, ((0, p - 1), (0, 0)), mode='reflect') if H > 1 else np.repeat(scene, p, 0)
        covered = np.zeros(H, dtype=int)
        for rank in range(world):
            step = -(-H // world)
            r0, r1 = min(H, rank * step), min(H, (rank + 1) * step)
            if r0 == r1:
                continue
            s0, s1 = band_slice(H, p, r0, r1)
            assert 0 <= s0 <= r0 and r1 <= s1 <= H
            band = scene[s0:s1]
            band_pad = np.pad(band, ((0, p - 1), (0, 0)), mode='reflect') if band.shape[0] > 1 else np.repeat(band, p, 0)
            for x in range(r0, r1):
                assert np.array_equal(band_pad[x - s0:x - s0 + p], whole[x:x + p]), (H, p, world, rank, x)
            covered[r0:r1] += 1
        assert (covered == 1).all()


@pytest.mark.parametrize('aligned', [0, 1])
def test_dense_class_tables_match_brute_force(aligned):
    """The window / box table of conv_pool4_kernel (csrc/dense.cu::build_pool4_cls) against a brute-force walk over a patch:
    for pooled border class (a, b), sub-position (s, t) and tap (dy, dx), the input the conv reads sits at a definite offset from
    the pooled cell's origin, carries a definite border variant, and (aligned pooling) lives in a definite parity phase.  The
    table must send that (offset) to a box holding exactly that plane, at exactly that cell shift."""
    import dmf._lib as L
    box_rows, box_cols, kq = (17, 9, 4) if aligned else (19, 11, 2)
    plane_bytes = box_rows * box_cols * 16
    for a in range(3):
        for b in range(3):
            win = (ctypes.c_int16 * 16)()
            bp = (ctypes.c_int16 * 9)()
            dr, dc = (ctypes.c_int8 * 9)(), (ctypes.c_int8 * 9)()
            nb, slot = ctypes.c_int32(), ctypes.c_int32()
            L.check(L.lib.dmf_dense_class_table(a, b, aligned, win, bp, dr, dc, ctypes.byref(nb), ctypes.byref(slot)))
            assert slot.value % 128 == 0 and slot.value >= kq * plane_bytes
            for P2 in (4, 8, 16):                                 # pooled cells per axis of the layer's OUTPUT (p/2 or p)
                S = 2 * P2                                        # conv positions per axis
                cell = {0: 0, 1: 1, 2: P2 - 1}                    # a representative cell index of each border class
                for s in range(2):
                    for t in range(2):
                        for dy in (-1, 0, 1):
                            for dx in (-1, 0, 1):
                                if aligned:
                                    i, j = 2 * cell[a] + s + dy, 2 * cell[b] + t + dx      # patch-relative input position
                                else:
                                    # stride-1 pooling: the layer's input and conv grids coincide; cell k covers conv rows k', k'+1 of
                                    # a patch whose pooled cell k sits at conv row 2k: same arithmetic with the conv row 2k + s
                                    i, j = 2 * cell[a] + s + dy, 2 * cell[b] + t + dx
                                orow, ocol = s + dy, t + dx
                                w = win[(orow + 1) * 4 + ocol + 1]
                                if not (0 <= i < S and 0 <= j < S):
                                    assert w == -1, (a, b, s, t, dy, dx)
                                    continue
                                assert w >= 0
                                vr = 0 if i == 0 else (2 if i == S - 1 else 1)
                                vc = 0 if j == 0 else (2 if j == S - 1 else 1)
                                k, rem = divmod(w * 16, slot.value)
                                assert k < nb.value and rem % 16 == 0
                                r_in, c_in = divmod(rem // 16, box_cols)
                                assert r_in + 16 <= box_rows and c_in + 8 <= box_cols
                                if aligned:
                                    want_plane = (vr * 3 + vc) * 4 + (orow & 1) * 2 + (ocol & 1)
                                    want_shift = (orow // 2, ocol // 2)                    # floor division: -1 -> -1
                                else:
                                    want_plane, want_shift = vr * 3 + vc, (orow, ocol)
                                assert bp[k] == want_plane, (a, b, s, t, dy, dx, bp[k], want_plane)
                                assert (dr[k] + r_in, dc[k] + c_in) == want_shift
