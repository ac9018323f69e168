"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions.

Run in the authoring container only (the reference lives at /root/reference, read-only, and
does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference imports libtiff / h5py / openpyxl / matplotlib at module scope; none is installed
and none is on the hot path, so empty stand-in modules are registered first (SURVEY.md
appendix B).  No reference file is modified or copied; only the numeric OUTPUTS of its functions
on seeded synthetic inputs are stored.
"""
import os
import random
import sys
import types

import numpy as np
import torch

REF = os.environ.get('DMF_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))

for name in ['libtiff', 'h5py', 'openpyxl', 'matplotlib', 'matplotlib.pyplot']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['libtiff'].TIFF = object
sys.modules['openpyxl'].Workbook = object
sys.modules['openpyxl'].load_workbook = lambda *a, **k: None
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
sys.path.insert(0, REF)

import function.function as rf          # noqa: E402
import train.dataset as rd              # noqa: E402
import indicators.kappa as rk           # noqa: E402
import image_convert.IHS as ri          # noqa: E402
import solver.basesolver as rbs         # noqa: E402
import solver.mainsolver as rms         # noqa: E402

sys.path.insert(0, REPO)
from oracle import dmf_oracle as orc    # noqa: E402
from oracle import gmfnet_ref            # noqa: E402


def quiet(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def small_cfg(H, W, p, n_classes, city='synthetic'):
    colors = [[(37 * i) % 256, (91 * i) % 256, (53 * i) % 256] for i in range(n_classes + 1)]
    return {'patch_size': p, 'data_city': city, 'DATA_DICT': {city: {'size': [H, W, 4], 'color': colors}}}


def golden_prep():
    """a1, a2, a3, a4: normalise + pad, index lists, dual/tri crops through the real DataLoader."""
    out = {}
    H, W, p = 21, 26, 8
    cfg = small_cfg(H, W, p, 5)
    ms, pan, label = orc.synthetic_scene(H, W, 5, seed=11, label_seed=12)
    out['ms_u16'], out['pan_u16'], out['label'] = ms, pan, label
    MS = quiet(rf.data_padding, ms, cfg, 'ms')
    PAN = quiet(rf.data_padding, pan, cfg, 'pan')
    out['MS_pad'], out['PAN_pad'] = MS, PAN
    # other raster dtypes (to_tensor follows numpy's promotion rules)
    rng = np.random.default_rng(5)
    for tag, arr in [('u8', rng.integers(3, 250, (9, 7, 4), dtype=np.uint8)),
                     ('f32', rng.normal(100, 30, (9, 7, 4)).astype(np.float32)),
                     ('f64', rng.normal(100, 30, (9, 7, 4)))]:
        out['raw_' + tag] = arr
        out['pad_' + tag] = quiet(rf.data_padding, arr, small_cfg(9, 7, 4, 2), 'ms')
        out['pad2d_' + tag] = quiet(rf.data_padding, arr[:, :, 0].copy(), small_cfg(9, 7, 4, 2), 'pan')
    xyl, mat_ = quiet(rf.split_data_old, label, cfg)
    out['xyl'] = np.stack([m[:, 0] for m in xyl])
    out['idx_unlabelled'] = np.asarray(mat_[0], dtype=np.int64)
    out['idx_labelled'] = np.asarray(mat_[1], dtype=np.int64)
    ds = rd.dataset_dual(MS, PAN, xyl, cfg)
    pick = [0, W - 1, (H - 1) * W, H * W - 1, 5 * W + 7, 13 * W + 25, 20 * W + 3, 137, 138, 139]
    loader = torch.utils.data.DataLoader(torch.utils.data.Subset(ds, pick), batch_size=len(pick), shuffle=False)
    d1, d2, tgt, x, y = next(iter(loader))
    out['pick'] = np.asarray(pick, dtype=np.int64)
    out['dual_ms'], out['dual_pan'] = d1.numpy(), d2.numpy()
    out['dual_target'], out['dual_x'], out['dual_y'] = tgt.numpy(), x.numpy(), y.numpy()
    # tri: third raster on the PAN grid (any float64 raster of PAN's padded shape)
    MSPAN = np.random.default_rng(6).random(PAN.shape)
    out['MSPAN_pad'] = MSPAN
    dt = rd.dataset_tri(MS, PAN, MSPAN, xyl[2], xyl[0], xyl[1], p)
    loader = torch.utils.data.DataLoader(torch.utils.data.Subset(dt, pick), batch_size=len(pick), shuffle=False)
    t1, t2, t3, ttgt, tx, ty = next(iter(loader))
    out['tri_ms'], out['tri_pan'], out['tri_mspan'] = t1.numpy(), t2.numpy(), t3.numpy()
    np.savez_compressed(os.path.join(HERE, 'prep_gather.npz'), **out)


def golden_ihs():
    """a7, a8: IHS_tran with a seeded Python RNG, pan2ms on integer and float rasters."""
    out = {}
    H, W = 6, 9
    rng = np.random.default_rng(21)
    MS = rng.random((H, W, 4))
    PAN = rng.random((4 * H, 4 * W))
    out['MS'], out['PAN'] = MS, PAN
    random.seed(1234)
    out['MSPAN'] = ri.IHS_tran(MS, PAN)
    random.seed(1234)
    out['offsets'] = orc.draw_unpooling_offsets(H, W, 4, 4)        # same stream, replayed
    random.seed(99)
    out['unpooled_seed99'] = ri.unpooling(MS, 4)
    pan_u16 = rng.integers(0, 2048, (4 * H, 4 * W), dtype=np.uint16)
    out['pan_u16'] = pan_u16
    out['pan2ms_u16'] = ri.pan2ms(pan_u16, [H, W, 4])
    out['pan2ms_f64'] = ri.pan2ms(PAN, [H, W, 4])
    pan_f32 = PAN.astype(np.float32)
    out['pan2ms_f32'] = ri.pan2ms(pan_f32, [H, W, 4])
    np.savez_compressed(os.path.join(HERE, 'ihs.npz'), **out)


def golden_metrics():
    """a11, a13, a14: the confusion loop of solver/mainsolver.py:139-141 run literally on torch
    tensors, then the reference aa_oa / kappa on the resulting matrices."""
    out = {}
    rng = np.random.default_rng(31)
    for tag, C, N in [('c8', 8, 700), ('c13', 13, 1500)]:
        logits = torch.tensor(rng.normal(0, 1, (N, C)).astype(np.float32))
        logits[::17, 2] = logits[::17, 5] = 9.0                     # exact ties -> first index wins
        target = torch.tensor(rng.integers(0, C, N).astype(np.float32))
        if tag == 'c13':
            target[target == 4] = 3.0                               # class 4 absent -> NaN accuracy
        M = np.zeros([C, C])
        pred = logits.data.max(1, keepdim=True)[1]
        for i in range(len(target)):
            M[int(pred[i].item())][int(target[i].item())] += 1
        aa, oa, k, rows = quiet(rk.aa_oa, M)
        out[tag + '_logits'], out[tag + '_target'] = logits.numpy(), target.numpy()
        out[tag + '_pred'] = pred.numpy()[:, 0]
        out[tag + '_M'] = M
        out[tag + '_aa_oa_k'] = np.array([aa, oa, k])
        out[tag + '_rows'] = np.asarray(rows, dtype=np.float64)
        out[tag + '_kappa'] = np.float64(rk.kappa(M))
    np.savez_compressed(os.path.join(HERE, 'metrics.npz'), **out)


def golden_solver_c1():
    """a6 + the whole C1 path through the reference's own Solver objects: BaseSolver.__init__,
    dataloader() under torch.manual_seed(3407) (test.py:8), whole-scene colour loop and the clean
    confusion loop (train/test.py:58-60) with the oracle Net injected as model.gmfnet."""
    import tempfile
    H = W = 128
    p, ncls = 16, 7
    ms, pan, label = orc.synthetic_scene(H, W, ncls, seed=0, label_seed=1)
    tmp = tempfile.mkdtemp() + '/'
    np.save(tmp + 'label.npy', label)
    cfg = small_cfg(H, W, p, ncls, city='c1')
    cfg.update({'task': 'classification', 'time': 1, 'index': 0, 'epoch': 1, 'device': 'cpu', 'gpu_mode': False,
                'data_new': 0, 'data_address': tmp, 'use_h5': False, 'nohup': 0, 'model_name': 'gmfnet',
                'batchsize': 256, 'test_batchsize': 300, 'color_batchsize': 300, 'train_rate': 0.02,
                'verify_rate': 0.02, 'Categories_Number': ncls + 1,
                'schedule': {'loss': 'Criterion', 'optimizer': 'ADAM', 'if_scheduler': 0, 'scheduler': 'ExponentialLR',
                             'activate': 'Relu', 'lr': 1e-3, 'base_lr': 5e-4},
                'train': {'index': 0, 'pretrained': 0, 'save_best': True}, 'test': {'index': 1, 'save_matrix': 1},
                'color': {'index': 1, 'supervised': 1, 'unsupervised': 1}})
    rbs.read_tif = lambda c, mode: ms if mode == 'ms' else pan
    mod = types.ModuleType('model.gmfnet')
    mod.Net = gmfnet_ref.Net
    sys.modules['model'] = types.ModuleType('model')
    sys.modules['model.gmfnet'] = mod
    torch.manual_seed(3407)
    s = quiet(rms.Solver, cfg)
    quiet(s.dataloader)
    out = {'train_idx': np.asarray([s.train_loader.dataset.dataset.indices[i] for i in s.train_loader.dataset.indices]),
           'test_idx': np.asarray([s.test_loader.dataset.dataset.indices[i] for i in s.test_loader.dataset.indices]),
           'valid_idx': np.asarray([s.valid_loader.dataset.dataset.indices[i] for i in s.valid_loader.dataset.indices])}
    # first shuffled training batch order (RandomSampler under the same generator state)
    it = iter(s.train_loader)
    b = next(it)
    out['train_batch0_x'], out['train_batch0_y'] = b[3].numpy(), b[4].numpy()
    quiet(s.init_model)
    net = s.model.eval()
    out['state_keys'] = np.asarray(list(net.state_dict().keys()))
    M = np.zeros([ncls + 1, ncls + 1])
    label_map = np.zeros([H, W])
    logits_all = np.zeros((H * W, ncls + 1), dtype=np.float32)
    with torch.no_grad():
        for loader in (s.color_loader1, s.color_loader2):
            for d1, d2, tgt, x, y in loader:
                o = net(d1, d2)
                pred = o.data.max(1, keepdim=True)[1]
                for i in range(len(tgt)):
                    M[int(pred[i].item())][int(tgt[i].item())] += 1
                    label_map[int(x[i])][int(y[i])] = int(pred[i])
                logits_all[(x * W + y).numpy()] = o.numpy()
    aa, oa, k, _ = quiet(rk.aa_oa, M)
    out['M'], out['label_map'] = M, label_map.astype(np.uint8)
    out['aa_oa_k'] = np.array([aa, oa, k])
    out['logits_first512'] = logits_all[:512]
    out['logits_checksum'] = np.float64(logits_all.astype(np.float64).sum())
    np.savez_compressed(os.path.join(HERE, 'solver_c1.npz'), **out)


def golden_solver_c1_fitted():
    """The C1 whole-scene path once more, on a network whose predictions VARY: the structured synthetic scene (labels depend
    on the rasters) and the fitted net of oracle/fitted_net.py (seed-3407 convolutions, calibrated BatchNorm, fitted head),
    again through the reference's own BaseSolver / DataLoader / dataset_dual objects and the clean confusion loop
    (train/test.py:58-60).  The default-initialised net of golden_solver_c1 predicts a single class everywhere."""
    import tempfile
    from oracle import fitted_net
    H = W = 128
    p, ncls = 16, 7
    ms, pan, label = fitted_net.scene('c1')
    tmp = tempfile.mkdtemp() + '/'
    np.save(tmp + 'label.npy', label)
    cfg = small_cfg(H, W, p, ncls, city='c1')
    cfg.update({'task': 'classification', 'time': 1, 'index': 0, 'epoch': 1, 'device': 'cpu', 'gpu_mode': False,
                'data_new': 0, 'data_address': tmp, 'use_h5': False, 'nohup': 0, 'model_name': 'gmfnet',
                'batchsize': 256, 'test_batchsize': 300, 'color_batchsize': 300, 'train_rate': 0.02,
                'verify_rate': 0.02, 'Categories_Number': ncls + 1,
                'schedule': {'loss': 'Criterion', 'optimizer': 'ADAM', 'if_scheduler': 0, 'scheduler': 'ExponentialLR',
                             'activate': 'Relu', 'lr': 1e-3, 'base_lr': 5e-4},
                'train': {'index': 0, 'pretrained': 0, 'save_best': True}, 'test': {'index': 1, 'save_matrix': 1},
                'color': {'index': 1, 'supervised': 1, 'unsupervised': 1}})
    rbs.read_tif = lambda c, mode: ms if mode == 'ms' else pan

    class FittedNet(gmfnet_ref.Net):
        def __init__(self, args):
            super().__init__(args)
            self.load_state_dict(fitted_net.fitted_state('c1'))

    mod = types.ModuleType('model.gmfnet')
    mod.Net = FittedNet
    sys.modules['model'] = types.ModuleType('model')
    sys.modules['model.gmfnet'] = mod
    torch.manual_seed(3407)
    s = quiet(rms.Solver, cfg)
    quiet(s.dataloader)
    quiet(s.init_model)
    net = s.model.eval()
    C = ncls + 1
    M = np.zeros([C, C])
    M_test = np.zeros([C, C])
    label_map = np.zeros([H, W])
    logits_all = np.zeros((H * W, C), dtype=np.float32)
    with torch.no_grad():
        for loader in (s.color_loader1, s.color_loader2):
            for d1, d2, tgt, x, y in loader:
                o = net(d1, d2)
                pred = o.data.max(1, keepdim=True)[1]
                for i in range(len(tgt)):
                    M[int(pred[i].item())][int(tgt[i].item())] += 1
                    label_map[int(x[i])][int(y[i])] = int(pred[i])
                logits_all[(x * W + y).numpy()] = o.numpy()
        # Solver.test()'s loop over the WHOLE test loader (train/test.py:58-60 semantics, no first-batch break)
        for d1, d2, tgt, x, y in s.test_loader:
            pred = net(d1, d2).data.max(1, keepdim=True)[1]
            for i in range(len(tgt)):
                M_test[int(pred[i].item())][int(tgt[i].item())] += 1
    aa, oa, k, _ = quiet(rk.aa_oa, M)
    aat, oat, kt, _ = quiet(rk.aa_oa, M_test)
    srt = np.sort(logits_all, axis=1)
    out = {'M': M, 'label_map': label_map.astype(np.uint8), 'aa_oa_k': np.array([aa, oa, k]),
           'M_test': M_test, 'test_aa_oa_k': np.array([aat, oat, kt]),
           'test_idx': np.asarray([s.test_loader.dataset.dataset.indices[i] for i in s.test_loader.dataset.indices]),
           'logits': logits_all.astype(np.float16),      # coarse copy of all 16 384 rows, for the tolerance check only
           'logits_first512': logits_all[:512], 'margin_top2': (srt[:, -1] - srt[:, -2]).astype(np.float32),
           'logits_checksum': np.float64(logits_all.astype(np.float64).sum())}
    print('c1 fitted: predicted classes %d, OA %.4f AA %.4f Kappa %.4f; min top-2 margin %.3g' %
          (len(np.unique(label_map)), oa, aa, k, out['margin_top2'].min()))
    np.savez_compressed(os.path.join(HERE, 'solver_c1_fitted.npz'), **out)


if __name__ == '__main__':
    golden_prep()
    golden_ihs()
    golden_metrics()
    golden_solver_c1()
    golden_solver_c1_fitted()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))
