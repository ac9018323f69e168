"""Python face of the C-ABI: thin object wrappers (Scene, NetHandle) and tensor-level helpers.

torch is used only for device memory, streams and (in solver/) torch.distributed.  Everything that
computes goes through libdmf_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import F32, F64, U8, U16, check, lib

_NP2DT = {np.dtype(np.uint8): U8, np.dtype(np.uint16): U16, np.dtype(np.float32): F32, np.dtype(np.float64): F64}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('dual-modal-fusion_b200 needs a CUDA device (sm_100a); there is no CPU fallback')


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def np_dtype_code(a):
    try:
        return _NP2DT[np.dtype(a.dtype)]
    except KeyError:
        raise TypeError('raster dtype %s is not supported (uint8, uint16, float32, float64)' % a.dtype)


def launch_count():
    return int(lib.dmf_launch_count())


def _as_dev(a, device):
    """numpy / torch, host or device -> contiguous tensor on `device`."""
    if isinstance(a, np.ndarray):
        if a.dtype == np.uint16:   # torch has limited uint16 support: move the bytes
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.int16))
        else:
            t = torch.from_numpy(np.ascontiguousarray(a))
    else:
        t = a.contiguous()
    return t.to(device, non_blocking=True)


class Scene:
    """Normalised, reflect-padded MS/PAN rasters resident in HBM
    (data_padding of function/function.py:99-117 for both rasters + the dataset's float32 cast)."""

    def __init__(self, handle, H, W, p, device):
        self._h, self.H, self.W, self.p, self.device = handle, H, W, p, device
        self.Hp, self.Wp = H + p - 1, W + p - 1
        self.H4p, self.W4p = 4 * H + 4 * p - 1, 4 * W + 4 * p - 1
        self.has_labels = False
        self.has_mspan = False

    @classmethod
    def from_raw(cls, ms, pan, p, device='cuda:0'):
        """ms [H,W,4], pan [4H,4W]: numpy arrays (host, copied) or CUDA tensors."""
        _require_cuda()
        H, W = int(ms.shape[0]), int(ms.shape[1])
        assert tuple(ms.shape) == (H, W, 4) and tuple(pan.shape) == (4 * H, 4 * W), 'MS must be [H,W,4], PAN [4H,4W]'
        h = C.c_void_p()
        with torch.cuda.device(device):
            if isinstance(ms, np.ndarray):
                ms_c, pan_c = np.ascontiguousarray(ms), np.ascontiguousarray(pan)
                check(lib.dmf_scene_create_raw(C.byref(h), ms_c.ctypes.data_as(C.c_void_p), np_dtype_code(ms_c),
                                               pan_c.ctypes.data_as(C.c_void_p), np_dtype_code(pan_c), H, W, p, 0, _stream()))
                torch.cuda.current_stream().synchronize()   # host buffers may be pageable
            else:                                  # torch tensors: CUDA, or (pinned) host memory
                code = {torch.uint8: U8, torch.int16: U16, torch.uint16: U16, torch.float32: F32, torch.float64: F64}
                ms_c, pan_c = ms.contiguous(), pan.contiguous()
                assert ms_c.is_cuda == pan_c.is_cuda
                check(lib.dmf_scene_create_raw(C.byref(h), _ptr(ms_c), code[ms_c.dtype], _ptr(pan_c), code[pan_c.dtype],
                                               H, W, p, 1 if ms_c.is_cuda else 0, _stream()))
                if not ms_c.is_cuda and not (ms_c.is_pinned() and pan_c.is_pinned()):
                    torch.cuda.current_stream().synchronize()
        return cls(h, H, W, p, device)

    def update_raw(self, ms, pan, ms_range=None, pan_range=None):
        """Re-fill this scene from new rasters of the same shape (pinned CPU or CUDA tensors, or ndarrays).
        ms_range / pan_range: float64 CUDA tensors {min, max} to normalise with instead of the rasters' own ranges (row-band
        scenes: the all-reduced ranges of the whole scene, see band_slice / raster_minmax)."""
        code = {torch.uint8: U8, torch.int16: U16, torch.uint16: U16, torch.float32: F32, torch.float64: F64}
        ranged = ms_range is not None
        assert ranged == (pan_range is not None), 'give both ranges or neither'
        with torch.cuda.device(self.device):
            if isinstance(ms, np.ndarray):
                a, b = np.ascontiguousarray(ms), np.ascontiguousarray(pan)
                if ranged:
                    check(lib.dmf_scene_update_raw_range(self._h, a.ctypes.data_as(C.c_void_p), np_dtype_code(a),
                                                         b.ctypes.data_as(C.c_void_p), np_dtype_code(b), 0, _ptr(ms_range), _ptr(pan_range), _stream()))
                else:
                    check(lib.dmf_scene_update_raw(self._h, a.ctypes.data_as(C.c_void_p), np_dtype_code(a),
                                                   b.ctypes.data_as(C.c_void_p), np_dtype_code(b), 0, _stream()))
                torch.cuda.current_stream().synchronize()
            else:
                a, b = ms.contiguous(), pan.contiguous()
                assert tuple(a.shape) == (self.H, self.W, 4) and tuple(b.shape) == (4 * self.H, 4 * self.W)
                if ranged:
                    check(lib.dmf_scene_update_raw_range(self._h, _ptr(a), code[a.dtype], _ptr(b), code[b.dtype], 1 if a.is_cuda else 0,
                                                         _ptr(ms_range), _ptr(pan_range), _stream()))
                else:
                    check(lib.dmf_scene_update_raw(self._h, _ptr(a), code[a.dtype], _ptr(b), code[b.dtype], 1 if a.is_cuda else 0, _stream()))
                if not a.is_cuda and not (a.is_pinned() and b.is_pinned()):
                    torch.cuda.current_stream().synchronize()
        return self

    @classmethod
    def from_padded(cls, ms_pad, pan_pad, p, device='cuda:0'):
        """ms_pad / pan_pad exactly as data_padding() returns them (float64 or float32 ndarrays)."""
        _require_cuda()
        H, W = ms_pad.shape[0] - p + 1, ms_pad.shape[1] - p + 1
        assert tuple(pan_pad.shape) == (4 * H + 4 * p - 1, 4 * W + 4 * p - 1), 'PAN padded shape mismatch'
        assert ms_pad.dtype == pan_pad.dtype
        h = C.c_void_p()
        a, b = np.ascontiguousarray(ms_pad), np.ascontiguousarray(pan_pad)
        with torch.cuda.device(device):
            check(lib.dmf_scene_create_padded(C.byref(h), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                              np_dtype_code(a), H, W, p, 0, _stream()))
            torch.cuda.current_stream().synchronize()
        return cls(h, H, W, p, device)

    def set_labels(self, label):
        """label: uint8 [H,W] ndarray, or a (pinned) CPU / CUDA uint8 tensor."""
        with torch.cuda.device(self.device):
            if isinstance(label, torch.Tensor):
                lab = label.contiguous()
                assert lab.dtype == torch.uint8 and tuple(lab.shape) == (self.H, self.W)
                check(lib.dmf_scene_set_labels(self._h, _ptr(lab), 1 if lab.is_cuda else 0, _stream()))
                if not lab.is_cuda and not lab.is_pinned():
                    torch.cuda.current_stream().synchronize()
            else:
                lab = np.ascontiguousarray(label, dtype=np.uint8)
                assert lab.shape == (self.H, self.W)
                check(lib.dmf_scene_set_labels(self._h, lab.ctypes.data_as(C.c_void_p), 0, _stream()))
                torch.cuda.current_stream().synchronize()
        self.has_labels = True

    def set_mspan(self, mspan_pad):
        a = np.ascontiguousarray(mspan_pad)
        assert a.shape == (self.H4p, self.W4p)
        with torch.cuda.device(self.device):
            check(lib.dmf_scene_set_mspan(self._h, a.ctypes.data_as(C.c_void_p), np_dtype_code(a), 0, _stream()))
            torch.cuda.current_stream().synchronize()
        self.has_mspan = True

    def set_mspan_ihs(self, ms, pan, offsets, ms_range=None, pan_range=None):
        """The IHS product of the scene's own rasters, computed and reflect-padded on the device (no host round trip):
        float32(IHS_tran(to_tensor(ms), to_tensor(pan))) of image_convert/IHS.py:40-54 as dataset_tri's third raster.
        ms [H,W,4] / pan [4H,4W]: the RAW rasters (ndarrays, pinned CPU or CUDA tensors) the scene was built from; offsets: int8
        [4,H,W,2] unpooling draws (image_convert.IHS.draw_offsets), ndarray or CUDA tensor; ms_range / pan_range as in update_raw."""
        code = {torch.uint8: U8, torch.int16: U16, torch.uint16: U16, torch.float32: F32, torch.float64: F64}
        assert (ms_range is None) == (pan_range is None), 'give both ranges or neither'
        with torch.cuda.device(self.device):
            off = _as_dev(np.asarray(offsets, dtype=np.int8) if isinstance(offsets, np.ndarray) else offsets, self.device)
            assert off.dtype == torch.int8 and tuple(off.shape) == (4, self.H, self.W, 2), 'offsets must be int8 [4,H,W,2]'
            if isinstance(ms, np.ndarray):
                a, b = np.ascontiguousarray(ms), np.ascontiguousarray(pan)
                assert a.shape == (self.H, self.W, 4) and b.shape == (4 * self.H, 4 * self.W)
                check(lib.dmf_scene_set_mspan_ihs(self._h, a.ctypes.data_as(C.c_void_p), np_dtype_code(a), b.ctypes.data_as(C.c_void_p),
                                                  np_dtype_code(b), 0, _ptr(off), _ptr(ms_range), _ptr(pan_range), _stream()))
                torch.cuda.current_stream().synchronize()
            else:
                a, b = ms.contiguous(), pan.contiguous()
                assert tuple(a.shape) == (self.H, self.W, 4) and tuple(b.shape) == (4 * self.H, 4 * self.W)
                check(lib.dmf_scene_set_mspan_ihs(self._h, _ptr(a), code[a.dtype], _ptr(b), code[b.dtype], 1 if a.is_cuda else 0,
                                                  _ptr(off), _ptr(ms_range), _ptr(pan_range), _stream()))
                if not a.is_cuda:
                    torch.cuda.current_stream().synchronize()
        self.has_mspan = True

    def export(self, which):
        """0 -> MS [Hp,Wp,4], 1 -> PAN [H4p,W4p], 2 -> MSPAN; float32 CUDA tensors."""
        shape = (self.Hp, self.Wp, 4) if which == 0 else (self.H4p, self.W4p)
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.dmf_scene_export(self._h, which, _ptr(out), _stream()))
        return out

    def gather(self, flat_idx, tri=False, want_target=None):
        """K1: flat_idx int64 (tensor / ndarray / list) -> (ms [N,4,p,p], pan [N,1,4p,4p][, mspan], target [N])."""
        idx = torch.as_tensor(flat_idx, dtype=torch.int64)
        if not idx.is_cuda and idx.numel():          # host index lists (the loaders') are range-checked; the kernels clamp
            lo, hi = int(idx.min()), int(idx.max())
            if lo < 0 or hi >= self.H * self.W:
                raise IndexError('gather: flat pixel index %d outside the %d x %d scene' % (lo if lo < 0 else hi, self.H, self.W))
        idx = idx.to(self.device)
        N, p = idx.numel(), self.p
        want_target = self.has_labels if want_target is None else want_target
        ms = torch.empty((N, 4, p, p), dtype=torch.float32, device=self.device)
        pan = torch.empty((N, 1, 4 * p, 4 * p), dtype=torch.float32, device=self.device)
        mspan = torch.empty_like(pan) if tri else None
        tgt = torch.empty((N,), dtype=torch.float32, device=self.device) if want_target else None
        with torch.cuda.device(self.device):
            check(lib.dmf_gather(self._h, _ptr(idx), N, _ptr(ms), _ptr(pan), _ptr(mspan), _ptr(tgt), _stream()))
        return (ms, pan, mspan, tgt) if tri else (ms, pan, tgt)

    def close(self):
        if self._h:
            lib.dmf_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def band_slice(H, p, r0, r1):
    """Scene rows [s0, s1) a rank needs to classify the anchors of rows [r0, r1) from a band-local Scene: the band, the next
    band's first p-1 rows (every window reaches p-1 rows down) and, for the last band, enough rows above it for the bottom
    reflect padding to be the band's own (BORDER_REFLECT_101 of function/function.py:104-110 reaches p-1 rows up from the
    scene's last row).  The anchors are rows [r0 - s0, r1 - s0) of the band scene."""
    s1 = min(H, r1 + p - 1)
    s0 = r0 if s1 < H else max(0, min(r0, H - p))
    return s0, s1


def raster_minmax(t):
    """{min, max} of a CUDA raster as a float64 CUDA tensor [2] (asynchronous; all-reduce it across the bands of a scene)."""
    code = {torch.uint8: U8, torch.int16: U16, torch.uint16: U16, torch.float32: F32, torch.float64: F64}
    assert t.is_cuda and t.is_contiguous()
    out = torch.empty((2,), dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        check(lib.dmf_raster_minmax(_ptr(t), code[t.dtype], t.numel(), _ptr(out), _stream()))
    return out


def normalize_pad(array, P, out_dtype=np.float64, device='cuda:0'):
    """data_padding()'s arithmetic for one raster on the GPU; returns a CUDA tensor."""
    _require_cuda()
    a = np.ascontiguousarray(array)
    H, W = a.shape[0], a.shape[1]
    bands = a.shape[2] if a.ndim == 3 else 1
    with torch.cuda.device(device):
        src = _as_dev(a, device)
        shape = (H + P - 1, W + P - 1) + ((bands,) if a.ndim == 3 else ())
        out = torch.empty(shape, dtype=torch.float64 if out_dtype == np.float64 else torch.float32, device=device)
        check(lib.dmf_normalize_pad(_ptr(src), np_dtype_code(a), H, W, bands, P, _ptr(out),
                                    F64 if out_dtype == np.float64 else F32, _stream()))
    return out


def ihs_tran(ms, pan, offsets, device='cuda:0'):
    """K2: ms f64 [H,W,4], pan f64 [4H,4W], offsets int8 [4,H,W,2] -> f64 CUDA tensor [4H,4W]."""
    _require_cuda()
    H, W = ms.shape[0], ms.shape[1]
    with torch.cuda.device(device):
        a = _as_dev(np.asarray(ms, dtype=np.float64) if isinstance(ms, np.ndarray) else ms.double(), device)
        b = _as_dev(np.asarray(pan, dtype=np.float64) if isinstance(pan, np.ndarray) else pan.double(), device)
        o = _as_dev(np.asarray(offsets, dtype=np.int8) if isinstance(offsets, np.ndarray) else offsets, device)
        out = torch.empty((4 * H, 4 * W), dtype=torch.float64, device=device)
        check(lib.dmf_ihs_tran(_ptr(a), _ptr(b), _ptr(o), _ptr(out), H, W, _stream()))
    return out


def pan2ms(pan, device='cuda:0'):
    """K2: pan [4H,4W] (u8/u16/f32/f64 ndarray) -> f64 CUDA tensor [H,W,4]."""
    _require_cuda()
    a = np.ascontiguousarray(pan)
    H4, W4 = a.shape
    with torch.cuda.device(device):
        src = _as_dev(a, device)
        out = torch.empty((H4 // 4, W4 // 4, 4), dtype=torch.float64, device=device)
        check(lib.dmf_pan2ms(_ptr(src), np_dtype_code(a), H4, W4, _ptr(out), _stream()))
    return out


def argmax_confusion(logits, target, C_, cm=None, want_pred=True):
    """K4: logits f32 [N,C] CUDA, target f32 or u8 [N] CUDA -> (pred int64 [N] | None, cm int64 [C,C])."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and logits.shape[1] == C_
    logits = logits.contiguous()
    N = logits.shape[0]
    if cm is None:
        cm = torch.zeros((C_, C_), dtype=torch.int64, device=logits.device)
    pred = torch.empty((N,), dtype=torch.int64, device=logits.device) if want_pred else None
    tcode = F32
    if target is not None:
        target = target.contiguous()
        tcode = {torch.float32: F32, torch.uint8: U8}[target.dtype]
    with torch.cuda.device(logits.device):
        check(lib.dmf_argmax_confusion(_ptr(logits), _ptr(target), tcode, N, C_, _ptr(pred),
                                       _ptr(cm) if target is not None else C.c_void_p(0), _stream()))
    return pred, cm


def confusion_at(pred_map, label_map, flat_idx, C_, cm=None):
    """cm[pred_map[k]][label_map[k]] += 1 for k in flat_idx (int64 CUDA tensor; None = every pixel); u8 maps on the device."""
    assert pred_map.is_cuda and pred_map.dtype == torch.uint8 and label_map.dtype == torch.uint8 and label_map.is_cuda
    pred_map, label_map = pred_map.contiguous(), label_map.contiguous()
    if cm is None:
        cm = torch.zeros((C_, C_), dtype=torch.int64, device=pred_map.device)
    idx = None if flat_idx is None else torch.as_tensor(flat_idx, dtype=torch.int64)
    if idx is not None and not idx.is_cuda and idx.numel():          # host index lists are range-checked (the kernel does not)
        lo, hi = int(idx.min()), int(idx.max())
        if lo < 0 or hi >= pred_map.numel():
            raise IndexError('confusion_at: flat pixel index %d outside a map of %d pixels' % (lo if lo < 0 else hi, pred_map.numel()))
    assert label_map.numel() == pred_map.numel(), 'prediction and label maps differ in size'
    idx = None if idx is None else idx.to(pred_map.device).contiguous()
    n = pred_map.numel() if idx is None else idx.numel()
    with torch.cuda.device(pred_map.device):
        check(lib.dmf_confusion_at(_ptr(pred_map), _ptr(label_map), _ptr(idx), n, C_, _ptr(cm), _stream()))
    return cm


def scatter_labels(label_map, x, y, pred):
    """K5a: label_map u8 [H,W] CUDA; x, y, pred int64 [N] (moved to the device if needed)."""
    dev = label_map.device
    x, y, pred = (torch.as_tensor(v, dtype=torch.int64).to(dev).contiguous() for v in (x, y, pred))
    with torch.cuda.device(dev):
        check(lib.dmf_scatter_labels(_ptr(x), _ptr(y), _ptr(pred), x.numel(), _ptr(label_map), label_map.shape[1], _stream()))
    return label_map


def paint_labels(label_map, colors):
    """K5b: label_map u8 [H,W] CUDA, colors [[r,g,b],...] -> u8 [H,W,3] CUDA."""
    pal = np.ascontiguousarray(np.asarray(colors, dtype=np.uint8))
    lm = label_map.contiguous()
    out = torch.empty(tuple(lm.shape) + (3,), dtype=torch.uint8, device=lm.device)
    with torch.cuda.device(lm.device):
        check(lib.dmf_paint_labels(_ptr(lm), lm.numel(), pal.ctypes.data_as(C.c_void_p), pal.shape[0], _ptr(out), _stream()))
    return out


class NetHandle:
    """GMFNet weights packed for the sm_100a kernels + activation workspace for `max_batch` patches."""

    def __init__(self, p, num_classes, max_batch=16384, device='cuda:0'):
        _require_cuda()
        self.p, self.C, self.max_batch, self.device = p, num_classes, max_batch, device
        self._h = C.c_void_p()
        check(lib.dmf_net_create(C.byref(self._h), p, num_classes, max_batch))
        self.flops_per_patch = int(lib.dmf_net_flops_per_patch(self._h))
        self.dense = True

    def load_state_dict(self, sd):
        with torch.cuda.device(self.device):
            for k, v in sd.items():
                if not torch.is_floating_point(v):
                    continue                      # num_batches_tracked
                a = np.ascontiguousarray(v.detach().to('cpu', torch.float32).numpy())
                check(lib.dmf_net_load_param(self._h, k.encode(), a.ctypes.data_as(C.c_void_p), a.size))
            check(lib.dmf_net_finalize(self._h, _stream()))

    def forward_patches(self, ms, pan):
        assert ms.is_cuda and pan.is_cuda and ms.dtype == torch.float32 and pan.dtype == torch.float32
        ms, pan = ms.contiguous(), pan.contiguous()
        N = ms.shape[0]
        assert tuple(ms.shape[1:]) == (4, self.p, self.p) and tuple(pan.shape) == (N, 1, 4 * self.p, 4 * self.p)
        out = torch.empty((N, self.C), dtype=torch.float32, device=ms.device)
        with torch.cuda.device(ms.device):
            check(lib.dmf_net_forward_patches(self._h, _ptr(ms), _ptr(pan), N, _ptr(out), _stream()))
        return out

    def forward_scene(self, scene, flat_idx=None, first=0, count=None, want_logits=True, want_pred=False, cm=None,
                      pred_map=None):
        idx = None
        if flat_idx is not None:
            idx = torch.as_tensor(flat_idx, dtype=torch.int64).to(self.device).contiguous()
            count = idx.numel()
        logits = torch.empty((count, self.C), dtype=torch.float32, device=self.device) if want_logits else None
        pred = torch.empty((count,), dtype=torch.uint8, device=self.device) if want_pred else None
        with torch.cuda.device(self.device):
            check(lib.dmf_net_forward_scene(self._h, scene._h, _ptr(idx), first, count, _ptr(logits), _ptr(pred),
                                            _ptr(cm), _ptr(pred_map), _stream()))
        return logits, pred

    def infer_scene(self, scene, row0=0, row1=None, pred_map=None, cm=None, want_logits=False):
        """Fused whole-band inference: returns (pred_map u8 [H,W], cm int64 [C,C]) (+ logits [(row1-row0)*W, C] when asked).
        Runs the scene-dense maps (csrc/dense.cu) unless set_dense(False) selected the per-patch kernels."""
        row1 = scene.H if row1 is None else row1
        if pred_map is None:
            pred_map = torch.zeros((scene.H, scene.W), dtype=torch.uint8, device=self.device)
        if cm is None and scene.has_labels:
            cm = torch.zeros((self.C, self.C), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            if want_logits:
                n = (row1 - row0) * scene.W
                logits = torch.empty((n, self.C), dtype=torch.float32, device=self.device)
                if self.dense:
                    check(lib.dmf_infer_scene_dense(self._h, scene._h, row0, row1, _ptr(logits), _ptr(pred_map), _ptr(cm), _stream()))
                else:
                    check(lib.dmf_net_forward_scene(self._h, scene._h, None, row0 * scene.W, n, _ptr(logits), None, _ptr(cm),
                                                    _ptr(pred_map), _stream()))
                return pred_map, cm, logits
            check(lib.dmf_infer_scene(self._h, scene._h, row0, row1, _ptr(pred_map), _ptr(cm), _stream()))
        return pred_map, cm

    def set_dense(self, enabled=True, band_rows=0):
        """Whole-scene inference mode: True = scene-dense maps (every layer once per scene position and border class),
        False = per-patch kernels.  band_rows = anchor rows per pass of the dense path (0 keeps the current value)."""
        check(lib.dmf_net_set_dense(self._h, 1 if enabled else 0, int(band_rows)))
        self.dense = bool(enabled)

    def set_pan_source(self, use_mspan):
        """Scene inference reads the scene's IHS product (Scene.set_mspan) instead of the PAN raster (IHS-input models)."""
        check(lib.dmf_net_set_pan_source(self._h, 1 if use_mspan else 0))

    def get_dense_timing(self, reset=True):
        buf = (C.c_float * 12)()
        check(lib.dmf_net_get_dense_timing(self._h, buf, 1 if reset else 0))
        names = ['ms_stem_maps', 'conv_ms2', 'pool_ms2', 'pan_stem_maps', 'conv_pan2', 'pool_pan2', 'conv_pan3', 'pool_pan3',
                 'conv_fuse', 'head', '_', 'total']      # conv_* include the fused pooling (pool_* stay 0); conv_fuse includes the row sums
        return {k: float(v) for k, v in zip(names, buf) if k != '_'}

    def dense_buffer(self, name):
        """test hook: a dense-path map as a flat tensor aliasing the library's workspace (fp16; "S" holds the row MEANS of F), and
        (rows, cols) of the MS grid"""
        ptr, nbytes, dims = C.c_void_p(), C.c_int64(), (C.c_int32 * 2)()
        check(lib.dmf_net_dense_buffer(self._h, name.encode(), C.byref(ptr), C.byref(nbytes), dims))
        t = _from_ptr(ptr.value, nbytes.value, self.device)
        return t.view(torch.float16), (int(dims[0]), int(dims[1]))

    def set_timing(self, on):
        check(lib.dmf_net_set_timing(self._h, 1 if on else 0))

    def get_timing(self):
        buf = (C.c_float * 8)()
        check(lib.dmf_net_get_timing(self._h, buf))
        names = ['stem_ms', 'conv_ms2', 'stem_pan', 'conv_pan2', 'conv_pan3', 'conv_fuse', 'head', 'total']
        return dict(zip(names, [float(v) for v in buf]))

    def debug_layer(self, layer, impl, x, out_shape, out=None):
        if out is None:
            out = torch.zeros(out_shape, dtype=torch.float16, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.dmf_net_debug_layer(self._h, layer, impl, _ptr(x.contiguous()), _ptr(out), x.shape[0], _stream()))
        return out

    def debug_stem(self, which, patches, out_shape):
        out = torch.zeros(out_shape, dtype=torch.float16, device=patches.device)
        with torch.cuda.device(patches.device):
            check(lib.dmf_net_debug_stem(self._h, which, _ptr(patches.contiguous()), _ptr(out), patches.shape[0], _stream()))
        return out

    def close(self):
        if self._h:
            lib.dmf_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def softmax_ce(logits, target, want_grad=True):
    """CrossEntropyLoss(reduction='mean') on the device: (loss f32 [], dlogits f32 [N,C] | None).
    target: float32 labels (as the loaders deliver them) or int64."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2
    logits, target = logits.contiguous(), target.contiguous()
    assert target.dtype in (torch.float32, torch.int64) and target.numel() == logits.shape[0]
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    dl = torch.empty_like(logits) if want_grad else None
    with torch.cuda.device(logits.device):
        check(lib.dmf_softmax_ce(_ptr(logits), _ptr(target), 1 if target.dtype == torch.int64 else 0, logits.shape[0],
                                 logits.shape[1], _ptr(loss), _ptr(dl), _stream()))
    return loss, dl


class TrainHandle:
    """Native training step for a GMFNet nn.Module (solver/mainsolver.py:49-55).

    The module's parameters are re-seated as views of ONE flat fp32 tensor (``flat``) with a matching flat
    gradient tensor (``flat_grad``; every ``p.grad`` is a view of it), so the data-parallel gradient all-reduce is
    one collective and Adam is one kernel.  The library reads / writes those tensors in place."""

    def __init__(self, module, p, num_classes, max_batch=512, device='cuda:0'):
        _require_cuda()
        self.p, self.C, self.max_batch, self.device = p, num_classes, max_batch, str(device)
        self._h = C.c_void_p()
        check(lib.dmf_train_create(C.byref(self._h), p, num_classes, max_batch))
        self.module = module
        self._bind()

    def _bind(self):
        named = list(self.module.named_parameters())
        total = sum(q.numel() for _, q in named)
        dev = named[0][1].device
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.forward_id = 0
        off = 0
        self._views = []
        with torch.cuda.device(dev):
            for name, q in named:
                k = q.numel()
                self.flat[off:off + k].copy_(q.data.reshape(-1))
                q.data = self.flat[off:off + k].view(q.shape)
                q.grad = self.flat_grad[off:off + k].view(q.shape)
                self._views.append((q, off, k))
                check(lib.dmf_train_bind(self._h, name.encode(), _ptr(q.data), _ptr(q.grad), k))
                off += k
            for name, b in self.module.named_buffers():
                check(lib.dmf_train_bind(self._h, name.encode(), _ptr(b), C.c_void_p(0), b.numel()))
            check(lib.dmf_train_finalize(self._h))
        self._key = self.key(self.module)

    @staticmethod
    def key(module):
        return tuple(t.data_ptr() for t in list(module.parameters()) + list(module.buffers()))

    def stale(self):
        return self._key != self.key(self.module)

    def reseat_grads(self):
        """Make every p.grad a view of flat_grad again (optimizer.zero_grad(set_to_none=True) drops them)."""
        for q, off, k in self._views:
            if q.grad is None or q.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                q.grad = self.flat_grad[off:off + k].view(q.shape)

    def forward(self, ms, pan):
        assert ms.is_cuda and pan.is_cuda
        ms, pan = ms.float().contiguous(), pan.float().contiguous()
        N = ms.shape[0]
        assert tuple(ms.shape[1:]) == (4, self.p, self.p) and tuple(pan.shape) == (N, 1, 4 * self.p, 4 * self.p)
        out = torch.empty((N, self.C), dtype=torch.float32, device=ms.device)
        self._keep = (ms, pan)               # the PAN stem's weight gradient re-reads the input in backward
        with torch.cuda.device(ms.device):
            check(lib.dmf_train_forward(self._h, _ptr(ms), _ptr(pan), N, _ptr(out), _stream()))
        return out

    def backward(self, dlogits):
        dlogits = dlogits.float().contiguous()
        with torch.cuda.device(dlogits.device):
            check(lib.dmf_train_backward(self._h, _ptr(dlogits), _stream()))

    def step_patches(self, ms, pan, target, loss_out=None):
        """zero grads -> forward -> CrossEntropyLoss(mean) -> backward; returns the loss (0-d CUDA tensor)."""
        ms, pan, target = ms.float().contiguous(), pan.float().contiguous(), target.contiguous()
        assert target.dtype in (torch.float32, torch.int64)
        loss = loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=ms.device)
        self._keep = (ms, pan)
        self.flat_grad.zero_()
        with torch.cuda.device(ms.device):
            check(lib.dmf_train_step_patches(self._h, _ptr(ms), _ptr(pan), _ptr(target), 1 if target.dtype == torch.int64 else 0,
                                             ms.shape[0], _ptr(loss), _stream()))
        return loss

    def step_scene(self, scene, flat_idx, use_mspan=False, loss_out=None):
        """Same, with the batch cropped from the device scene by flat pixel index (no patches leave the library)."""
        idx = torch.as_tensor(flat_idx, dtype=torch.int64).to(self.device).contiguous()
        loss = loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=self.device)
        self.flat_grad.zero_()
        with torch.cuda.device(self.device):
            check(lib.dmf_train_step_scene(self._h, scene._h, _ptr(idx), idx.numel(), 1 if use_mspan else 0, _ptr(loss), _stream()))
        return loss

    def buffer(self, name, dtype, shape, alias=False):
        """Test hook: a copy (or, with alias=True, a writable view) of an internal activation / gradient buffer."""
        ptr, nbytes = C.c_void_p(), C.c_int64()
        check(lib.dmf_train_buffer(self._h, name.encode(), C.byref(ptr), C.byref(nbytes)))
        n = int(np.prod(shape))
        item = torch.empty((), dtype=dtype).element_size()
        assert n * item <= nbytes.value, 'buffer %s holds %d bytes' % (name, nbytes.value)
        view = _from_ptr(ptr.value, n * item, self.device).view(dtype)[:n].view(shape)
        return view if alias else view.clone()

    def debug_op(self, op, layer, N):
        """Test hook: 'pack' | 'fwd' | 'wgrad' | 'dgrad' of one layer on the internal buffers."""
        code = {'pack': 0, 'fwd': 1, 'wgrad': 2, 'dgrad': 3}[op]
        lay = {'ms1': 0, 'ms2': 1, 'pan1': 2, 'pan2': 3, 'pan3': 4, 'fuse': 5}[layer]
        with torch.cuda.device(self.device):
            check(lib.dmf_train_debug_op(self._h, code, lay, N, _stream()))

    def set_debug(self, swap_lbo_sbo):
        check(lib.dmf_train_set_debug(self._h, int(swap_lbo_sbo)))

    def close(self):
        if self._h:
            lib.dmf_train_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ScenePipeline:
    """End-to-end whole-scene classification of a STREAM of same-sized scenes from pinned host rasters, software-pipelined:
    while the kernels classify scene i on the caller's stream, scene i+1 is already being uploaded on a copy stream (H2D of
    this rank's rows, its min / max, and under torch.distributed the 4-value range all-reduce), so the copy engine and the
    SMs work at the same time.  Per scene: H2D (band + p-1 halo rows of MS, PAN, labels) -> to_tensor range of the WHOLE scene
    (function/function.py:120-124) -> normalise + reflect-pad -> dmf_infer_scene on the rank's row band -> int64 C x C
    all-reduce -> D2H of the label band and the matrix (solver/mainsolver.py:104-141, 167-185 for every pixel of the scene).

        pipe = ScenePipeline(net_handle, H, W, p, r0, r1)          # rows [r0, r1) of the scene are this rank's anchors
        t = pipe.submit(ms_pin, pan_pin, label_pin)                # whole-scene pinned tensors (int16 views of uint16 are fine)
        pred_band, cm = pipe.result(t)                             # pinned host tensors, valid until `depth` more submits
    """

    class _Slot:
        pass

    def __init__(self, net, H, W, p, r0, r1, ms_dtype=torch.int16, pan_dtype=torch.int16, depth=2, device=None, group=None):
        import torch.distributed as dist
        self.net, self.H, self.W, self.p, self.r0, self.r1 = net, H, W, p, r0, r1
        self.device = device or net.device
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.group, self.dist = group, dist
        self.s0, self.s1 = band_slice(H, p, r0, r1)
        Hb, C_ = self.s1 - self.s0, net.C
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots, self.n = [], 0
        with torch.cuda.device(self.device):
            for _ in range(depth):
                s = ScenePipeline._Slot()
                s.ms = torch.empty((Hb, W, 4), dtype=ms_dtype, device=self.device)
                s.pan = torch.empty((4 * Hb, 4 * W), dtype=pan_dtype, device=self.device)
                s.lab = torch.empty((Hb, W), dtype=torch.uint8, device=self.device)
                s.scene = Scene.from_raw(s.ms, s.pan, p, self.device)          # allocates the padded rasters once (contents replaced per scene)
                s.pm = torch.zeros((Hb, W), dtype=torch.uint8, device=self.device)
                s.cm = torch.zeros((C_, C_), dtype=torch.int64, device=self.device)
                s.pm_host = torch.empty((r1 - r0, W), dtype=torch.uint8).pin_memory()
                s.cm_host = torch.empty((C_, C_), dtype=torch.int64).pin_memory()
                s.uploaded, s.consumed, s.done = (torch.cuda.Event() for _ in range(3))
                s.rng = None
                s.ranges = None
                s.consumed.record()
                self.slots.append(s)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in (self.slots[0].ms, self.slots[0].pan, self.slots[0].lab))
        self.d2h_bytes = (r1 - r0) * W + C_ * C_ * 8

    def submit(self, ms_pin, pan_pin, lab_pin, ms_range=None, pan_range=None):
        """Queue one scene.  to_tensor (function/function.py:120-124) normalises with the range of the WHOLE raster: under
        torch.distributed the ranks' bands cover the scene and their ranges are all-reduced; a single process that classifies only a
        part of the scene (r0 > 0 or r1 < H) uploads only that part and must be GIVEN the ranges (ms_range / pan_range = (min, max))."""
        if self.world == 1 and (self.r0 != 0 or self.r1 != self.H) and (ms_range is None or pan_range is None):
            raise ValueError('ScenePipeline: a partial band in a single process needs ms_range / pan_range of the whole rasters')
        s = self.slots[self.n % len(self.slots)]
        self.n += 1
        s0, s1, r0, r1 = self.s0, self.s1, self.r0, self.r1
        with torch.cuda.device(self.device):
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(s.consumed)               # the slot's raw buffers were read by its previous scene
                s.ms.copy_(ms_pin[s0:s1], non_blocking=True)
                s.pan.copy_(pan_pin[4 * s0:4 * s1], non_blocking=True)
                s.lab.copy_(lab_pin[s0:s1], non_blocking=True)
                if self.world > 1:
                    a = raster_minmax(s.ms[r0 - s0:r1 - s0])
                    b = raster_minmax(s.pan[4 * (r0 - s0):4 * (r1 - s0)])
                    s.rng = torch.stack([a[0], -a[1], b[0], -b[1]])
                    self.dist.all_reduce(s.rng, op=self.dist.ReduceOp.MIN, group=self.group)
                    s.ranges = (torch.stack([s.rng[0], -s.rng[1]]), torch.stack([s.rng[2], -s.rng[3]]))
                elif ms_range is not None:
                    s.ranges = (torch.tensor([float(ms_range[0]), float(ms_range[1])], dtype=torch.float64).to(self.device, non_blocking=True),
                                torch.tensor([float(pan_range[0]), float(pan_range[1])], dtype=torch.float64).to(self.device, non_blocking=True))
                else:
                    s.ranges = None
                s.uploaded.record()
            cur = torch.cuda.current_stream()
            cur.wait_event(s.uploaded)
            if s.ranges is not None:
                s.scene.update_raw(s.ms, s.pan, *s.ranges)
            else:
                s.scene.update_raw(s.ms, s.pan)
            s.scene.set_labels(s.lab)
            s.consumed.record()
            s.cm.zero_()
            self.net.infer_scene(s.scene, r0 - s0, r1 - s0, pred_map=s.pm, cm=s.cm)
            if self.world > 1:
                self.dist.all_reduce(s.cm, group=self.group)
            s.pm_host.copy_(s.pm[r0 - s0:r1 - s0], non_blocking=True)
            s.cm_host.copy_(s.cm, non_blocking=True)
            s.done.record()
        return s

    def result(self, ticket):
        ticket.done.synchronize()
        return ticket.pm_host, ticket.cm_host


def _from_ptr(ptr, nbytes, device):
    """uint8 CUDA tensor aliasing `nbytes` of device memory at `ptr` (no ownership)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 3, 'strides': None}
    return torch.as_tensor(h, device=device)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps; no weight decay / amsgrad) with the update done by dmf_adam_step.
    When the parameters are contiguous slices of one flat tensor (TrainHandle re-seats them that way) the whole
    model is ONE kernel launch; otherwise one launch per tensor.  State layout follows torch's names
    (step / exp_avg / exp_avg_sq) so checkpoints stay readable (utils/utils.py:82-88)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._flat = None

    def _flat_view(self, group):
        """(param_flat, grad_flat) if the group's tensors tile one contiguous block in order, else None."""
        ps = [q for q in group['params'] if q.requires_grad]
        if not ps or any(q.grad is None for q in ps):
            return None
        p0, g0 = ps[0].data_ptr(), ps[0].grad.data_ptr()
        off = 0
        for q in ps:
            if q.data_ptr() != p0 + 4 * off or q.grad.data_ptr() != g0 + 4 * off or q.dtype != torch.float32:
                return None
            off += q.numel()
        key = (p0, g0, off)
        if self._flat is None or self._flat[0] != key:
            self._flat = (key, _from_ptr(p0, 4 * off, ps[0].device).view(torch.float32), _from_ptr(g0, 4 * off, ps[0].device).view(torch.float32))
        return self._flat[1], self._flat[2]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            lr, (b1, b2), eps = float(group['lr']), group['betas'], group['eps']
            flat = self._flat_view(group)
            # flat mode: one state entry (whole-model exp_avg / exp_avg_sq) stored under the group's first parameter
            items = [(group['params'][0], flat[0], flat[1])] if flat is not None else \
                    [(q, q.data, q.grad) for q in group['params'] if q.grad is not None]
            for key, pt, gt in items:
                st = self.state[key]
                if 'step' not in st or st['exp_avg'].numel() != pt.numel():
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(pt)
                    st['exp_avg_sq'] = torch.zeros_like(pt)
                st['step'] += 1
                assert pt.is_cuda and pt.is_contiguous() and gt.is_contiguous(), 'FusedAdam needs contiguous CUDA tensors (no CPU path)'
                with torch.cuda.device(pt.device):
                    check(lib.dmf_adam_step(_ptr(pt), _ptr(gt), _ptr(st['exp_avg']), _ptr(st['exp_avg_sq']), pt.numel(),
                                            lr, b1, b2, eps, st['step'], _stream()))
        return loss

    def zero_grad(self, set_to_none=False):
        """Gradients stay allocated (they are views of the flat buffer the library writes into)."""
        for group in self.param_groups:
            flat = self._flat_view(group)
            if flat is not None:
                flat[1].zero_()
            else:
                for q in group['params']:
                    if q.grad is not None:
                        q.grad.zero_()
