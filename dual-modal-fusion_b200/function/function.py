"""Data preparation, same public names as the reference's function/function.py.

to_tensor / data_padding run on the GPU (dmf_normalize_pad) and return exactly the float64 (or, for
float32 rasters, float32) arrays the reference returns; split_data_old / split_data are vectorised on
the host and return identical index lists (order included).  File readers keep their signatures.
"""
import os

import numpy as np

import dmf


def read_tif(cfg, mode):
    """reference: function/function.py:34-43 (libtiff: bands in FILE order).  Without libtiff the raster is read with OpenCV, whose
    TIFF decoder hands 3- / 4-sample images back as BGR(A): the first three bands are swapped back so that the array is in file
    order like libtiff's.  That restoration is announced once (band order matters to trained weights) and can be overridden with
    cfg['b200']['tif_band_order'] = [i0, i1, i2, i3] (indices into what OpenCV returned)."""
    if mode not in ('ms', 'pan'):
        raise ValueError("mode")
    filename = cfg['data_address'] + ('ms4.tif' if mode == 'ms' else 'pan.tif')
    try:
        from libtiff import TIFF
        return TIFF.open(filename, mode='r').read_image()
    except ImportError:
        pass
    import warnings
    import cv2
    image = cv2.imread(filename, cv2.IMREAD_UNCHANGED)
    if image is None:
        raise FileNotFoundError(filename)
    if image.ndim == 3:
        b200 = cfg.get('b200') if isinstance(cfg.get('b200'), dict) else {}
        order = b200.get('tif_band_order') or ([2, 1, 0] + list(range(3, image.shape[2])) if image.shape[2] >= 3 else list(range(image.shape[2])))
        warnings.warn('read_tif: libtiff is not installed; %s read with OpenCV and its bands taken in the order %s '
                      '(OpenCV decodes multi-sample TIFFs as BGR[A]; set b200.tif_band_order to override)' % (filename, order), stacklevel=2)
        image = image[:, :, list(order)]
    return np.ascontiguousarray(image)


def label_mat2np(cfg):
    """reference: function/function.py:11-17 (needs h5py; not on the hot path)."""
    import h5py
    path = cfg['data_address']
    label = np.array(h5py.File(path + 'label.mat')['label'], dtype='uint8')
    np.save(path + 'label.npy', np.transpose(label))


def to_tensor(image):
    """Global min-max normalisation over the whole raster (reference: function/function.py:120-124)."""
    image = np.asarray(image)
    out = dmf.normalize_pad(image, 1).cpu().numpy()
    return out.astype(np.float32) if image.dtype == np.float32 else out


def data_padding(array, cfg, mode=None):
    """Normalise, then reflect-101 pad bottom/right by P-1 (reference: function/function.py:99-117);
    P = patch_size for the 3-D MS raster, 4*patch_size for the 2-D PAN raster."""
    array = np.asarray(array)
    P = cfg['patch_size'] if array.ndim == 3 else cfg['patch_size'] * 4
    out = dmf.normalize_pad(array, P).cpu().numpy()
    return out.astype(np.float32) if array.dtype == np.float32 else out


def data_show(matrix):
    elements, counts = np.unique(matrix, return_counts=True)
    rows, cols = np.shape(matrix)
    print("labels {} counts {} rows {} cols {} classes {}".format(elements, counts, rows, cols, len(elements) - 1))


def _enumerate_pixels(label, size):
    H, W = int(size[0]), int(size[1])
    lab = np.asarray(label)[:H, :W]
    rows = np.repeat(np.arange(H, dtype=np.float64), W).reshape(-1, 1)
    cols = np.tile(np.arange(W, dtype=np.float64), H).reshape(-1, 1)
    return [rows, cols, lab.reshape(-1, 1).astype(np.float64)], lab.reshape(-1)


def split_data_old(label, cfg):
    """Row-major pixel enumeration t = i*W + j and the unlabelled / labelled index lists
    (reference: function/function.py:149-169, a 15 s Python loop at 2001x2101)."""
    the_matrix, flat = _enumerate_pixels(label, cfg['DATA_DICT'][cfg['data_city']]['size'])
    matrix_ = [np.flatnonzero(flat == 0).tolist(), np.flatnonzero(flat != 0).tolist()]
    for i in range(2):
        print("label-{} index set size {}".format(i, len(matrix_[i])))
    return the_matrix, matrix_


def split_data(train_label, test_label, label, cfg):
    """reference: function/function.py:172-194: [neither, train-labelled, test-labelled (and not train)]."""
    size = cfg['DATA_DICT'][cfg['data_city']]['size']
    the_matrix, _ = _enumerate_pixels(label, size)
    H, W = int(size[0]), int(size[1])
    tr = np.asarray(train_label)[:H, :W].reshape(-1) != 0
    te = np.asarray(test_label)[:H, :W].reshape(-1) != 0
    matrix_ = [np.flatnonzero(~tr & ~te).tolist(), np.flatnonzero(tr).tolist(), np.flatnonzero(~tr & te).tolist()]
    for i in range(3):
        print("label-{} index set size {}".format(i, len(matrix_[i])))
    return the_matrix, matrix_
