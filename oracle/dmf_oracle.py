"""ORACLE — CPU restatement of the reference's per-pixel scene-classification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under dual-modal-fusion_b200/ may import this module; only
tests/, bench.py's ``cpu_baseline`` / ``--impl reference`` legs and __graft_entry__.smoke()
use it, and only as the checker.

Every function restates one reference function in vectorised numpy and cites the reference
file:line it follows (paths relative to the reference repo root).  The restatement is pinned:
tests/test_oracle_golden.py checks each function bit-for-bit against tests/golden/*.npz, which
tests/golden/make_golden.py produced by importing and running the reference's own Python
functions in the authoring container (the reference ships no tests or fixtures of its own,
SURVEY.md section 4).  The network arithmetic is NOT pinned by the reference (its model/ package
was never committed); see oracle/gmfnet_ref.py.
"""
import random as _pyrandom

import numpy as np

# --------------------------------------------------------------------------------------
# a1 / a2: normalisation and padding                         function/function.py:99-124
# --------------------------------------------------------------------------------------


def to_tensor(image):
    """Global min-max normalisation over the WHOLE array (all bands share one min/max).

    function/function.py:120-124.  For integer rasters numpy keeps ``image - min`` in the
    integer dtype and true-divides into float64; float32 rasters stay float32.
    """
    hi = np.max(image)
    lo = np.min(image)
    return (image - lo) / (hi - lo)


def reflect101_index(i, n):
    """cv2.BORDER_REFLECT_101 source index for a coordinate i >= 0 on an axis of length n that is
    padded on the far side only: the pattern 0..n-1, n-2..1 repeats with period 2(n-1) (no edge
    duplication), so i < n -> i and n <= i < 2n-1 -> 2(n-1) - i."""
    i = np.asarray(i)
    if n == 1:
        return np.zeros_like(i)
    t = i % (2 * (n - 1))
    return np.where(t < n, t, 2 * (n - 1) - t)


def data_padding(array, patch_size):
    """function/function.py:99-117: normalise, then pad bottom/right by P-1 with
    BORDER_REFLECT_101, P = patch_size for the 3-D MS array, 4*patch_size for the 2-D PAN."""
    P = patch_size if array.ndim == 3 else 4 * patch_size
    a = to_tensor(array)
    rows = reflect101_index(np.arange(a.shape[0] + P - 1), a.shape[0])
    cols = reflect101_index(np.arange(a.shape[1] + P - 1), a.shape[1])
    return a[rows][:, cols]


# --------------------------------------------------------------------------------------
# a3: pixel index lists                                      function/function.py:149-169
# --------------------------------------------------------------------------------------


def split_data_old(label, size):
    """function/function.py:149-169.  Row-major enumeration t = i*W + j; returns
    (the_matrix, matrix_) with the_matrix = [x, y, label] as float64 (H*W, 1) columns and
    matrix_ = [flat indices with label == 0, flat indices with label != 0] as python int lists."""
    H, W = int(size[0]), int(size[1])
    lab = np.asarray(label)[:H, :W]
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    the_matrix = [ii.reshape(-1, 1).astype(np.float64),
                  jj.reshape(-1, 1).astype(np.float64),
                  lab.reshape(-1, 1).astype(np.float64)]
    flat = lab.reshape(-1)
    matrix_ = [np.flatnonzero(flat == 0).tolist(), np.flatnonzero(flat != 0).tolist()]
    return the_matrix, matrix_


# --------------------------------------------------------------------------------------
# a4 / a5: co-registered patch crop + default collate        train/dataset.py:158-188, 248-282
# --------------------------------------------------------------------------------------


def gather_dual(MS, PAN, xs, ys, p):
    """Batched restatement of dataset_dual.__getitem__ + torch default_collate
    (train/dataset.py:168-185).  (x, y) is the TOP-LEFT corner of the MS window; the PAN
    window starts at (4x, 4y).  float64 -> float32 is the one rounding on this path."""
    xs = np.asarray(xs, dtype=np.int64)
    ys = np.asarray(ys, dtype=np.int64)
    r = np.arange(p)
    ms = MS[(xs[:, None] + r)[:, :, None], (ys[:, None] + r)[:, None, :], :]       # [B,p,p,4]
    ms = np.ascontiguousarray(ms.transpose(0, 3, 1, 2)).astype(np.float32)
    R = np.arange(4 * p)
    pan = PAN[(4 * xs[:, None] + R)[:, :, None], (4 * ys[:, None] + R)[:, None, :]]
    pan = pan[:, None, :, :].astype(np.float32)
    return ms, pan


def gather_tri(MS, PAN, MSPAN, xs, ys, p):
    """dataset_tri.__getitem__ (train/dataset.py:259-279): gather_dual plus the same PAN-grid
    window cut from the IHS product MSPAN."""
    ms, pan = gather_dual(MS, PAN, xs, ys, p)
    _, mspan = gather_dual(MS, MSPAN, xs, ys, p)
    return ms, pan, mspan


# --------------------------------------------------------------------------------------
# a7 / a8: IHS                                               image_convert/IHS.py:6-54
# --------------------------------------------------------------------------------------


def draw_unpooling_offsets(H, W, bands, time, rng=None):
    """The (m, n) stream that ``unpooling`` draws (image_convert/IHS.py:25-28): loop order
    band -> row -> col, ``m = random.randint(0, time-1)`` then ``n = ...``, from Python's
    Mersenne-Twister.  Returns int8 [bands, H, W, 2]."""
    rng = rng or _pyrandom
    out = np.empty((bands, H, W, 2), dtype=np.int8)
    flat = out.reshape(-1)
    for t in range(flat.size):
        flat[t] = rng.randint(0, time - 1)
    return out


def unpooling_from_offsets(pic, offs, time):
    """image_convert/IHS.py:22-29 with the random draws supplied as a table."""
    H, W, B = pic.shape
    up = np.zeros((H * time, W * time, B))
    jj, kk = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    for i in range(B):
        up[time * jj + offs[i, :, :, 0], time * kk + offs[i, :, :, 1], i] = pic[:, :, i]
    return up


def ihs_tran_from_offsets(MS, PAN, offs):
    """image_convert/IHS.py:40-54, float64, operation order kept literally:
    I = running band mean (I*i + up_i)/(i+1); delta = PAN - I; result = up + delta;
    MSPAN = running band mean of result."""
    B = MS.shape[2]
    up = unpooling_from_offsets(MS, offs, B)
    I = up[:, :, 0]
    for i in range(1, B):
        I = (I * i + up[:, :, i]) / (i + 1)
    delta = PAN - I
    result = up + delta[:, :, None]
    out = result[:, :, 0]
    for i in range(1, B):
        out = (out * i + result[:, :, i]) / (i + 1)
    return out


def unsampling(im, scale):
    """image_convert/IHS.py:6-12: block mean.  ``np.mean`` of a (scale, scale) slice adds the
    elements in row-major order starting from the first one, then divides once."""
    H, W = im.shape
    h, w = H // scale, W // scale
    blk = im[:h * scale, :w * scale].reshape(h, scale, w, scale)
    # np.mean accumulates float32 rasters in float32, everything else here in float64
    acc_t = np.float32 if im.dtype == np.float32 else np.float64
    acc = blk[:, 0, :, 0].astype(acc_t)
    for a in range(scale):
        for b in range(scale):
            if a or b:
                acc = acc + blk[:, a, :, b].astype(acc_t)
    return (acc / acc_t(scale * scale)).astype(np.float64)


def pan2ms(pan, size):
    """image_convert/IHS.py:14-19: 2x block mean, then 2x2 space-to-depth:
    band i = p[i % 2 :: 2, i // 2 :: 2]."""
    p = unsampling(pan, 2)
    out = np.zeros(size)
    for i in range(size[2]):
        out[:, :, i] = p[i % 2::2, i // 2::2]
    return out


# --------------------------------------------------------------------------------------
# a11 / a12: argmax, confusion matrix, label map              solver/mainsolver.py:139-141, 167-185
# --------------------------------------------------------------------------------------


def argmax_first(logits):
    """``output.data.max(1, keepdim=True)[1]`` (solver/mainsolver.py:139): lowest index on ties."""
    return np.argmax(np.asarray(logits), axis=1)


def confusion(pred, target, C, M=None):
    """``M[pred][target] += 1`` per sample (solver/mainsolver.py:140-141; clean copy
    train/test.py:58-60).  float64 [C, C], rows = prediction, columns = target."""
    M = np.zeros((C, C)) if M is None else M
    np.add.at(M, (np.asarray(pred, dtype=np.int64), np.asarray(target).astype(np.int64)), 1.0)
    return M


def scatter_labels(label_map, xs, ys, pred):
    """``label_np[x][y] = pred`` (solver/mainsolver.py:171-173, 182-183)."""
    label_map[np.asarray(xs, dtype=np.int64), np.asarray(ys, dtype=np.int64)] = np.asarray(pred)
    return label_map


def paint(label_map, colors):
    """solver/mainsolver.py:186-189: label -> RGB through cfg DATA_DICT[city]['color']."""
    lut = np.asarray(colors, dtype=np.float64)
    return np.uint8(lut[np.asarray(label_map).astype(np.int64)])


# --------------------------------------------------------------------------------------
# a13 / a14: metrics                                          indicators/kappa.py:10-22, 69-84
# --------------------------------------------------------------------------------------


def kappa(matrix):
    """Cohen's kappa over the full matrix including class 0 (indicators/kappa.py:10-22).
    All partial sums are integers < 2**53, so the order of the additions cannot matter."""
    M = np.asarray(matrix, dtype=np.float64)
    n = np.sum(M)
    sum_po = 0
    sum_pe = 0
    for i in range(M.shape[1]):
        sum_po += M[i][i]
        sum_pe += np.sum(M[i, :]) * np.sum(M[:, i])
    po = sum_po / n
    pe = sum_pe / (n * n)
    return (po - pe) / (1 - pe)


def aa_oa(matrix):
    """indicators/kappa.py:69-84 without the prints.  Class 0 is skipped in the per-class
    accuracies and in the OA numerator but not in the OA denominator."""
    M = np.asarray(matrix, dtype=np.float64)
    b = np.sum(M, axis=0)
    acc, rows, c = [], [], 0
    with np.errstate(invalid='ignore', divide='ignore'):
        for i in range(1, M.shape[0]):
            a = M[i][i] / b[i]
            c += M[i][i]
            acc.append(a)
            rows.append([b[i], M[i][i], a])
        aa = np.mean(acc)
        oa = c / np.sum(b, axis=0)
        k = kappa(M)
    return [aa, oa, k, rows]


# --------------------------------------------------------------------------------------
# synthetic inputs — shared by tests, bench.py and the golden generator
# --------------------------------------------------------------------------------------


def synthetic_scene(H, W, n_classes, seed=0, label_seed=1, blocky=False):
    """MS uint16 [H,W,4] and PAN uint16 [4H,4W] uniform in [0, 2047]; label uint8 [H,W]
    uniform in [0, n_classes] (0 = unlabelled)."""
    rng = np.random.default_rng(seed)
    if blocky:
        g = 16
        base = rng.integers(200, 1800, (H // g + 1, W // g + 1, 4)).astype(np.float64)
        ms = np.kron(base, np.ones((g, g, 1)))[:H, :W] + rng.normal(0, 60, (H, W, 4))
        pan = np.kron(base.mean(2), np.ones((4 * g, 4 * g)))[:4 * H, :4 * W] + rng.normal(0, 80, (4 * H, 4 * W))
        ms = np.clip(ms, 0, 2047).astype(np.uint16)
        pan = np.clip(pan, 0, 2047).astype(np.uint16)
    else:
        ms = rng.integers(0, 2048, (H, W, 4), dtype=np.uint16)
        pan = rng.integers(0, 2048, (4 * H, 4 * W), dtype=np.uint16)
    lrng = np.random.default_rng(label_seed)
    label = lrng.integers(0, n_classes + 1, (H, W), dtype=np.uint8)
    return ms, pan, label


def synthetic_scene_structured(H, W, n_classes, seed=0, label_seed=1, cell=20):
    """A scene whose labels DEPEND on the rasters, so that a classifier can have Kappa > 0 on it (the uniform-noise scene
    above cannot: its labels are independent of the image).  A coarse grid of `cell` x `cell` regions carries a class each;
    every class has a 4-band spectral signature and a PAN texture (oriented stripes of class-specific period and amplitude
    on top of the band mean); sensor noise on both rasters.  uint16 in [0, 2047] like the 11-bit sensors.  label = class + 1,
    with ~15 % of the pixels unlabelled (0) in 8 x 8 blocks.  Returns (ms [H,W,4], pan [4H,4W], label [H,W])."""
    rng = np.random.default_rng(seed)
    gh, gw = H // cell + 2, W // cell + 2
    grid = rng.integers(0, n_classes, (gh, gw))
    oy, ox = int(rng.integers(0, cell)), int(rng.integers(0, cell))
    cmap = np.kron(grid, np.ones((cell, cell), dtype=np.int64))[oy:oy + H, ox:ox + W]
    sig = rng.uniform(350.0, 1650.0, (n_classes, 4))
    period = rng.integers(3, 9, n_classes).astype(np.float64)
    amp = rng.uniform(40.0, 220.0, n_classes)
    angle = rng.uniform(0.0, np.pi, n_classes)
    ms = sig[cmap] + rng.normal(0.0, 45.0, (H, W, 4))
    c4 = np.repeat(np.repeat(cmap, 4, axis=0), 4, axis=1)
    rr, cc = np.meshgrid(np.arange(4 * H, dtype=np.float32), np.arange(4 * W, dtype=np.float32), indexing='ij')
    phase = (rr * np.cos(angle)[c4].astype(np.float32) + cc * np.sin(angle)[c4].astype(np.float32)) / period[c4].astype(np.float32)
    pan = sig.mean(1)[c4] + amp[c4] * np.sin(2.0 * np.pi * phase) + rng.normal(0.0, 60.0, (4 * H, 4 * W)).astype(np.float32)
    ms = np.clip(np.rint(ms), 0, 2047).astype(np.uint16)
    pan = np.clip(np.rint(pan), 0, 2047).astype(np.uint16)
    lrng = np.random.default_rng(label_seed)
    holes = np.kron(lrng.random((H // 8 + 1, W // 8 + 1)) < 0.15, np.ones((8, 8), dtype=bool))[:H, :W]
    label = (cmap + 1).astype(np.uint8)
    label[holes] = 0
    return ms, pan, label

