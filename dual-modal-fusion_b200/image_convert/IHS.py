"""IHS transforms with the reference's names (image_convert/IHS.py:6-54), computed by the K2 kernels.

The reference draws unpooling's sub-pixel offsets from Python's global Mersenne Twister
(random.randint, band -> row -> col order).  draw_offsets() reproduces that stream without the
16.8 M-iteration Python loop: it lifts the generator state into numpy's MT19937, replays
getrandbits' rejection sampling vectorised, and puts the advanced state back, so `random` ends up
exactly where the reference would leave it."""
import random

import numpy as np

import dmf


def draw_offsets(H, W, bands, time, rng=random):
    """int8 [bands, H, W, 2] of (m, n) = (randint(0,time-1), randint(0,time-1)) per (band,row,col)."""
    need = bands * H * W * 2
    k = int(time).bit_length()                       # random._randbelow: getrandbits(k) until < time
    version, internal, gauss = rng.getstate()
    key0, pos0 = np.array(internal[:-1], dtype=np.uint32), int(internal[-1])
    bg = np.random.MT19937()
    bg.state = {'bit_generator': 'MT19937', 'state': {'key': key0, 'pos': pos0}}
    out = np.empty(need, dtype=np.int8)
    have, consumed = 0, 0
    while have < need:
        block = max(4096, int((need - have) * (2 ** k / time) * 1.05) + 64)
        raw = bg.random_raw(block).astype(np.uint64)
        r = (raw >> np.uint64(32 - k)).astype(np.int64)
        ok = np.flatnonzero(r < time)
        take = min(ok.size, need - have)
        out[have:have + take] = r[ok[:take]]
        have += take
        consumed += (int(ok[take - 1]) + 1) if have == need else block
    bg.state = {'bit_generator': 'MT19937', 'state': {'key': key0, 'pos': pos0}}
    if consumed:
        bg.random_raw(consumed)
    st = bg.state['state']
    rng.setstate((version, tuple(int(v) for v in st['key']) + (int(st['pos']),), gauss))
    return out.reshape(bands, H, W, 2)


def unsampling(im, scale):
    """Block mean (reference: image_convert/IHS.py:6-12); scale 2 runs on the GPU via pan2ms."""
    im = np.asarray(im)
    H, W = im.shape
    h, w = H // scale, W // scale
    acc_t = np.float32 if im.dtype == np.float32 else np.float64
    blk = im[:h * scale, :w * scale].reshape(h, scale, w, scale)
    acc = blk[:, 0, :, 0].astype(acc_t)
    for a in range(scale):
        for b in range(scale):
            if a or b:
                acc = acc + blk[:, a, :, b].astype(acc_t)
    return (acc / acc_t(scale * scale)).astype(np.float64)


def pan2ms(pan, size):
    """2x block mean then 2x2 space-to-depth into 4 bands (reference: image_convert/IHS.py:14-19)."""
    pan = np.asarray(pan)
    if size[2] != 4 or pan.shape[0] != 4 * size[0] or pan.shape[1] != 4 * size[1]:
        raise ValueError("pan2ms: expects PAN [4H,4W] and size [H,W,4]")
    return dmf.pan2ms(pan).cpu().numpy()


def unpooling(pic, time):
    """Zero-stuffed random-phase upsampling (reference: image_convert/IHS.py:22-29)."""
    pic = np.asarray(pic)
    H, W, B = pic.shape
    offs = draw_offsets(H, W, B, time)
    up = np.zeros([H * time, W * time, B])
    jj, kk = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    for i in range(B):
        up[time * jj + offs[i, :, :, 0], time * kk + offs[i, :, :, 1], i] = pic[:, :, i]
    return up


def raw_3copy(image_raw, n):
    return np.repeat(np.asarray(image_raw)[:, :, np.newaxis], n, axis=2)


def IHS_tran(MS, PAN, device='cuda:0', return_tensor=False):
    """Intensity substitution (reference: image_convert/IHS.py:40-54): float64 [4H,4W]."""
    MS = np.asarray(MS, dtype=np.float64)
    PAN = np.asarray(PAN, dtype=np.float64)
    H, W, B = MS.shape
    if B != 4 or PAN.shape != (4 * H, 4 * W):
        raise ValueError("IHS_tran: expects MS [H,W,4] and PAN [4H,4W]")
    offs = draw_offsets(H, W, B, B)
    out = dmf.ihs_tran(MS, PAN, offs, device)
    return out if return_tensor else out.cpu().numpy()
