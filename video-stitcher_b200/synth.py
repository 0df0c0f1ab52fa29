"""Synthetic camera content of SURVEY.md 8(d) (shared by tests and bench.py; no oracle dependency).

v(x, y, c) = clamp(128 + 64 sin(2 pi (x/97 + y/61 + c/3 + view/7)) + U[-32, 32]); noise seeded by
1234 + 1000*frame + view (numpy MT19937).  Gains g_i = 1 + 0.03 ((i mod 3) - 1).
"""
import math

import numpy as np


def frame(view, frame_idx, w, h):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    rs = np.random.RandomState(1234 + 1000 * frame_idx + view)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        img[..., c] = 128.0 + 64.0 * np.sin(2.0 * math.pi * (x / 97.0 + y / 61.0 + c / 3.0 + view / 7.0))
    img += rs.uniform(-32.0, 32.0, size=(h, w, 3)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def gains(n_views):
    return [1.0 + 0.03 * ((i % 3) - 1) for i in range(n_views)]


def identity_mesh(W, H, rows=10, cols=10):
    mx = np.empty((rows, cols), np.float32)
    my = np.empty((rows, cols), np.float32)
    for i in range(rows):
        for j in range(cols):
            mx[i, j] = np.float32(np.float32(j) * np.float32(W)) / np.float32(cols - 1)
            my[i, j] = np.float32(np.float32(i) * np.float32(H)) / np.float32(rows - 1)
    return mx, my


def mesh(W, H, rows=10, cols=10, phase=0.0):
    mx, my = identity_mesh(W, H, rows, cols)
    i = np.arange(rows, dtype=np.float64)[:, None]
    j = np.arange(cols, dtype=np.float64)[None, :]
    dx = 6.0 * np.sin(math.pi * i / 9 + phase) * np.cos(math.pi * j / 9)
    dy = 4.0 * np.sin(math.pi * j / 9 + phase) * np.ones_like(i)
    return (mx + dx.astype(np.float32)).astype(np.float32), (my + dy.astype(np.float32)).astype(np.float32)


def frame_nv12(view, frame_idx, w, h):
    """Synthetic NV12 wire frame (A/defs.h:10-17): (h * 3 // 2, w) uint8 -- luma from frame(), smooth chroma plus noise, with
    values outside the nominal [16, 235] / [16, 240] ranges so the saturating branches of the conversion are exercised."""
    rng = np.random.default_rng(4321 + 1000 * frame_idx + view)
    bgr = frame(view, frame_idx, w, h)
    out = np.empty((h * 3 // 2, w), np.uint8)
    out[:h] = np.clip(bgr[..., 1].astype(np.int32) + rng.integers(-40, 41, (h, w)), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[0:h // 2, 0:w // 2]
    u = 128 + 100 * np.sin(xx / 37.0 + view) + rng.integers(-30, 31, (h // 2, w // 2))
    v = 128 + 100 * np.cos(yy / 29.0 - view) + rng.integers(-30, 31, (h // 2, w // 2))
    out[h:, 0::2] = np.clip(u, 0, 255).astype(np.uint8)
    out[h:, 1::2] = np.clip(v, 0, 255).astype(np.uint8)
    return out
