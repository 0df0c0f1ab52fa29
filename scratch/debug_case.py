import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vsb200
from oracle import oracle as og, pipeline as op
from tests.test_gpu_parity import _rigs
case = sys.argv[1]
og.set_num_threads(8)
orig, grig, kw = _rigs(case, False)
frames = [vsb200.synth.frame(i, 0, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
want, wm = orig.compose(frames)
got = grig.compose([frames])[0]
d = (got != want).any(axis=2)
ys, xs = np.nonzero(d)
print('roi', orig.roi_final, orig.roi_padded, 'nb', orig.num_bands)
for i in range(kw['n_views']): print(i, orig.rois[i], orig.blender.view_geom(i))
print('ndiff px', d.sum(), 'x range', xs.min(), xs.max(), 'y range', ys.min(), ys.max())
print('x hist (64 px tiles):', np.bincount(xs // 64, minlength=16))
print('y hist (32 px tiles):', np.bincount(ys // 32, minlength=6))
print('max abs', np.abs(got.astype(int) - want).max())
# dw compare
for k in range(orig.num_bands + 1):
    acc = None
    W, H = orig.roi_padded[2] >> k, orig.roi_padded[3] >> k
    acc = np.zeros((H, W), np.float32)
    for i in range(kw['n_views']):
        g = orig.blender.view_geom(i)
        w = orig.blender.view_weight(i, k)
        acc[g['y_tl'] >> k:(g['y_br'] >> k), g['x_tl'] >> k:(g['x_br'] >> k)] += w
    dw = grig.read(3, 0, k, (H, W), np.float32)
    print('dw level', k, (dw != acc).sum())
k = 0
W, H = orig.roi_padded[2], orig.roi_padded[3]
acc = np.zeros((H, W), np.float32)
contrib = []
for i in range(kw['n_views']):
    g = orig.blender.view_geom(i)
    w = np.zeros((H, W), np.float32)
    w[g['y_tl']:g['y_br'], g['x_tl']:g['x_br']] = orig.blender.view_weight(i, 0)
    contrib.append(w)
    acc += w
dw = grig.read(3, 0, 0, (H, W), np.float32)
ys, xs = np.nonzero(dw != acc)
print('dw diff positions x range', xs.min(), xs.max(), 'y', ys.min(), ys.max())
for j in range(0, len(ys), max(1, len(ys)//12)):
    y, x = ys[j], xs[j]
    print((y, x), 'gpu', dw[y, x], 'np', acc[y, x], 'contribs', [float(c[y, x]) for c in contrib])
miss = dw != acc
print('total mismatches', miss.sum())
for i, c in enumerate(contrib):
    m = miss & (c != 0)
    if m.any():
        yy, xx = np.nonzero(m)
        print('view', i, 'n', m.sum(), 'bbox x', xx.min(), xx.max(), 'y', yy.min(), yy.max(), 'of nonzero', (c != 0).sum())
# does gpu == sum without view v?
for i in range(len(contrib)):
    alt = np.zeros_like(acc)
    for j, c in enumerate(contrib):
        if j != i: alt += c
    print('without view', i, 'explains', (miss & (dw == alt)).sum())
yy, xx = np.nonzero(miss)
print(np.bincount(yy, minlength=H)[40:160])
