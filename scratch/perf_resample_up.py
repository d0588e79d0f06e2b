"""
Plain cubic-spline up-sampling of one float32 band (SrcSpaceModel's reference up-sampling, C3: 10000^2 -> 20000^2):
scratch/perf_resample_up.py [coarse size] [ratio]
"""
import sys
import torch
sys.path.insert(0, '.')
from homonim_b200 import Affine, kernel_model as km, _native
from homonim_b200.kernel_model import KernelTimer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ratio = int(sys.argv[2]) if len(sys.argv) > 2 else 2
src = torch.rand((n, n), device='cuda') + 0.2
src[100:120, 300:340] = float('nan')
stf = Affine(1.0 * ratio, 0, 0, 0, -1.0 * ratio, 0)
dtf = Affine(1.0, 0, 0, 0, -1.0, 0)
for _ in range(2):
    out = km._resample_up(src, stf, float('nan'), (n * ratio, n * ratio), dtf, _native.HB_UP_CUBIC_SPLINE)
with KernelTimer() as t:
    for _ in range(5):
        out = km._resample_up(src, stf, float('nan'), (n * ratio, n * ratio), dtf, _native.HB_UP_CUBIC_SPLINE)
    res = t.results()['hb_resample_up']
ms = sorted(res)[len(res) // 2]
px = (n * ratio) ** 2
print(f'resample_up {n}^2 x{ratio}: {ms:.3f} ms  {px / ms / 1e6:.1f} Gpix/s  {(px * 4 + n * n * 4) / ms / 1e6:.0f} GB/s')
