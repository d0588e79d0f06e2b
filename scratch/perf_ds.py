"""
Down-sampler timing on one C2 band (uint16 10000 x 10000 -> 500 x 500) and a float32 60000-wide slab:
scratch/perf_ds.py   (HOMONIM_B200_LIB selects a library variant)
"""
import os
import sys
import torch
sys.path.insert(0, '.')
from homonim_b200 import Affine, kernel_model as km
from homonim_b200.kernel_model import KernelTimer

tag = os.path.basename(os.environ.get('HOMONIM_B200_LIB', 'default'))
for dtype, hs, ws, ratio in (('uint16', 10000, 10000, 20), ('uint8', 10000, 10000, 20), ('float32', 7500, 60000, 20)):
    if dtype == 'float32':
        src = torch.rand((hs, ws), device='cuda')
    else:
        src = torch.randint(1, 200, (hs, ws), device='cuda', dtype=torch.int32).to(getattr(torch, dtype))
    stf = Affine(0.5, 0, 0, 0, -0.5, 0)
    dtf = Affine(0.5 * ratio, 0, 0, 0, -0.5 * ratio, 0)
    for _ in range(3):
        km._downsample_average(src, stf, 0 if dtype != 'float32' else float('nan'), (hs // ratio, ws // ratio), dtf)
    with KernelTimer() as t:
        for _ in range(20):
            km._downsample_average(src, stf, 0 if dtype != 'float32' else float('nan'), (hs // ratio, ws // ratio), dtf)
        res = t.results()['hb_downsample_average']
    ms = sorted(res)[len(res) // 2]
    print(f'{tag:24s} {dtype:8s} {hs}x{ws}: {ms * 1e3:7.1f} us  {src.numel() * src.element_size() / ms / 1e6:7.0f} GB/s')
