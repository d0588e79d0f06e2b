"""
Block statistics timing: scratch/perf_norm.py  -- hb_block_norm on planes of several sizes (CUDA events, 5 calls each)
and the three shard-protocol level passes (hb_block_norm_partial) on a 20000 x 20000 plane (> 2^28 pixels: C5b regime).
"""
import sys
import torch
sys.path.insert(0, '.')
from homonim_b200 import KernelModel, Model
from homonim_b200.dist import NativeBlockNorm
from homonim_b200.kernel_model import KernelTimer

nan = float('nan')
km = KernelModel(Model.gain_blk_offset, (5, 5))
g = torch.Generator(device='cuda').manual_seed(0)
for (h, w) in ((400, 400), (3000, 3000), (16384, 16384), (20000, 20000)):
    src = torch.rand((h, w), generator=g, device='cuda') * 0.5 + 0.2
    ref = 0.7 * src + 0.05 + 0.01 * torch.rand((h, w), generator=g, device='cuda')
    src[10:30, 50:90] = nan
    for _ in range(2):
        norm = km._block_norm(src, nan, ref, nan)
    with KernelTimer() as t:
        for _ in range(5):
            norm = km._block_norm(src, nan, ref, nan)
        res = t.results()
    ms = sum(sum(v) for v in res.values()) / 5
    print(f'hb_block_norm {h}x{w}: {ms:.3f} ms  {h*w/ms/1e6:.1f} Gpix/s  {3*8*h*w/ms/1e6:.0f} GB/s (3 passes x 8 B)  norm={norm.cpu().tolist()}')
    if h == 20000:
        nb = NativeBlockNorm(src, nan, ref, nan, h * w, h * w, 0)
        times = {0: [], 1: [], 2: []}
        for rep in range(4):
            for level in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                msg = nb.partial(level)
                e1.record()
                nb.merge(level, msg, 1)
                torch.cuda.synchronize()
                if rep:
                    times[level].append(e0.elapsed_time(e1))
        for level, v in times.items():
            ms = sum(v) / len(v)
            print(f'   partial level {level}: {ms:.3f} ms  {8*h*w/ms/1e6:.0f} GB/s')
        print('   sharded norm', nb.norm.cpu().tolist())
    del src, ref
