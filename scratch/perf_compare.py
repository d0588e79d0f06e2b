""" hb_compare_sums throughput on two float32 planes (CUDA events, inputs larger than L2). """
import json
import torch
from homonim_b200.compare import compare_sums_device

res = {}
for n_side in (500, 10000, 16384):
    g = torch.Generator(device='cuda').manual_seed(1)
    ref = torch.randn((n_side, n_side), device='cuda', generator=g) * 300 + 900
    src = 0.7 * ref + 40 + torch.randn((n_side, n_side), device='cuda', generator=g) * 50
    src[::97, ::13] = float('nan')
    for _ in range(5):
        compare_sums_device(src, float('nan'), ref, float('nan'))
    torch.cuda.synchronize()
    k = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(k):
        compare_sums_device(src, float('nan'), ref, float('nan'))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / k
    res[n_side] = dict(ms=round(ms, 4), gbs=round(8.0 * n_side * n_side / ms / 1e6, 1))
print(json.dumps(res))
