"""
Same-grid fit kernel timing (SrcSpace / C3 regime): scratch/perf_fit.py [n] [case]
cases: c3 = gain-offset 31x31 fused apply; go15r2 = gain-offset 15x15 + R2 (params); g5 = gain 5x5; gbo15 = gain-blk-offset
15x15 (fused apply, statistics excluded).  Prints ms, Gpix/s and algorithmic GB/s (12 B/px fused, 16/20 B/px params).
"""
import sys
import torch
sys.path.insert(0, '.')
from homonim_b200 import KernelModel, Model
from homonim_b200.kernel_model import KernelTimer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
only = sys.argv[2] if len(sys.argv) > 2 else None
g = torch.Generator(device='cuda').manual_seed(0)
src = torch.rand((n, n), generator=g, device='cuda') * 0.5 + 0.2
ref = 0.7 * src + 0.05 + 0.01 * torch.rand((n, n), generator=g, device='cuda')
src[100:300, 500:900] = float('nan')
nan = float('nan')
out = torch.empty_like(src)
cases = {'c3': (Model.gain_offset, (31, 31), False, True), 'go15r2': (Model.gain_offset, (15, 15), True, False),
         'g5': (Model.gain, (5, 5), False, True), 'gbo15': (Model.gain_blk_offset, (15, 15), False, True)}
for name, (model, k, r2, fused) in cases.items():
    if only and name != only:
        continue
    km = KernelModel(model, k, find_r2=r2, r2_inpaint_thresh=None)
    norm = km._block_norm(src, nan, ref, nan) if model == Model.gain_blk_offset else None
    run = (lambda: km._fit_apply_rows(src, nan, ref, nan, 0, n, norm=norm, out=out)) if fused else \
        (lambda: km._fit_planes(src, nan, ref, nan, norm=norm))
    for _ in range(2):
        run()
    with KernelTimer() as t:
        for _ in range(3):
            run()
        res = t.results()
    ms = sum(sum(v) for v in res.values()) / 3
    nb = 12 if fused else 8 + 4 * (3 if r2 else 2)
    print(name, model.value, k, 'r2' if r2 else '', 'fused' if fused else 'params',
          f'{ms:.3f} ms  {n*n/ms/1e6:.1f} Gpix/s  {n*n*nb/ms/1e6:.0f} GB/s ({100*n*n*nb/ms/1e6/6547:.1f}% of measured peak)')
