import sys, torch, time
sys.path.insert(0, '.')
from homonim_b200 import *
from homonim_b200.kernel_model import KernelTimer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device='cuda').manual_seed(0)
src = torch.rand((n, n), generator=g, device='cuda') * 0.5 + 0.2
ref = 0.7 * src + 0.05 + 0.01 * torch.rand((n, n), generator=g, device='cuda')
src[100:300, 500:900] = float('nan')
crs = CRS.from_epsg(32735); tf = Affine(1, 0, 0, 0, -1, 0)
s_ra, r_ra = RasterArray(src, crs, tf), RasterArray(ref, crs, tf)
for model, k, r2 in ((Model.gain_offset, (31, 31), False), (Model.gain_offset, (15, 15), True), (Model.gain, (5, 5), False), (Model.gain_blk_offset, (15, 15), False)):
    km = KernelModel(model, k, find_r2=r2, r2_inpaint_thresh=None)
    for _ in range(2): km.fit(s_ra, r_ra)
    with KernelTimer() as t:
        for _ in range(3): p = km.fit(s_ra, r_ra)
        res = t.results()
    ms = sum(res['hb_fit_same_grid']) / 3
    nb = 8 + 4 * (3 if r2 else 2)
    print(model.value, k, 'r2' if r2 else '', f'fit {ms:.3f} ms  {n*n/ms/1e6:.1f} Gpix/s  {n*n*nb/ms/1e6:.0f} GB/s ({100*n*n*nb/ms/1e6/6547:.1f}% of peak)', {kk: round(sum(v)/3, 3) for kk, v in res.items()})
