"""Time RefSpaceModel.apply (hb_upsample_apply) alone on a C2-like band: python scratch/perf_up.py [dtype] [ratio] [n]"""
import sys, math, torch
sys.path.insert(0, '.')
from homonim_b200 import Model, RasterArray, RefSpaceModel
from homonim_b200.synthetic import make_pair
dtype = sys.argv[1] if len(sys.argv) > 1 else 'uint16'
ratio = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
hp = n // ratio
nodata = float('nan') if dtype == 'float32' else 0.0
mu = 0.3 if dtype == 'float32' else (120.0 if dtype == 'uint8' else 3000.0)
src_ra, ref_ra = make_pair(hp, hp, ratio, bands=1, dtype=dtype, mu=mu, seed=2, device='cuda', src_nodata=nodata)
src1 = RasterArray(src_ra.array[0], src_ra.crs, src_ra.transform, nodata=src_ra.nodata)
ref1 = RasterArray(ref_ra.array[0], ref_ra.crs, ref_ra.transform, nodata=ref_ra.nodata)
km = RefSpaceModel(Model.gain_offset, (15, 15), r2_inpaint_thresh=0.25)
params = km.fit(src1, ref1)
out = torch.empty(src1.array.shape, dtype=torch.float32, device='cuda')
for _ in range(5):
    km.apply(src1, params, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
ev[0].record()
for i in range(20):
    km.apply(src1, params, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
px = src1.array.numel()
b = px * (src1.array.element_size() + 4)
print(f'{dtype} ratio {ratio} {n}x{n}: apply median {ts[10]*1e3:.1f} us  min {ts[0]*1e3:.1f} us  -> {b/ts[10]/1e6:.0f} GB/s (median)')
