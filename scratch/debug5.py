import sys, numpy as np, torch
sys.path.insert(0, '.')
from homonim_b200 import *
from homonim_b200 import kernel_model as hkm
NAN=float('nan')
hp = wp = 500; ratio = 20
g = torch.Generator(device='cuda').manual_seed(1)
src = torch.randint(1, 5000, (hp * ratio, wp * ratio), generator=g, device='cuda', dtype=torch.int32)
src[:777, :1234] = 0
src = src.to(torch.uint16)
crs = CRS.from_epsg(32735)
src_tf = Affine(0.5, 0, 0, 0, -0.5, 0); ref_tf = Affine(10, 0, 0, 0, -10, 0)
src_ra = RasterArray(src, crs, src_tf, nodata=0)
avg = hkm._downsample_average(src, src_tf, 0, (hp, wp), ref_tf)
ref_ra = RasterArray((2 * avg).nan_to_num(1.0), crs, ref_tf, nodata=NAN)
km = RefSpaceModel(Model.gain, (1, 1))
param_ra = km.fit(src_ra, ref_ra)
corr = km.apply(src_ra, param_ra).array
s32 = src.to(torch.int32).to(torch.float32)
bad = (corr != 2*s32) & ~torch.isnan(corr)
ys, xs = torch.where(bad)
print('n bad', int(bad.sum()), 'rows', int(ys.min()), int(ys.max()), 'cols', int(xs.min()), int(xs.max()))
print('unique rows (first 30)', torch.unique(ys)[:30].tolist()); print('unique cols', torch.unique(xs)[:30].tolist())
print('sample', [(int(y),int(x), float(corr[y,x]), float(s32[y,x])) for y,x in list(zip(ys.tolist(), xs.tolist()))[:10]])
gain=param_ra.array[0]; print('gain unique', torch.unique(gain[~torch.isnan(gain)])[:5], 'offset unique', torch.unique(param_ra.array[1][~torch.isnan(gain)])[:5])
