import sys, numpy as np, torch, cv2
sys.path.insert(0, '.')
from homonim_b200 import *
from homonim_b200 import _native, kernel_model as hkm
from homonim_b200.synthetic import make_pair
from oracle import kernel_model_np as kmnp, gdal_restate as gr
NAN=float('nan')
src_ra, ref_ra = make_pair(120, 101, 20, bands=2, dtype='uint16', mu=3000.0, seed=5, device='cuda', src_nodata=0)
src, ref = src_ra.to_host(), ref_ra.to_host()
src_blk, src_blk_tf, ref_blk, ref_blk_tf, _ = kmnp.block_windows(src.array[0], tuple(src.transform), 0, ref.array[0], tuple(ref.transform))
ds=gr.reproject_array(src_blk, src_blk_tf, 0, ref_blk.shape, ref_blk_tf, NAN, 'average')
ds_gpu = hkm._downsample_average(src_ra.array[0].contiguous(), src_ra.transform, 0, ref_blk.shape, ref_blk_tf).cpu().numpy()
print('downsample identical', np.array_equal(ds, ds_gpu, equal_nan=True))
km = KernelModel(Model.gain, (1,1), find_r2=True)
crs=src_ra.crs
p = km.fit(RasterArray(ds.copy(), crs, ref_blk_tf), RasterArray(ref_blk.copy(), crs, ref_blk_tf)).array
exp = kmnp.fit_same_grid(ds, NAN, ref_blk, NAN, 'gain', (1,1), True, None)
d=np.abs(p[2].astype('f8')-exp[2]); d[np.isnan(d)]=0
i=np.unravel_index(np.argmax(d), d.shape); print('worst', i, p[:,i[0],i[1]], exp[:,i[0],i[1]], 'ndiff', (d>0).sum())
s=np.float32(ds[i]); r=np.float32(ref_blk[i]); print('s,r', s.view(np.uint32), r.view(np.uint32), repr(s), repr(r))
g=np.float32(r)/np.float32(s)
S2=np.float64(s)*np.float64(s); R2=np.float64(r)*np.float64(r); P=np.float32(s*r); N=np.float32(1)
tss=(N*R2)-np.float64(np.float32(r*r))
t1=np.float64(np.float32(g*g))*S2; t3=np.float64(np.float32(np.float32(2*g)*P))
rss=((t1-t3)+R2)*np.float64(N)
print('manual r2', np.float32(1)-np.float32(rss/tss), 'tss', tss, 'rss', rss, 't1', repr(t1), 't3', repr(t3), 'R2', repr(R2))
# neighbours (running-sum effects): previous rows in the same column
print('col values above', ds[max(0,i[0]-3):i[0]+2, i[1]], ref_blk[max(0,i[0]-3):i[0]+2, i[1]])
