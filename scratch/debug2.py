import sys, numpy as np, torch, cv2
sys.path.insert(0, '.')
from homonim_b200 import *
from homonim_b200 import _native, kernel_model as hkm
from homonim_b200.synthetic import make_pair
from oracle import kernel_model_np as kmnp, gdal_restate as gr
NAN=float('nan')
def report(name, got, exp, extra=None):
    d=np.abs(got.astype('f8')-exp.astype('f8'))/np.maximum(np.abs(exp),1e-3)
    d[~np.isfinite(d)]=0
    i=np.unravel_index(np.argmax(d), d.shape)
    print(name, 'max rel', d.max(), 'at', i, 'got', got[i], 'exp', exp[i], 'count>1e-4', (d>1e-4).sum())
    ys,xs=np.where(d>1e-4) if d.ndim==2 else (None,None)
    if ys is not None and len(ys): print('   rows', ys.min(), ys.max(), 'cols', xs.min(), xs.max(), 'uniq rows', np.unique(ys)[:20], 'uniq cols', np.unique(xs)[:20])
    return i
# srcspace
src_ra, ref_ra = make_pair(300, 260, 2, bands=1, dtype='float32', mu=0.3, seed=3, device='cuda', src_nodata=NAN)
km = SrcSpaceModel(Model.gain_offset, (31,31), find_r2=True, r2_inpaint_thresh=None)
p = km.fit(RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=NAN), RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, ref_ra.transform, nodata=NAN)).to_host().array
src, ref = src_ra.to_host(), ref_ra.to_host()
exp = kmnp.srcspace_fit(src.array[0], tuple(src.transform), NAN, ref.array[0], tuple(ref.transform), NAN, 'gain-offset', (31,31), True, None)
for b in range(3): report(f'srcspace band{b}', p[b], exp[b])
# same with find_r2 False
km = SrcSpaceModel(Model.gain_offset, (31,31), find_r2=False, r2_inpaint_thresh=None)
p2 = km.fit(RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=NAN), RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, ref_ra.transform, nodata=NAN)).to_host().array
report('srcspace nor2 gain', p2[0], exp[0])
# exact-math check of the worst pixel
ref_us = gr.reproject_array(ref.array[0], tuple(ref.transform), NAN, src.shape, tuple(src.transform), NAN, 'cubic_spline')
s = src.array[0].astype('f8').copy(); r = ref_us.astype('f8').copy(); m = ~np.isnan(s)&~np.isnan(r); s[~m]=0; r[~m]=0
box=lambda x: cv2.boxFilter(x,-1,(31,31),normalize=False,borderType=cv2.BORDER_CONSTANT)
N=box(m.astype('f8')); S=box(s); R=box(r); P=box(s*r); S2=box(s*s)
g64=(N*P-S*R)/(N*S2-S*S)
i=report('srcspace gain vs exact64: oracle', exp[0], np.where(m,g64,np.nan))
report('srcspace gain vs exact64: gpu', p[0], np.where(m,g64,np.nan))
# 1k blk offset
src_ra, ref_ra = make_pair(1000, 1203, 1, bands=1, dtype='float32', mu=0.3, seed=11, device='cuda', src_nodata=NAN, ref_pad=0)
s_ra = RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=NAN); r_ra = RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, src_ra.transform, nodata=NAN)
km = KernelModel(Model.gain_blk_offset, (5,5), find_r2=True)
p = km.fit(s_ra, r_ra).to_host().array
exp = kmnp.fit_same_grid(s_ra.to_host().array, NAN, r_ra.to_host().array, NAN, 'gain-blk-offset', (5,5), True, None)
for b in range(3): 
    i=report(f'1k blk band{b}', p[b], exp[b])
sn, rn = s_ra.to_host().array.astype('f8'), r_ra.to_host().array.astype('f8')
m = ~np.isnan(sn)&~np.isnan(rn)
norm = kmnp.block_norm(s_ra.to_host().array, r_ra.to_host().array, m)
sp = sn*norm[0]+norm[1]; sp[~m]=0; rr=rn.copy(); rr[~m]=0
box=lambda x: cv2.boxFilter(x,-1,(5,5),normalize=False,borderType=cv2.BORDER_CONSTANT)
S=box(sp); R=box(rr)
i=report('1k blk gain0 oracle vs exact', exp[0]/np.float32(norm[0]), np.where(m, R/S, np.nan))
print('S at worst', S[i], 'R', R[i])
