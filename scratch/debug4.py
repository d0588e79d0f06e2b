import sys, numpy as np, torch, cv2
sys.path.insert(0, '.')
from homonim_b200 import *
from homonim_b200.synthetic import make_pair
from oracle import kernel_model_np as kmnp
NAN=float('nan')
src_ra, ref_ra = make_pair(1000, 1203, 1, bands=1, dtype='float32', mu=0.3, seed=11, device='cuda', src_nodata=NAN, ref_pad=0)
s_ra = RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=NAN); r_ra = RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, src_ra.transform, nodata=NAN)
km = KernelModel(Model.gain_blk_offset, (5,5), find_r2=True)
p = km.fit(s_ra, r_ra).to_host().array
sn32, rn32 = s_ra.to_host().array, r_ra.to_host().array
exp = kmnp.fit_same_grid(sn32, NAN, rn32, NAN, 'gain-blk-offset', (5,5), True, None)
m = ~np.isnan(sn32)&~np.isnan(rn32)
norm = kmnp.block_norm(sn32, rn32, m)
box=lambda x: cv2.boxFilter(x,-1,(5,5),normalize=False,borderType=cv2.BORDER_CONSTANT)
def exact_r2(n0, n1, g0=None):
    sp = sn32.astype('f8')*n0+n1; sp[~m]=0; rr=rn32.astype('f8').copy(); rr[~m]=0
    N=box(m.astype('f8')); S=box(sp); R=box(rr); P=box(sp*rr); S2=box(sp*sp); R2=box(rr*rr)
    g = R/S if g0 is None else g0
    rss=(g*g*S2-2*g*P+R2)*N; tss=N*R2-R*R
    return np.where(m, 1-rss/tss, np.nan), g
r2a, g_a = exact_r2(norm[0], norm[1])
r2b, g_b = exact_r2(0.78105617, -0.02534578)
d=lambda a,b: np.nanmax(np.abs(a-b))
print('exact r2: numpy-norm vs gpu-norm', d(r2a, r2b))
print('oracle vs exact(numpy norm)', d(exp[2], r2a), ' gpu vs exact(gpu norm)', d(p[2], r2b), 'gpu vs exact(numpy norm)', d(p[2], r2a), 'oracle vs gpu', d(exp[2], p[2]))
# effect of f32 rounding of g0 on r2
g32 = (exp[0]/np.float32(norm[0])).astype('f4')
r2c,_ = exact_r2(norm[0], norm[1], g0=g_a.astype('f4').astype('f8'))
print('exact r2 with f32-rounded g0 vs exact', d(r2c, r2a))
i=np.unravel_index(np.nanargmax(np.abs(exp[2]-p[2])), p[2].shape); print('worst', i, exp[2][i], p[2][i], r2a[i], r2b[i], 'gain', exp[0][i], p[0][i])
print('-----')
dd=np.abs(exp[2].astype('f8')-p[2]); ok=np.isfinite(dd)&(np.abs(exp[0])<5)&(dd>1e-4)
ys,xs=np.where(ok); print('n moderate-gain px with r2 diff>1e-4:', len(ys))
sp = sn32.astype('f8')*norm[0]+norm[1]; sp[~m]=0; rr=rn32.astype('f8').copy(); rr[~m]=0
N=box(m.astype('f8')); S=box(sp); R=box(rr); P=box(sp*rr); S2=box(sp*sp); R2=box(rr*rr)
for y,x in list(zip(ys,xs))[:8]:
    print((y,x),'r2 oracle',exp[2][y,x],'gpu',p[2][y,x],'exactA',r2a[y,x],'exactB',r2b[y,x],'gain',exp[0][y,x],p[0][y,x],'N',N[y,x],'S',S[y,x],'R',R[y,x],'tss',N[y,x]*R2[y,x]-R[y,x]**2, 'rss', (g_a[y,x]**2*S2[y,x]-2*g_a[y,x]*P[y,x]+R2[y,x]))
