import sys, numpy as np, torch
sys.path.insert(0, '.')
from homonim_b200 import *
from homonim_b200 import _native, kernel_model as hkm
from homonim_b200.synthetic import make_pair
from oracle import kernel_model_np as kmnp, gdal_restate as gr
NAN=float('nan')
lib=_native.lib()
# (a) block norm on the 1k case
src_ra, ref_ra = make_pair(1000, 1203, 1, bands=1, dtype='float32', mu=0.3, seed=11, device='cuda', src_nodata=NAN, ref_pad=0)
s=src_ra.array[0].contiguous(); r=ref_ra.array[0].contiguous()
n=s.numel()
norm=torch.zeros(2,dtype=torch.float64,device='cuda'); nb=lib.hb_block_norm_workspace_bytes(n); work=torch.empty(nb,dtype=torch.uint8,device='cuda')
_native.check(lib.hb_block_norm(s.data_ptr(),1,NAN,r.data_ptr(),1,NAN,n,norm.data_ptr(),work.data_ptr(),nb,torch.cuda.current_stream().cuda_stream))
sn, rn = s.cpu().numpy(), r.cpu().numpy()
mask=~np.isnan(sn)&~np.isnan(rn)
exp=kmnp.block_norm(sn, rn, mask)
print('norm gpu', norm.cpu().numpy(), 'numpy', exp)
print('std', np.std(rn[mask]), np.std(sn[mask]), 'p1', np.percentile(rn[mask],1), np.percentile(sn[mask],1), 'n', mask.sum())
st=np.frombuffer(work.cpu().numpy()[:120].tobytes(), dtype=np.uint8)
import struct
raw=work.cpu().numpy().tobytes()
nval=struct.unpack_from('Q',raw,0)[0]; sums=struct.unpack_from('2d',raw,8); ssd=struct.unpack_from('2d',raw,24); mean=struct.unpack_from('2d',raw,40)
rank=struct.unpack_from('4Q',raw,56); prefix=struct.unpack_from('4I',raw,88); gamma=struct.unpack_from('f',raw,104)
print('state n',nval,'sums',sums,'ssd',ssd,'mean',mean,'rank',rank,'prefix',[hex(p) for p in prefix],'gamma',gamma)
def key_float(k):
    b = (k & 0x7fffffff) if (k & 0x80000000) else (~k & 0xffffffff)
    return np.frombuffer(struct.pack('I', b), dtype=np.float32)[0]
print('vals', [key_float(p) for p in prefix])
ss=np.sort(sn[mask]); rs=np.sort(rn[mask]); k=int(np.floor(0.01*(mask.sum()-1)))
print('sorted src', ss[k-1:k+3], 'ref', rs[k-1:k+3], 'k', k)
# (b) srcspace upsample
src_ra, ref_ra = make_pair(300, 260, 2, bands=1, dtype='float32', mu=0.3, seed=3, device='cuda', src_nodata=NAN)
refn=ref_ra.to_host().array[0]; 
exp_us=gr.reproject_array(refn, tuple(ref_ra.transform), NAN, src_ra.shape, tuple(src_ra.transform), NAN, 'cubic_spline')
got_us=hkm._resample_up(ref_ra.array[0].contiguous(), ref_ra.transform, NAN, src_ra.shape, src_ra.transform).cpu().numpy()
d=np.abs(got_us-exp_us); print('upsample maxdiff', np.nanmax(d), 'nan mismatch', (np.isnan(got_us)!=np.isnan(exp_us)).sum(), 'argmax', np.unravel_index(np.nanargmax(d), d.shape))
bad=np.argwhere(d>1e-6); print('n bad', len(bad), bad[:10], bad[-5:] if len(bad) else None)
# (c) refspace gain 1x1 r2
src_ra, ref_ra = make_pair(120, 101, 20, bands=2, dtype='uint16', mu=3000.0, seed=5, device='cuda', src_nodata=0)
with RasterFuse(src_ra, ref_ra) as fuse:
    corr_ra, param_ra = fuse.process(model=Model.gain, kernel_shape=(1,1), param_filename='p', model_config=dict(r2_inpaint_thresh=None))
src, ref = src_ra.to_host(), ref_ra.to_host()
exp_params, _, exp_corr = kmnp.fuse_band_blocks(src.array[0], tuple(src.transform), 0, ref.array[0], tuple(ref.transform), NAN, 'gain', (1,1), 'ref', True, None)
got=np.stack([param_ra.to_host().array[p*2] for p in range(3)])
d=np.abs(got[2].astype('f8')-exp_params[2]); i=np.unravel_index(np.nanargmax(d), d.shape); print('r2 maxdiff', np.nanmax(d), i, got[:,i[0],i[1]], exp_params[:,i[0],i[1]], 'n differing', (d>0).sum(), 'gain equal', np.array_equal(got[0],exp_params[0],equal_nan=True))
src_blk, src_blk_tf, ref_blk, ref_blk_tf, _ = kmnp.block_windows(src.array[0], tuple(src.transform), 0, ref.array[0], tuple(ref.transform))
ds=gr.reproject_array(src_blk, src_blk_tf, 0, ref_blk.shape, ref_blk_tf, NAN, 'average')
print('s,r at px', repr(ds[i]), repr(ref_blk[i]))
