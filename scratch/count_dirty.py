import sys, numpy as np, time
sys.path.insert(0, '.')
from homonim_b200.synthetic import make_pair
from oracle import kernel_model_np as kmnp
src_ra, ref_ra = make_pair(500, 500, 20, bands=1, dtype='uint16', mu=3000.0, seed=2, device='cpu', src_nodata=0.0)
src, ref = src_ra.to_host(), ref_ra.to_host()
t=time.time()
params, _, corr = kmnp.fuse_band_blocks(src.array[0], tuple(src.transform), 0.0, ref.array[0], tuple(ref.transform), float('nan'), 'gain-offset', (15, 15), 'ref', False, 0.25)
print('oracle s', time.time()-t, params.shape)
g, o = params[0], params[1]
hp, wp = g.shape
allv = ~np.isnan(g) & ~np.isnan(o)
anyv = ~np.isnan(g) | ~np.isnan(o)
pad = np.zeros((hp+6, wp+6), bool); pad[2:2+hp,2:2+wp] = allv
pada = np.zeros((hp+6, wp+6), bool); pada[2:2+hp,2:2+wp] = anyv
clean = np.ones((hp+2, wp+2), bool); 
for j in range(4):
    for i in range(4):
        clean &= pad[j:j+hp+2, i:i+wp+2]   # cell (ky,kx) idx (ky+1,kx+1): taps rows ky-1+j -> pad index ky-1+j+2 = (ky+1)+j
cand = np.zeros((hp+2, wp+2), bool)
for j in (1,2):
    for i in (1,2):
        cand |= pada[j:j+hp+2, i:i+wp+2]
dead = ~cand
dirty = ~clean & ~dead
print('cells', clean.size, 'clean', clean.sum(), 'dead', dead.sum(), 'dirty', dirty.sum())
# dynamic range
ga = np.abs(np.where(allv, g, np.nan))
print('gain stats', np.nanmin(ga), np.nanmedian(ga), np.nanmax(ga), 'neg', (g<0).sum())
