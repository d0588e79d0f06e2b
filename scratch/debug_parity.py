"""
Where does the corrected-pixel error of a bench workload peak?  scratch/debug_parity.py [workload]
Prints the worst pixels (GPU vs the reference on band 1) with the parameters around them.
"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from homonim_b200 import Model, ProcCrs, RasterArray, RasterFuse
from homonim_b200.synthetic import make_pair

name = sys.argv[1] if len(sys.argv) > 1 else 'c4'
cfg = bench.WORKLOADS[name]
src_ra, ref_ra = make_pair(cfg['hp'], cfg['wp'], cfg['ratio'], bands=cfg['bands'], dtype=cfg['dtype'], mu=cfg['mu'],
                           seed=2, device='cuda', src_nodata=cfg['src_nodata'])
cpu = bench.CpuReference()
src_np, ref_np, src_tf, ref_tf, note = bench._cpu_sample(cfg, src_ra, ref_ra, 1)
exp_params, exp_corr = cpu.one_block(cfg, src_np[0], src_tf, ref_np[0], ref_tf, find_r2=True)
s = RasterArray(torch.from_numpy(src_np).cuda(), src_ra.crs, src_ra.transform, nodata=src_ra.nodata)
r = RasterArray(torch.from_numpy(ref_np).cuda(), ref_ra.crs, ref_ra.transform, nodata=ref_ra.nodata)
with RasterFuse(s, r, proc_crs=ProcCrs(cfg['proc_crs'])) as f:
    corr_ra, param_ra = f.process(model=Model(cfg['model']), kernel_shape=cfg['kernel_shape'],
                                  model_config=dict(r2_inpaint_thresh=cfg['r2_inpaint_thresh']), param_filename='p')
got = corr_ra.array[0].cpu().numpy()
gp = param_ra.array.cpu().numpy()
fin = np.isfinite(exp_corr)
floor = 1e-3 * np.abs(exp_corr[fin]).mean()
err = np.zeros(exp_corr.shape)
err[fin] = np.abs(got[fin].astype('f8') - exp_corr[fin]) / np.maximum(np.abs(exp_corr[fin]), floor)
ratio = cfg['ratio']
order = np.argsort(err.ravel())[::-1][:8]
print('floor', floor, 'median |gain|', np.nanmedian(np.abs(exp_params[0])))
for o in order:
    y, x = divmod(int(o), err.shape[1])
    py, px = y // ratio + 1, x // ratio + 1          # (reference padded by one pixel)
    print(f'err {err[y, x]:.3e} at ({y},{x}) got {got[y, x]:.6f} exp {exp_corr[y, x]:.6f} src {src_np[0][y, x]}',
          'gain 3x3:', np.array2string(exp_params[0][py - 1:py + 2, px - 1:px + 2], precision=3).replace('\n', ''),
          'offset centre', exp_params[1][py, px])
