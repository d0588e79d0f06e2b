"""CPU time of one RasterFuse.process() call (C2 shapes) with the GPU kept busy: python scratch/host_overhead.py"""
import sys, time, torch
sys.path.insert(0, '.')
from homonim_b200 import Model, ProcCrs, RasterFuse
from homonim_b200.synthetic import make_pair
src_ra, ref_ra = make_pair(500, 500, 20, bands=4, dtype='uint16', mu=3000.0, seed=2, device='cuda', src_nodata=0.0)
fuse = RasterFuse(src_ra, ref_ra, proc_crs=ProcCrs.ref); fuse.open()
cfg = dict(r2_inpaint_thresh=0.25)
for _ in range(5): fuse.process(model=Model.gain_offset, kernel_shape=(15, 15), model_config=cfg)
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n): fuse.process(model=Model.gain_offset, kernel_shape=(15, 15), model_config=cfg)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f'host time per process() call: {(t1-t0)/n*1e3:.3f} ms (enqueue only); incl. GPU drain: {(t2-t0)/n*1e3:.3f} ms')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): fuse.process(model=Model.gain_offset, kernel_shape=(15, 15), model_config=cfg)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
